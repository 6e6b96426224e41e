// Depthwise causal conv1d (+ SiLU), both time directions from one read of x.  sm_100a.
//
// Reference: src/models/modules/mamba_block.py:24-31 (Conv1d groups=d_inner, k=4, padding=k-1),
// :52-55 (crop to L, SiLU).  Direction 0 is that causal conv; direction 1 is the causal conv of
// the time-reversed sequence written back in natural order (taps t..t+K-1), which is what
// conv(flip(x)) of src/models/DualStreamSEMamba.py:476-478 computes.
//
// Each thread owns V consecutive time steps of one (batch, channel) row and produces both
// directions from one register window x[t0-(K-1) .. t0+V+K-2].
#include "common.cuh"

namespace bimamba {

void set_err(const char* msg);  // scan.cu

constexpr int kConvV = 8;
constexpr int kConvThreads = 256;
constexpr int kMaxK = 4;

template <int K>
__global__ void __launch_bounds__(kConvThreads)
conv_fwd_kernel(const void* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                void* __restrict__ out, int batch, int ndir, int dim, int L, int Lp, int64_t x_bs, int64_t x_rs,
                int64_t o_bs, int64_t o_ds, int64_t o_rs, int dt, int silu) {
  const int strips = (Lp + kConvV - 1) / kConvV;
  const int64_t total = (int64_t)batch * dim * strips;
  const int64_t gid = (int64_t)blockIdx.x * kConvThreads + threadIdx.x;
  if (gid >= total) return;
  const int s = (int)(gid % strips);
  const int64_t row = gid / strips;
  const int d = (int)(row % dim);
  const int b = (int)(row / dim);
  const int t0 = s * kConvV;

  float wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wk[k] = __ldg(w + d * K + k);
  const float bs = bias ? __ldg(bias + d) : 0.f;

  constexpr int W = kConvV + 2 * (K - 1);
  float xv[W];
  const int64_t xb = (int64_t)b * x_bs + (int64_t)d * x_rs;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int t = t0 - (K - 1) + j;
    xv[j] = (t >= 0 && t < L) ? ld_f(x, xb + t, dt) : 0.f;
  }
  for (int dir = 0; dir < ndir; ++dir) {
    const int64_t ob = (int64_t)b * o_bs + (int64_t)dir * o_ds + (int64_t)d * o_rs;
#pragma unroll
    for (int i = 0; i < kConvV; ++i) {
      const int t = t0 + i;
      if (t < L) {
        float acc = bs;
        // window index of x[t] is i + K - 1
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = dir == 0 ? (i + k) : (i + 2 * (K - 1) - k);  // x[t-(K-1)+k]  |  x[t+(K-1)-k]
          acc = fmaf(wk[k], xv[j], acc);
        }
        if (silu) acc *= sigmoid_f(acc);
        st_f(out, ob + t, acc, dt);
      } else if (t < Lp) {
        st_f(out, ob + t, 0.f, dt);
      }
    }
  }
}

// Backward.  One thread owns V consecutive time steps of dx for one (batch, channel) row.
//   pre_dir[t]  = bias + sum_k w[k] x[t -/+ ((K-1)-k)];  g_dir[t] = dout_dir[t] * silu'(pre_dir[t])
//   dx[tau]     = sum_k w[k] ( g_0[tau+(K-1)-k] + g_1[tau-(K-1)+k] )
//   dw[k]      += sum_t g_0[t] x[t-(K-1)+k] + g_1[t] x[t+(K-1)-k];   dbias += sum_t g_0[t] + g_1[t]
// A block covers `kConvThreads` strips of ONE row-major range; dw/dbias are reduced per
// (batch, channel) row by a warp-per-row layout: blockDim = (32 lanes over strips) x (8 rows).
template <int K>
__global__ void __launch_bounds__(kConvThreads)
conv_bwd_kernel(const void* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                const void* __restrict__ dout, void* __restrict__ dx, float* __restrict__ dwb_part,
                int batch, int ndir, int dim, int L, int Lp, int64_t x_bs, int64_t x_rs, int64_t g_bs, int64_t g_ds,
                int64_t g_rs, int64_t dx_bs, int64_t dx_rs, int dt, int silu) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (kConvThreads / 32) + (threadIdx.x >> 5);
  if (row >= (int64_t)batch * dim) return;  // whole warp exits together
  const int d = (int)(row % dim);
  const int b = (int)(row / dim);

  float wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wk[k] = __ldg(w + d * K + k);
  const float bs = bias ? __ldg(bias + d) : 0.f;
  const int64_t xb = (int64_t)b * x_bs + (int64_t)d * x_rs;
  const int64_t dxb = (int64_t)b * dx_bs + (int64_t)d * dx_rs;

  float dwl[K];
#pragma unroll
  for (int k = 0; k < K; ++k) dwl[k] = 0.f;
  float dbl = 0.f;

  constexpr int H = K - 1;              // halo
  constexpr int WX = kConvV + 4 * H;    // x window: [t0-2H, t0+V+2H)
  constexpr int WG = kConvV + 2 * H;    // g windows

  for (int t = L + lane; t < Lp; t += 32) st_f(dx, dxb + t, 0.f, dt);
  for (int t0 = lane * kConvV; t0 < L; t0 += 32 * kConvV) {
    float xv[WX];
#pragma unroll
    for (int j = 0; j < WX; ++j) {
      const int t = t0 - 2 * H + j;
      xv[j] = (t >= 0 && t < L) ? ld_f(x, xb + t, dt) : 0.f;
    }
    float acc[kConvV];
#pragma unroll
    for (int i = 0; i < kConvV; ++i) acc[i] = 0.f;

    for (int dir = 0; dir < ndir; ++dir) {
      const int64_t gb = (int64_t)b * g_bs + (int64_t)dir * g_ds + (int64_t)d * g_rs;
      // dir 0 needs g_0[t0 .. t0+V+H);  dir 1 needs g_1[t0-H .. t0+V)
      const int gstart = dir == 0 ? t0 : t0 - H;
      float gv[WG];
#pragma unroll
      for (int j = 0; j < kConvV + H; ++j) {
        const int t = gstart + j;
        float gval = 0.f;
        if (t >= 0 && t < L) {
          gval = ld_f(dout, gb + t, dt);
          if (silu) {
            float pre = bs;
            // x[t] sits at window index (t - t0 + 2H)
            const int c0 = t - t0 + 2 * H;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              const int j2 = dir == 0 ? (c0 - H + k) : (c0 + H - k);
              pre = fmaf(wk[k], xv[j2], pre);
            }
            const float sg = sigmoid_f(pre);
            gval *= sg * (1.f + pre * (1.f - sg));
          }
        }
        gv[j] = gval;
      }
      // dx contributions
#pragma unroll
      for (int i = 0; i < kConvV; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          // dir 0: g_0[tau+H-k] -> window index (i + H - k);  dir 1: g_1[tau-H+k] -> index (i + k)
          const int j = dir == 0 ? (i + H - k) : (i + k);
          acc[i] = fmaf(wk[k], gv[j], acc[i]);
        }
      }
      // dw / dbias contributions of the V positions this thread owns (t = t0+i)
#pragma unroll
      for (int i = 0; i < kConvV; ++i) {
        const int jg = dir == 0 ? i : (i + H);  // window index of g_dir[t0+i]
        const float gval = gv[jg];               // zero when t0+i >= L
        dbl += gval;
        const int c0 = i + 2 * H;                // window index of x[t0+i]
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j2 = dir == 0 ? (c0 - H + k) : (c0 + H - k);
          dwl[k] = fmaf(gval, xv[j2], dwl[k]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kConvV; ++i)
      if (t0 + i < L) st_f(dx, dxb + t0 + i, acc[i], dt);
  }

#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dwl[k] += __shfl_xor_sync(kFull, dwl[k], off);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) dbl += __shfl_xor_sync(kFull, dbl, off);
  if (lane == 0) {
    float* o = dwb_part + row * (K + 1);
#pragma unroll
    for (int k = 0; k < K; ++k) o[k] = dwl[k];
    o[K] = dbl;
  }
}

// out[g*out_gs + j] (+)= sum_i part[g*part_gs + i*row_stride + j]
__global__ void __launch_bounds__(256)
reduce_kernel(const float* __restrict__ part, void* __restrict__ out, int64_t groups, int64_t rows, int64_t cols,
              int64_t part_gs, int64_t row_stride, int64_t out_gs, int dt, int accumulate) {
  const int64_t total = groups * cols;
  for (int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x; gid < total; gid += (int64_t)gridDim.x * 256) {
    const int64_t g = gid / cols, j = gid - g * cols;
    const float* src = part + g * part_gs + j;
    float s = 0.f;
    for (int64_t i = 0; i < rows; ++i) s += __ldg(src + i * row_stride);
    if (accumulate) s += ld_f(out, g * out_gs + j, dt);
    st_f(out, g * out_gs + j, s, dt);
  }
}

}  // namespace bimamba

using namespace bimamba;

#define CONV_DISPATCH_K(CALL)                 \
  switch (width) {                            \
    case 2: { constexpr int K = 2; CALL; } break; \
    case 3: { constexpr int K = 3; CALL; } break; \
    default: { constexpr int K = 4; CALL; } break; \
  }

extern "C" int bimamba_causal_conv1d_fwd(const void* x, const float* weight, const float* bias, void* out, int batch,
                                         int ndir, int dim, int seqlen, int pad_to, int width, int64_t x_bs, int64_t x_rs,
                                         int64_t out_bs, int64_t out_ds, int64_t out_rs, int dtype, int flags,
                                         bimamba_stream_t stream) {
  if (batch == 0) return 0;
  if (!x || !weight || !out) { set_err("conv fwd: null operand"); return -1; }
  if (width < 2 || width > kMaxK) { set_err("conv width must be 2, 3 or 4"); return -2; }
  if (ndir < 1 || ndir > 2 || dtype < 0 || dtype > 2 || batch < 0 || dim < 1 || seqlen < 0) { set_err("conv fwd: bad sizes"); return -3; }
  if (pad_to < seqlen) pad_to = seqlen;
  if (batch == 0 || pad_to == 0) return 0;
  const int strips = (pad_to + kConvV - 1) / kConvV;
  const int64_t total = (int64_t)batch * dim * strips;
  const unsigned blocks = (unsigned)((total + kConvThreads - 1) / kConvThreads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CONV_DISPATCH_K((conv_fwd_kernel<K><<<blocks, kConvThreads, 0, st>>>(x, weight, bias, out, batch, ndir, dim, seqlen, pad_to, x_bs, x_rs,
                                                                       out_bs, out_ds, out_rs, dtype, flags & BIMAMBA_FLAG_SILU)))
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int bimamba_causal_conv1d_bwd(const void* x, const float* weight, const float* bias, const void* dout, void* dx,
                                         float* dwb_part, int batch, int ndir, int dim, int seqlen, int pad_to, int width, int64_t x_bs,
                                         int64_t x_rs, int64_t dout_bs, int64_t dout_ds, int64_t dout_rs, int64_t dx_bs,
                                         int64_t dx_rs, int dtype, int flags, bimamba_stream_t stream) {
  if (batch == 0) return 0;
  if (!x || !weight || !dout || !dx || !dwb_part) { set_err("conv bwd: null operand"); return -1; }
  if (width < 2 || width > kMaxK) { set_err("conv width must be 2, 3 or 4"); return -2; }
  if (ndir < 1 || ndir > 2 || dtype < 0 || dtype > 2 || batch < 0 || dim < 1 || seqlen < 0) { set_err("conv bwd: bad sizes"); return -3; }
  if (pad_to < seqlen) pad_to = seqlen;
  if (batch == 0) return 0;
  const int64_t rows = (int64_t)batch * dim;
  const unsigned blocks = (unsigned)((rows + kConvThreads / 32 - 1) / (kConvThreads / 32));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CONV_DISPATCH_K((conv_bwd_kernel<K><<<blocks, kConvThreads, 0, st>>>(x, weight, bias, dout, dx, dwb_part, batch, ndir, dim, seqlen, pad_to,
                                                                       x_bs, x_rs, dout_bs, dout_ds, dout_rs, dx_bs, dx_rs, dtype,
                                                                       flags & BIMAMBA_FLAG_SILU)))
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int bimamba_reduce_partials(const float* part, void* out, int64_t groups, int64_t rows, int64_t cols,
                                       int64_t part_gs, int64_t row_stride, int64_t out_gs, int out_dtype, int accumulate,
                                       bimamba_stream_t stream) {
  if (groups * cols == 0) return 0;
  if (!part || !out) { set_err("reduce: null operand"); return -1; }
  if (out_dtype < 0 || out_dtype > 2 || groups < 0 || rows < 0 || cols < 0) { set_err("reduce: bad sizes"); return -2; }
  const int64_t total = groups * cols;
  if (total == 0) return 0;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(part, out, groups, rows, cols, part_gs, row_stride, out_gs, out_dtype, accumulate);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
