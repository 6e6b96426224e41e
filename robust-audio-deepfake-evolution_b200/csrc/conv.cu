// Depthwise causal conv1d (+ SiLU), channel-last, both time directions from one read of x.  sm_100a.
//
// Reference: src/models/modules/mamba_block.py:24-31 (Conv1d groups=d_inner, k=4, padding=k-1),
// :52-55 (crop to L, SiLU).  Direction 0 is that causal conv; direction 1 is the causal conv of
// the time-reversed sequence written back in natural order (taps t..t+K-1), which is what
// conv(flip(x)) of src/models/DualStreamSEMamba.py:476-478 computes.
//
// A thread owns V consecutive channels (one 8/16-byte vector of a (batch, time) row) and walks
// kSeg consecutive time steps with a register window of 2K-1 rows, producing both directions from
// one read of x.  Consecutive threads own consecutive channel vectors, so every load and store of a
// warp is one contiguous run of the row.
#include <initializer_list>

#include "common.cuh"

namespace bimamba {

constexpr int kSeg = 16;          // time steps per thread
constexpr int kConvThreads = 128;
constexpr int kMaxK = 4;

template <typename T, int V> struct Vec;
template <> struct Vec<float, 4> { using type = float4; };
template <> struct Vec<float, 1> { using type = float; };
template <> struct Vec<__nv_bfloat16, 4> { using type = uint2; };
template <> struct Vec<__nv_bfloat16, 1> { using type = __nv_bfloat16; };
template <> struct Vec<__half, 4> { using type = uint2; };
template <> struct Vec<__half, 1> { using type = __half; };

template <typename T, int V>
__device__ __forceinline__ void load_row(const T* __restrict__ p, float (&v)[V]) {
  using VT = typename Vec<T, V>::type;
  const VT raw = *reinterpret_cast<const VT*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = to_f(e[i]);
}
template <typename T, int V>
__device__ __forceinline__ void store_row(T* __restrict__ p, const float (&v)[V]) {
  using VT = typename Vec<T, V>::type;
  VT raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) e[i] = from_f<T>(v[i]);
  *reinterpret_cast<VT*>(p) = raw;
}

__device__ __forceinline__ float silu_grad(float pre) {
  const float sg = sigmoid_f(pre);
  return sg * (1.f + pre * (1.f - sg));
}

template <typename T, int V, int K>
__global__ void __launch_bounds__(kConvThreads)
conv_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
                int batch, int ndir, int dim, int L, int64_t x_bs, int64_t x_ts, int64_t o_bs, int64_t o_ds,
                int64_t o_ts, int silu) {
  const int nvec = (dim + V - 1) / V;
  const int nseg = (L + kSeg - 1) / kSeg;
  const int64_t total = (int64_t)batch * nseg * nvec;
  const int64_t gid = (int64_t)blockIdx.x * kConvThreads + threadIdx.x;
  if (gid >= total) return;
  const int v = (int)(gid % nvec);
  const int s = (int)((gid / nvec) % nseg);
  const int b = (int)(gid / ((int64_t)nvec * nseg));
  const int d = v * V, t0 = s * kSeg;
  constexpr int H = K - 1;

  float wk[V][K], bs[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) wk[i][k] = __ldg(w + (int64_t)(d + i) * K + k);
    bs[i] = bias ? __ldg(bias + d + i) : 0.f;
  }
  const T* xb = x + (int64_t)b * x_bs + d;
  // window rows: win[j] = x[t - H + j], j = 0..2H, for the current t
  float win[2 * H + 1][V];
#pragma unroll
  for (int j = 0; j < 2 * H; ++j) {
    const int t = t0 - H + j;
    if (t >= 0 && t < L) {
      load_row<T, V>(xb + (int64_t)t * x_ts, win[j + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) win[j + 1][i] = 0.f;
    }
  }
#pragma unroll
  for (int s_ = 0; s_ < kSeg; ++s_) {
    const int t = t0 + s_;
    if (t >= L) break;
#pragma unroll
    for (int j = 0; j < 2 * H; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) win[j][i] = win[j + 1][i];
    const int tn = t + H;
    if (tn < L) {
      load_row<T, V>(xb + (int64_t)tn * x_ts, win[2 * H]);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) win[2 * H][i] = 0.f;
    }
    float o0[V], o1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float a0 = bs[i], a1 = bs[i];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        a0 = fmaf(wk[i][k], win[k][i], a0);           // x[t-H+k]
        a1 = fmaf(wk[i][k], win[2 * H - k][i], a1);   // x[t+H-k]
      }
      if (silu) {
        a0 *= sigmoid_f(a0);
        a1 *= sigmoid_f(a1);
      }
      o0[i] = a0;
      o1[i] = a1;
    }
    T* ob = out + (int64_t)b * o_bs + (int64_t)t * o_ts + d;
    store_row<T, V>(ob, o0);
    if (ndir > 1) store_row<T, V>(ob + o_ds, o1);
  }
}

// Backward.  Thread = (batch b, time slot y, channel vector v); it walks the segments
// s = y, y+SY, ... of its row (SY = time slots per batch row, conv_sy()).  For a segment [t0, t0+kSeg):
//   pre_dir[t] = bias + sum_k w[k] x[t -/+ (H-k)];  g_dir[t] = dout_dir[t] * silu'(pre_dir[t])
//   dx[tau]    = sum_k w[k] ( g_0[tau+H-k] + g_1[tau-H+k] )
//   dw[k]     += sum_t g_0[t] x[t-H+k] + g_1[t] x[t+H-k];   dbias += sum_t g_0[t] + g_1[t]
// (dw / dbias over the thread's OWN positions only; halo positions are recomputed for dx.)
template <typename T, int V, int K>
__global__ void __launch_bounds__(kConvThreads)
conv_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                const T* __restrict__ dout, T* __restrict__ dx, const T* __restrict__ dz_in, T* __restrict__ dz_out,
                float* __restrict__ part, int batch, int ndir, int dim, int L, int SY, int64_t x_bs, int64_t x_ts,
                int64_t g_bs, int64_t g_ds, int64_t g_ts, int64_t dx_bs, int64_t dx_ts, int silu) {
  const int nvec = (dim + V - 1) / V;
  const int nseg = (L + kSeg - 1) / kSeg;
  const int64_t total = (int64_t)batch * SY * nvec;
  const int64_t gid = (int64_t)blockIdx.x * kConvThreads + threadIdx.x;
  if (gid >= total) return;
  const int v = (int)(gid % nvec);
  const int y = (int)((gid / nvec) % SY);
  const int b = (int)(gid / ((int64_t)nvec * SY));
  const int d = v * V;
  constexpr int H = K - 1;

  float wk[V][K], bs[V], dwl[V][K], dbl[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      wk[i][k] = __ldg(w + (int64_t)(d + i) * K + k);
      dwl[i][k] = 0.f;
    }
    bs[i] = bias ? __ldg(bias + d + i) : 0.f;
    dbl[i] = 0.f;
  }
  const T* xb = x + (int64_t)b * x_bs + d;
  const T* gb = dout + (int64_t)b * g_bs + d;
  T* dxb = dx + (int64_t)b * dx_bs + d;

  for (int s = y; s < nseg; s += SY) {
    const int t0 = s * kSeg;
    // Walk tau = t0-H .. t0+kSeg+H-1.  At each tau we have the x window xw[j] = x[tau-H+j] (j=0..2H),
    // form g0[tau], g1[tau], and scatter them into the dx accumulators of the positions they touch:
    //   g0[tau] contributes w[k] g0[tau] to dx[tau-H+k];   g1[tau] contributes w[k] g1[tau] to dx[tau+H-k]
    // dx accumulators form a sliding window acc[j] = dx[tau-H+j], j = 0..2H; dx[tau-H] is complete
    // once tau has been processed (g0 reaches back H, g1 reaches forward H).
    float xw[2 * H + 1][V], acc[2 * H + 1][V];
#pragma unroll
    for (int j = 0; j < 2 * H + 1; ++j)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * H; ++j) {
      const int t = t0 - 2 * H + j;
      if (t >= 0 && t < L) {
        load_row<T, V>(xb + (int64_t)t * x_ts, xw[j + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) xw[j + 1][i] = 0.f;
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < kSeg + 2 * H; ++s_) {
      const int tau = t0 - H + s_;
#pragma unroll
      for (int j = 0; j < 2 * H; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          xw[j][i] = xw[j + 1][i];
          acc[j][i] = acc[j + 1][i];
        }
#pragma unroll
      for (int i = 0; i < V; ++i) acc[2 * H][i] = 0.f;
      const int tn = tau + H;
      if (tn >= 0 && tn < L) {
        load_row<T, V>(xb + (int64_t)tn * x_ts, xw[2 * H]);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) xw[2 * H][i] = 0.f;
      }
      if (tau >= 0 && tau < L) {
        float g0[V], g1[V];
        load_row<T, V>(gb + (int64_t)tau * g_ts, g0);
        if (ndir > 1) {
          load_row<T, V>(gb + g_ds + (int64_t)tau * g_ts, g1);
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) g1[i] = 0.f;
        }
        const bool own = tau >= t0 && tau < t0 + kSeg;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          if (silu) {
            float p0 = bs[i], p1 = bs[i];
#pragma unroll
            for (int k = 0; k < K; ++k) {
              p0 = fmaf(wk[i][k], xw[k][i], p0);
              p1 = fmaf(wk[i][k], xw[2 * H - k][i], p1);
            }
            g0[i] *= silu_grad(p0);
            g1[i] *= silu_grad(p1);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            acc[k][i] = fmaf(wk[i][k], g0[i], acc[k][i]);                  // dx[tau-H+k]
            acc[2 * H - k][i] = fmaf(wk[i][k], g1[i], acc[2 * H - k][i]);  // dx[tau+H-k]
          }
          if (own) {
            dbl[i] += g0[i] + g1[i];
#pragma unroll
            for (int k = 0; k < K; ++k) dwl[i][k] += g0[i] * xw[k][i] + g1[i] * xw[2 * H - k][i];
          }
        }
      }
      // dx[tau-H] is complete (for positions inside this segment)
      const int td = tau - H;
      if (td >= t0 && td < t0 + kSeg && td < L) {
        store_row<T, V>(dxb + (int64_t)td * dx_ts, acc[0]);
        if (dz_in) {
          float z0[V], z1[V];
          load_row<T, V>(dz_in + (int64_t)b * g_bs + (int64_t)td * g_ts + d, z0);
          if (ndir > 1) {
            load_row<T, V>(dz_in + (int64_t)b * g_bs + g_ds + (int64_t)td * g_ts + d, z1);
#pragma unroll
            for (int i = 0; i < V; ++i) z0[i] += z1[i];
          }
          store_row<T, V>(dz_out + (int64_t)b * dx_bs + (int64_t)td * dx_ts + d, z0);
        }
      }
    }
  }
  float* o = part + ((int64_t)(b * SY + y) * dim + d) * (K + 1);
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) o[i * (K + 1) + k] = dwl[i][k];
    o[i * (K + 1) + K] = dbl[i];
  }
}

// out[g*out_gs + j] (+)= sum_i part[g*part_gs + i*row_stride + j]
// Block = (256 / RL columns) x (RL row lanes); row lane r sums rows r, r+RL, ... in order and the
// lanes are combined in order through shared memory, so the result does not depend on timing.
template <int RL>
__global__ void __launch_bounds__(256)
reduce_kernel(const float* __restrict__ part, void* __restrict__ out, int64_t groups, int64_t rows, int64_t cols,
              int64_t part_gs, int64_t row_stride, int64_t out_gs, int dt, int accumulate) {
  constexpr int CPB = 256 / RL;
  __shared__ float sm[RL][CPB + 1];
  const int cx = threadIdx.x % CPB, ry = threadIdx.x / CPB;
  const int64_t tiles = (cols + CPB - 1) / CPB;
  for (int64_t bid = blockIdx.x; bid < groups * tiles; bid += gridDim.x) {
    const int64_t g = bid / tiles, j = (bid - g * tiles) * CPB + cx;
    float s = 0.f;
    if (j < cols) {
      const float* src = part + g * part_gs + j;
      for (int64_t i = ry; i < rows; i += RL) s += __ldg(src + i * row_stride);
    }
    if (RL > 1) {
      sm[ry][cx] = s;
      __syncthreads();
      if (ry == 0) {
#pragma unroll
        for (int r = 1; r < RL; ++r) s += sm[r][cx];
      }
    }
    if (ry == 0 && j < cols) {
      if (accumulate) s += ld_f(out, g * out_gs + j, dt);
      st_f(out, g * out_gs + j, s, dt);
    }
    if (RL > 1) __syncthreads();
  }
}

// time slots per batch row in the backward: enough threads to fill the GPU, few enough partials
// Column sums of a (rows, cols) activation matrix (bias gradients of the Linear layers): stage 1 writes one fp32
// partial row per slice of kColRows rows; bimamba_reduce_partials finishes the sum in fixed order.
constexpr int kColRows = 256;
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, float* __restrict__ part, int64_t rows, int cols, int64_t ld) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * kColRows;
  float s = 0.f;
  if (col < cols) {
#pragma unroll 4
    for (int i = ry; i < kColRows; i += 8) {
      const int64_t r = r0 + i;
      if (r < rows) s += to_f(x[r * ld + col]);
    }
  }
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && col < cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += sm[k][cx];
    part[(int64_t)blockIdx.y * cols + col] = s;
  }
}

static int conv_sy(int batch, int seqlen, int dim) {
  const int nseg = (seqlen + kSeg - 1) / kSeg;
  const int64_t per_slot = (int64_t)(batch < 1 ? 1 : batch) * ((dim + 3) / 4);
  int64_t sy = (150000 + per_slot - 1) / per_slot;
  if (sy > nseg) sy = nseg;
  return (int)(sy < 1 ? 1 : sy);
}

template <typename T, int V>
static void launch_conv_fwd(const void* x, const float* w, const float* bias, void* out, int batch, int ndir, int dim,
                            int L, int width, int64_t x_bs, int64_t x_ts, int64_t o_bs, int64_t o_ds, int64_t o_ts,
                            int silu, cudaStream_t st) {
  const int nvec = (dim + V - 1) / V, nseg = (L + kSeg - 1) / kSeg;
  const int64_t total = (int64_t)batch * nseg * nvec;
  const unsigned blocks = (unsigned)((total + kConvThreads - 1) / kConvThreads);
  const T* xp = reinterpret_cast<const T*>(x);
  T* op = reinterpret_cast<T*>(out);
  switch (width) {
    case 2: conv_fwd_kernel<T, V, 2><<<blocks, kConvThreads, 0, st>>>(xp, w, bias, op, batch, ndir, dim, L, x_bs, x_ts, o_bs, o_ds, o_ts, silu); break;
    case 3: conv_fwd_kernel<T, V, 3><<<blocks, kConvThreads, 0, st>>>(xp, w, bias, op, batch, ndir, dim, L, x_bs, x_ts, o_bs, o_ds, o_ts, silu); break;
    default: conv_fwd_kernel<T, V, 4><<<blocks, kConvThreads, 0, st>>>(xp, w, bias, op, batch, ndir, dim, L, x_bs, x_ts, o_bs, o_ds, o_ts, silu); break;
  }
}

template <typename T, int V>
static void launch_conv_bwd(const void* x, const float* w, const float* bias, const void* dout, void* dx,
                            const void* dz_in, void* dz_out, float* part, int batch, int ndir, int dim, int L, int width,
                            int64_t x_bs, int64_t x_ts, int64_t g_bs, int64_t g_ds, int64_t g_ts, int64_t dx_bs,
                            int64_t dx_ts, int silu, cudaStream_t st) {
  const int nvec = (dim + V - 1) / V, SY = conv_sy(batch, L, dim);
  const int64_t total = (int64_t)batch * SY * nvec;
  const unsigned blocks = (unsigned)((total + kConvThreads - 1) / kConvThreads);
  const T* xp = reinterpret_cast<const T*>(x);
  const T* gp = reinterpret_cast<const T*>(dout);
  const T* zi = reinterpret_cast<const T*>(dz_in);
  T* dxp = reinterpret_cast<T*>(dx);
  T* zo = reinterpret_cast<T*>(dz_out);
  switch (width) {
    case 2: conv_bwd_kernel<T, V, 2><<<blocks, kConvThreads, 0, st>>>(xp, w, bias, gp, dxp, zi, zo, part, batch, ndir, dim, L, SY, x_bs, x_ts, g_bs, g_ds, g_ts, dx_bs, dx_ts, silu); break;
    case 3: conv_bwd_kernel<T, V, 3><<<blocks, kConvThreads, 0, st>>>(xp, w, bias, gp, dxp, zi, zo, part, batch, ndir, dim, L, SY, x_bs, x_ts, g_bs, g_ds, g_ts, dx_bs, dx_ts, silu); break;
    default: conv_bwd_kernel<T, V, 4><<<blocks, kConvThreads, 0, st>>>(xp, w, bias, gp, dxp, zi, zo, part, batch, ndir, dim, L, SY, x_bs, x_ts, g_bs, g_ds, g_ts, dx_bs, dx_ts, silu); break;
  }
}

static bool vec4_ok(int dtype, int dim, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> strides) {
  const int es = dtype == BIMAMBA_F32 ? 4 : 2;
  if (dim % 4) return false;
  for (const void* p : ptrs)
    if (p && (reinterpret_cast<uintptr_t>(p) % (4 * es))) return false;
  for (int64_t s : strides)
    if (s % 4) return false;
  return true;
}

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_causal_conv1d_fwd(const void* x, const float* weight, const float* bias, void* out, int batch,
                                         int ndir, int dim, int seqlen, int width, int64_t x_bs, int64_t x_ts,
                                         int64_t out_bs, int64_t out_ds, int64_t out_ts, int dtype, int flags,
                                         bimamba_stream_t stream) {
  if (batch == 0 || seqlen == 0) return 0;
  if (!x || !weight || !out) { set_err("conv fwd: null operand"); return -1; }
  if (width < 2 || width > kMaxK) { set_err("conv width must be 2, 3 or 4"); return -2; }
  if (ndir < 1 || ndir > 2 || dtype < 0 || dtype > 2 || batch < 0 || dim < 1 || seqlen < 0) { set_err("conv fwd: bad sizes"); return -3; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int silu = flags & BIMAMBA_FLAG_SILU;
  const bool v4 = vec4_ok(dtype, dim, {x, out}, {x_bs, x_ts, out_bs, out_ds, out_ts});
#define CONV_FWD(T, V) launch_conv_fwd<T, V>(x, weight, bias, out, batch, ndir, dim, seqlen, width, x_bs, x_ts, out_bs, out_ds, out_ts, silu, st)
  if (dtype == BIMAMBA_F32) { if (v4) CONV_FWD(float, 4); else CONV_FWD(float, 1); }
  else if (dtype == BIMAMBA_BF16) { if (v4) CONV_FWD(__nv_bfloat16, 4); else CONV_FWD(__nv_bfloat16, 1); }
  else { if (v4) CONV_FWD(__half, 4); else CONV_FWD(__half, 1); }
#undef CONV_FWD
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int bimamba_conv_bwd_slices(int batch, int seqlen, int dim) { return batch * conv_sy(batch, seqlen, dim); }

extern "C" int bimamba_causal_conv1d_bwd(const void* x, const float* weight, const float* bias, const void* dout, void* dx,
                                         const void* dz_in, void* dz_out, float* dwb_part, int batch, int ndir, int dim,
                                         int seqlen, int width, int64_t x_bs, int64_t x_ts, int64_t dout_bs,
                                         int64_t dout_ds, int64_t dout_ts, int64_t dx_bs, int64_t dx_ts, int dtype,
                                         int flags, bimamba_stream_t stream) {
  if (batch == 0 || seqlen == 0) return 0;
  if (!x || !weight || !dout || !dx || !dwb_part) { set_err("conv bwd: null operand"); return -1; }
  if ((dz_in == nullptr) != (dz_out == nullptr)) { set_err("conv bwd: dz_in and dz_out go together"); return -1; }
  if (width < 2 || width > kMaxK) { set_err("conv width must be 2, 3 or 4"); return -2; }
  if (ndir < 1 || ndir > 2 || dtype < 0 || dtype > 2 || batch < 0 || dim < 1 || seqlen < 0) { set_err("conv bwd: bad sizes"); return -3; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int silu = flags & BIMAMBA_FLAG_SILU;
  const bool v4 = vec4_ok(dtype, dim, {x, dout, dx, dz_in, dz_out}, {x_bs, x_ts, dout_bs, dout_ds, dout_ts, dx_bs, dx_ts});
#define CONV_BWD(T, V) launch_conv_bwd<T, V>(x, weight, bias, dout, dx, dz_in, dz_out, dwb_part, batch, ndir, dim, seqlen, width, x_bs, x_ts, dout_bs, dout_ds, dout_ts, dx_bs, dx_ts, silu, st)
  if (dtype == BIMAMBA_F32) { if (v4) CONV_BWD(float, 4); else CONV_BWD(float, 1); }
  else if (dtype == BIMAMBA_BF16) { if (v4) CONV_BWD(__nv_bfloat16, 4); else CONV_BWD(__nv_bfloat16, 1); }
  else { if (v4) CONV_BWD(__half, 4); else CONV_BWD(__half, 1); }
#undef CONV_BWD
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
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto nblocks = [&](int cpb) {
    int64_t n = groups * ((cols + cpb - 1) / cpb);
    return (unsigned)(n > 148 * 64 ? 148 * 64 : n);
  };
  if (rows <= 16)
    reduce_kernel<1><<<nblocks(256), 256, 0, st>>>(part, out, groups, rows, cols, part_gs, row_stride, out_gs, out_dtype, accumulate);
  else if (rows <= 96)
    reduce_kernel<8><<<nblocks(32), 256, 0, st>>>(part, out, groups, rows, cols, part_gs, row_stride, out_gs, out_dtype, accumulate);
  else
    reduce_kernel<32><<<nblocks(8), 256, 0, st>>>(part, out, groups, rows, cols, part_gs, row_stride, out_gs, out_dtype, accumulate);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int bimamba_colsum_slices(int64_t rows) { return (int)((rows + kColRows - 1) / kColRows); }

extern "C" int bimamba_colsum(const void* x, float* part, int64_t rows, int cols, int64_t ld, int dtype,
                              bimamba_stream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  if (!x || !part) { set_err("colsum: null operand"); return -1; }
  if (rows < 0 || cols < 0 || dtype < 0 || dtype > 2) { set_err("colsum: bad sizes"); return -3; }
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)bimamba_colsum_slices(rows));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == BIMAMBA_F32) colsum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), part, rows, cols, ld);
  else if (dtype == BIMAMBA_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), part, rows, cols, ld);
  else colsum_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(x), part, rows, cols, ld);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
