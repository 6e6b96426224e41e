// LayerNorm over the channel axis of (rows, C) activations, forward and backward.  sm_100a.
//
// Reference: the two nn.LayerNorm(d_model) of PN_BiMambas_Encoder (src/models/DualStreamSEMamba.py:
// 458-459, applied at :472 and :482) and norm_f (:703, :759).  One warp owns one row: the row lives
// in registers (C <= 1024), statistics are two-pass in fp32, and the output is written directly in
// the dtype the next GEMM consumes (no separate cast kernel).  The backward produces dx and per-CTA
// partial sums of dgamma / dbeta (rows are walked grid-stride; lanes keep their columns' partial sums
// in registers), reduced in fixed order by bimamba_reduce_partials.
#include "common.cuh"

namespace bimamba {

constexpr int kLnWarps = 8;
constexpr int kLnThreads = kLnWarps * 32;
constexpr int kLnMaxPerLane = 32;  // C <= 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

template <typename Tin, typename Tout, int NPL>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const Tin* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              Tout* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int C,
              float eps) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const Tin* xr = x + row * C;
  float v[NPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < C ? to_f(xr[c]) : 0.f;
    s += v[i];
  }
  const float mu = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    const float d = c < C ? v[i] - mu : 0.f;
    q = fmaf(d, d, q);
  }
  const float rs = rsqrtf(warp_sum(q) / C + eps);
  Tout* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) yr[c] = from_f<Tout>(fmaf((v[i] - mu) * rs, __ldg(gamma + c), __ldg(beta + c)));
  }
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
}

template <typename Tin, typename Tg, int NPL>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_kernel(const Tin* __restrict__ x, const Tg* __restrict__ dy, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, const Tin* __restrict__ addend,
              Tin* __restrict__ dx, float* __restrict__ part /* (gridDim.x, 2, C) */, int64_t rows, int C) {
  pdl_prologue();
  constexpr bool kWide = NPL > 8;     // wide rows: the warps add into one [2][C] buffer in turn (fixed order)
  __shared__ float sm[kWide ? 1 : kLnWarps][2][32 * NPL];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[NPL], dg[NPL], db[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < C ? __ldg(gamma + c) : 0.f;
    dg[i] = 0.f;
    db[i] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < rows; row += (int64_t)gridDim.x * kLnWarps) {
    const Tin* xr = x + row * C;
    const Tg* gr = dy + row * C;
    const float mu = mean[row], rs = rstd[row];
    float xh[NPL], dxh[NPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      float xv = 0.f, gv = 0.f;
      if (c < C) {
        xv = (to_f(xr[c]) - mu) * rs;
        gv = to_f(gr[c]);
      }
      xh[i] = xv;
      dxh[i] = gv * g[i];
      dg[i] = fmaf(gv, xv, dg[i]);
      db[i] += gv;
      s1 += dxh[i];
      s2 = fmaf(dxh[i], xv, s2);
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    Tin* dr = dx + row * C;
    const Tin* ar = addend ? addend + row * C : nullptr;    // gradient of the residual branch that bypasses the norm
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) dr[c] = from_f<Tin>(rs * (dxh[i] - s1 - xh[i] * s2) + (ar ? to_f(ar[c]) : 0.f));
    }
  }
  if constexpr (kWide) {
    for (int w = 0; w < kLnWarps; ++w) {
      if (warp == w) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          sm[0][0][lane + 32 * i] = w ? sm[0][0][lane + 32 * i] + dg[i] : dg[i];
          sm[0][1][lane + 32 * i] = w ? sm[0][1][lane + 32 * i] + db[i] : db[i];
        }
      }
      __syncthreads();
    }
    for (int e = threadIdx.x; e < 2 * C; e += kLnThreads) {
      const int k = e / C, c = e - k * C;
      part[((int64_t)blockIdx.x * 2 + k) * C + c] = sm[0][k][c];
    }
    return;
  } else {
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      sm[warp][0][lane + 32 * i] = dg[i];
      sm[warp][1][lane + 32 * i] = db[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * C; e += kLnThreads) {
      const int k = e / C, c = e - k * C;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kLnWarps; ++w) s += sm[w][k][c];
      part[((int64_t)blockIdx.x * 2 + k) * C + c] = s;
    }
  }
}

static int ln_blocks(int64_t rows) {
  int64_t b = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = 148 * 4;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

template <typename Tin, typename Tout>
static void launch_ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                          int64_t rows, int C, float eps, cudaStream_t st) {
  const unsigned blocks = (unsigned)((rows + kLnWarps - 1) / kLnWarps);
  const Tin* xp = reinterpret_cast<const Tin*>(x);
  Tout* yp = reinterpret_cast<Tout*>(y);
  if (C <= 160) launch_k(ln_fwd_kernel<Tin, Tout, 5>, blocks, kLnThreads, 0, st, xp, gamma, beta, yp, mean, rstd, rows, C, eps);
  else if (C <= 256) launch_k(ln_fwd_kernel<Tin, Tout, 8>, blocks, kLnThreads, 0, st, xp, gamma, beta, yp, mean, rstd, rows, C, eps);
  else launch_k(ln_fwd_kernel<Tin, Tout, kLnMaxPerLane>, blocks, kLnThreads, 0, st, xp, gamma, beta, yp, mean, rstd, rows, C, eps);
}

template <typename Tin, typename Tg>
static void launch_ln_bwd(const void* x, const void* dy, const float* gamma, const float* mean, const float* rstd,
                          const void* addend, void* dx, float* part, int64_t rows, int C, cudaStream_t st) {
  const unsigned blocks = (unsigned)ln_blocks(rows);
  const Tin* xp = reinterpret_cast<const Tin*>(x);
  const Tg* gp = reinterpret_cast<const Tg*>(dy);
  Tin* dxp = reinterpret_cast<Tin*>(dx);
  const Tin* ap = reinterpret_cast<const Tin*>(addend);
  if (C <= 160) launch_k(ln_bwd_kernel<Tin, Tg, 5>, blocks, kLnThreads, 0, st, xp, gp, gamma, mean, rstd, ap, dxp, part, rows, C);
  else if (C <= 256) launch_k(ln_bwd_kernel<Tin, Tg, 8>, blocks, kLnThreads, 0, st, xp, gp, gamma, mean, rstd, ap, dxp, part, rows, C);
  else launch_k(ln_bwd_kernel<Tin, Tg, kLnMaxPerLane>, blocks, kLnThreads, 0, st, xp, gp, gamma, mean, rstd, ap, dxp, part, rows, C);
}

}  // namespace bimamba

using namespace bimamba;

#define LN_DISPATCH2(A, B, CALL)                                                          \
  do {                                                                                     \
    if (A == BIMAMBA_F32 && B == BIMAMBA_F32) { using T1 = float; using T2 = float; CALL; }   \
    else if (A == BIMAMBA_F32 && B == BIMAMBA_BF16) { using T1 = float; using T2 = __nv_bfloat16; CALL; } \
    else if (A == BIMAMBA_F32 && B == BIMAMBA_F16) { using T1 = float; using T2 = __half; CALL; } \
    else if (A == BIMAMBA_BF16 && B == BIMAMBA_F32) { using T1 = __nv_bfloat16; using T2 = float; CALL; } \
    else if (A == BIMAMBA_BF16 && B == BIMAMBA_BF16) { using T1 = __nv_bfloat16; using T2 = __nv_bfloat16; CALL; } \
    else if (A == BIMAMBA_F16 && B == BIMAMBA_F32) { using T1 = __half; using T2 = float; CALL; } \
    else if (A == BIMAMBA_F16 && B == BIMAMBA_F16) { using T1 = __half; using T2 = __half; CALL; } \
    else { set_err("layernorm: unsupported dtype pair"); return -6; }                    \
  } while (0)

extern "C" int bimamba_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                     float* rstd, int64_t rows, int channels, float eps, int in_dtype, int out_dtype,
                                     bimamba_stream_t stream) {
  if (rows == 0) return 0;
  if (!x || !gamma || !beta || !y) { set_err("layernorm fwd: null operand"); return -1; }
  if (channels < 1 || channels > 32 * kLnMaxPerLane || rows < 0) { set_err("layernorm: channels must be 1..1024"); return -3; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LN_DISPATCH2(in_dtype, out_dtype, (launch_ln_fwd<T1, T2>(x, gamma, beta, y, mean, rstd, rows, channels, eps, st)));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int bimamba_layernorm_bwd_blocks(int64_t rows) { return ln_blocks(rows); }

extern "C" int bimamba_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* mean,
                                     const float* rstd, const void* dx_addend, void* dx, float* dgb_part, int64_t rows,
                                     int channels, int x_dtype, int dy_dtype, bimamba_stream_t stream) {
  if (rows == 0) return 0;
  if (!x || !dy || !gamma || !mean || !rstd || !dx || !dgb_part) { set_err("layernorm bwd: null operand"); return -1; }
  if (channels < 1 || channels > 32 * kLnMaxPerLane || rows < 0) { set_err("layernorm backward: channels must be 1..1024"); return -3; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LN_DISPATCH2(x_dtype, dy_dtype, (launch_ln_bwd<T1, T2>(x, dy, gamma, mean, rstd, dx_addend, dx, dgb_part, rows, channels, st)));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
