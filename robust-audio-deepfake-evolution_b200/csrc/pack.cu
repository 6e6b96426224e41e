// Per-step weight preparation for the Bi-Mamba block in ONE launch.  sm_100a.
//
// The parameters stay fp32 masters with the reference's names and shapes (mamba_block.py:22-39); every step
// the GEMM / scan kernels need them in the activation dtype and in a few derived arrangements:
//   Wi   (2D, dm)        in_proj.weight                        forward in_proj          (mamba_block.py:48)
//   WiT  (dm, 2D)        its transpose                         data gradient of in_proj
//   Wxp  (48, D)         x_proj.weight repacked [B | C | dt_r | 0]   forward x_proj     (mamba_block.py:73-75)
//   WxpT (D, 48)         its transpose                         data gradient of x_proj
//   Wo2  (dm, ndir*D)    [out_proj.weight | out_proj.weight]   forward out_proj over both directions (:62)
//   WoT  (D, dm)         out_proj.weight transposed            data gradient of out_proj
//   WdT  (16, D)         dt_proj.weight transposed, zero padded   data gradient of dt_proj
//   A    (D, N) fp32     -exp(A_log)                           (mamba_block.py:82)
// Done with torch ops this is ~15 tiny kernels per layer per step; here it is one.
#include "common.cuh"

namespace bimamba {

struct PackArgs {
  const float *W_in, *W_x, *W_dt, *A_log, *W_out;
  void *Wi, *WiT, *Wxp, *WxpT, *Wo2, *WoT, *WdT;
  float* A;
  int dm, D, N, R, ndir;
};

template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(const PackArgs a) {
  pdl_prologue();
  const int dm = a.dm, D = a.D, N = a.N, R = a.R;
  const int64_t n_wi = (int64_t)2 * D * dm, n_xp = (int64_t)kXW * D, n_wo2 = (int64_t)dm * a.ndir * D,
                n_wot = (int64_t)D * dm, n_wdt = (int64_t)BIMAMBA_MAX_DT_RANK * D, n_a = (int64_t)D * N;
  const int64_t total = 2 * n_wi + 2 * n_xp + n_wo2 + n_wot + n_wdt + n_a;
  T* Wi = reinterpret_cast<T*>(a.Wi);
  T* WiT = reinterpret_cast<T*>(a.WiT);
  T* Wxp = reinterpret_cast<T*>(a.Wxp);
  T* WxpT = reinterpret_cast<T*>(a.WxpT);
  T* Wo2 = reinterpret_cast<T*>(a.Wo2);
  T* WoT = reinterpret_cast<T*>(a.WoT);
  T* WdT = reinterpret_cast<T*>(a.WdT);
  auto xp = [&](int r, int c) -> float {  // row r of the repacked x_proj weight
    if (r < 2 * N) return __ldg(a.W_x + (int64_t)(R + r) * D + c);
    if (r < 2 * N + R) return __ldg(a.W_x + (int64_t)(r - 2 * N) * D + c);
    return 0.f;
  };
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t j = i;
    if (j < n_wi) { Wi[j] = from_f<T>(__ldg(a.W_in + j)); continue; }
    j -= n_wi;
    if (j < n_wi) {  // WiT[c][r] = W_in[r][c], c < dm, r < 2D
      const int c = (int)(j / (2 * D)), r = (int)(j % (2 * D));
      WiT[j] = from_f<T>(__ldg(a.W_in + (int64_t)r * dm + c));
      continue;
    }
    j -= n_wi;
    if (j < n_xp) { Wxp[j] = from_f<T>(xp((int)(j / D), (int)(j % D))); continue; }
    j -= n_xp;
    if (j < n_xp) { WxpT[j] = from_f<T>(xp((int)(j % kXW), (int)(j / kXW))); continue; }
    j -= n_xp;
    if (j < n_wo2) {
      const int m = (int)(j / (a.ndir * D)), c = (int)(j % (a.ndir * D)) % D;
      Wo2[j] = from_f<T>(__ldg(a.W_out + (int64_t)m * D + c));
      continue;
    }
    j -= n_wo2;
    if (j < n_wot) {  // WoT[c][m] = W_out[m][c]
      const int c = (int)(j / dm), m = (int)(j % dm);
      WoT[j] = from_f<T>(__ldg(a.W_out + (int64_t)m * D + c));
      continue;
    }
    j -= n_wot;
    if (j < n_wdt) {  // WdT[r][c] = W_dt[c][r] (zero beyond R)
      const int r = (int)(j / D), c = (int)(j % D);
      WdT[j] = from_f<T>(r < R ? __ldg(a.W_dt + (int64_t)c * R + r) : 0.f);
      continue;
    }
    j -= n_wdt;
    a.A[j] = -expf(__ldg(a.A_log + j));
  }
}

// dst (rows, cols) = cast(src), dstT (cols, rows) = cast(src)^T : a Linear weight and its data-gradient operand
template <typename T>
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ src, T* __restrict__ dst, T* __restrict__ dstT, int rows, int cols) {
  pdl_prologue();
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < 2 * n; i += (int64_t)gridDim.x * 256) {
    if (i < n) {
      dst[i] = from_f<T>(__ldg(src + i));
    } else {
      const int64_t j = i - n;
      const int c = (int)(j / rows), r = (int)(j % rows);
      dstT[j] = from_f<T>(__ldg(src + (int64_t)r * cols + c));
    }
  }
}

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_pack_weights(const float* W_in, const float* W_x, const float* W_dt, const float* A_log,
                                    const float* W_out, void* Wi, void* WiT, void* Wxp, void* WxpT, void* Wo2,
                                    void* WoT, void* WdT, float* A, int d_model, int d_inner, int d_state,
                                    int dt_rank, int ndir, int dtype, bimamba_stream_t stream) {
  if (!W_in || !W_x || !W_dt || !A_log || !W_out || !Wi || !WiT || !Wxp || !WxpT || !Wo2 || !WoT || !WdT || !A) {
    set_err("pack: null operand");
    return -1;
  }
  if (d_state != kN || dt_rank < 1 || dt_rank > BIMAMBA_MAX_DT_RANK || ndir < 1 || ndir > 2 || d_model < 1 || d_inner < 1 ||
      dtype < 0 || dtype > 2) {
    set_err("pack: bad sizes");
    return -3;
  }
  PackArgs a{W_in, W_x, W_dt, A_log, W_out, Wi, WiT, Wxp, WxpT, Wo2, WoT, WdT, A, d_model, d_inner, d_state, dt_rank, ndir};
  const int64_t total = (int64_t)4 * d_inner * d_model + 2 * (int64_t)kXW * d_inner + (int64_t)d_model * ndir * d_inner +
                        (int64_t)d_inner * d_model + (int64_t)BIMAMBA_MAX_DT_RANK * d_inner + (int64_t)d_inner * d_state;
  const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == BIMAMBA_F32) launch_k(pack_kernel<float>, blocks, 256, 0, st, a);
  else if (dtype == BIMAMBA_BF16) launch_k(pack_kernel<__nv_bfloat16>, blocks, 256, 0, st, a);
  else launch_k(pack_kernel<__half>, blocks, 256, 0, st, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int bimamba_cast_transpose(const float* src, void* dst, void* dstT, int rows, int cols, int dtype,
                                      bimamba_stream_t stream) {
  if (rows == 0 || cols == 0) return 0;
  if (!src || !dst || !dstT) { set_err("cast_transpose: null operand"); return -1; }
  if (rows < 0 || cols < 0 || dtype < 0 || dtype > 2) { set_err("cast_transpose: bad sizes"); return -3; }
  const int64_t total = 2 * (int64_t)rows * cols;
  const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == BIMAMBA_F32) launch_k(cast_transpose_kernel<float>, blocks, 256, 0, st, src, reinterpret_cast<float*>(dst), reinterpret_cast<float*>(dstT), rows, cols);
  else if (dtype == BIMAMBA_BF16) launch_k(cast_transpose_kernel<__nv_bfloat16>, blocks, 256, 0, st, src, reinterpret_cast<__nv_bfloat16*>(dst), reinterpret_cast<__nv_bfloat16*>(dstT), rows, cols);
  else launch_k(cast_transpose_kernel<__half>, blocks, 256, 0, st, src, reinterpret_cast<__half*>(dst), reinterpret_cast<__half*>(dstT), rows, cols);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
