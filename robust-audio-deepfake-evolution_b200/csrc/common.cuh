// Shared device helpers for the sm_100a Bi-Mamba kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bimamba.h"

namespace bimamba {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kT = BIMAMBA_CHUNK;   // steps per chunk == checkpoint interval
constexpr int kN = BIMAMBA_DSTATE;  // d_state handled by these kernels
constexpr int kXW = 48;             // staged row of projection features: B(16) | C(16) | dt_r(<=16)

void set_err(const char* msg);  // api.cu
extern int g_tune[BIMAMBA_TUNE_COUNT];  // api.cu: bimamba_set_tuning

// ---- element conversion (arithmetic is always fp32; storage type is the template parameter)
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// runtime-tagged access (cold paths only)
__device__ __forceinline__ float ld_f(const void* __restrict__ p, int64_t i, int dt) {
  if (dt == BIMAMBA_F32) return __ldg(reinterpret_cast<const float*>(p) + i);
  unsigned short raw = __ldg(reinterpret_cast<const unsigned short*>(p) + i);
  if (dt == BIMAMBA_BF16) return __uint_as_float(static_cast<unsigned>(raw) << 16);
  return __half2float(__ushort_as_half(raw));
}
__device__ __forceinline__ void st_f(void* __restrict__ p, int64_t i, float v, int dt) {
  if (dt == BIMAMBA_F32) {
    reinterpret_cast<float*>(p)[i] = v;
  } else if (dt == BIMAMBA_BF16) {
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  } else {
    reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// softplus with the reference's threshold (torch.nn.functional.softplus: x if x > 20).
// log1p(e) for small e is taken as e - e^2/2 + e^3/3 (lg2(1+e) would lose e's low bits).
// Branch-free (lanes of a warp mix both ranges).
__device__ __forceinline__ float softplus_f(float v) {
  const float e = ex2_approx(fminf(v, 20.f) * kLog2e);
  const float small = e * fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.33333334f), -0.5f), 1.f);
  const float big = lg2_approx(1.f + e) * kLn2;
  const float r = e < 0.03125f ? small : big;
  return v > 20.f ? v : r;
}

__device__ __forceinline__ float sigmoid_f(float v) { return rcp_approx(1.f + ex2_approx(-v * kLog2e)); }

// ---- cp.async (LDGSTS) helpers: 16-byte copies global -> shared, zero-filled when !pred
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const int n = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Stage a [rows x width] tile of T into shared memory (row stride `sw` elements in smem).
// Row i of the tile is global row `row_of(i)` (or invalid if row_of(i) < 0 -> zero-filled);
// columns [col0, col0+width) of which only those < col_end are valid.  `vec` selects the
// 16-byte cp.async path (caller checked alignment), else a scalar synchronous copy.
template <typename T, typename RowFn>
__device__ __forceinline__ void stage_tile(T* __restrict__ sm, int sw, const T* __restrict__ g, int64_t g_ts,
                                           int rows, int width, int col0, int col_end, bool vec, RowFn row_of,
                                           int tid, int nthreads) {
  if (vec) {
    constexpr int kV = 16 / sizeof(T);
    const int vpr = width / kV;  // vectors per row
    for (int e = tid; e < rows * vpr; e += nthreads) {
      const int i = e / vpr, v = e - i * vpr;
      const int64_t r = row_of(i);
      const int c = col0 + v * kV;
      const bool ok = r >= 0 && c < col_end;
      cp_async16(sm + i * sw + v * kV, ok ? (g + r * g_ts + c) : g, ok);
    }
  } else {
    for (int e = tid; e < rows * width; e += nthreads) {
      const int i = e / width, v = e - i * width;
      const int64_t r = row_of(i);
      const int c = col0 + v;
      sm[i * sw + v] = (r >= 0 && c < col_end) ? g[r * g_ts + c] : from_f<T>(0.f);
    }
  }
}

// ---- programmatic dependent launch (PDL).  Every kernel of this library is launched with the programmatic-stream-
// serialization attribute and begins with pdl_prologue() = griddepcontrol.wait: the grid may be scheduled as soon as the
// previous kernel's CTAs have exited, and waits there until that kernel's memory is visible, instead of paying a full
// launch after its completion - inside a captured graph the edges between this library's kernels become programmatic.
// Measured (config 2, graph replay): 2.401 -> 2.317 ms per step.  An EARLY trigger (griddepcontrol.launch_dependents at
// the top of every kernel) was measured slower (2.475 ms): the pre-scheduled dependents sit on registers / shared memory
// the running kernel could use; it stays available behind BIMAMBA_PDL_EARLY_TRIGGER.  No-ops without the attribute.
__device__ __forceinline__ void pdl_prologue() {
#ifdef BIMAMBA_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Launch with the programmatic-stream-serialization attribute (BIMAMBA_TUNE_PDL = 1 switches it off for A/B runs).
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = g_tune[BIMAMBA_TUNE_PDL] == 1 ? 0 : 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace bimamba
