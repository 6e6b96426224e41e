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

// ---- element IO: the arithmetic is always fp32; storage type is a runtime tag so one
// kernel body serves fp32 / bf16 / fp16 (the branch is warp-uniform and outside hot loops).
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

// softplus with the reference's threshold (torch.nn.functional.softplus: x if x > 20).
__device__ __forceinline__ float softplus_f(float v) { return v > 20.f ? v : log1pf(expf(v)); }

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ size_t dtype_size(int dt) { return dt == BIMAMBA_F32 ? 4 : 2; }

}  // namespace bimamba
