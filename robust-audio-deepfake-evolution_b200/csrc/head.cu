// Backend head, forward (scoring): final LayerNorm -> attention pooling over time -> classifier, one kernel.  sm_100a.
//
// Reference: src/models/DualStreamSEMamba.py:759 (norm_f), :762-763 (softmax over T of attention_pool(f), weighted sum
// of the normalised frames), :767 (classifier); dropout (:764) is the identity in eval mode, which is where this
// kernel is used (produce_evaluation_file, src/main.py:958-995: score = logits[:, 1]).
//
// One CTA per utterance.  A warp owns frames t = warp, warp + 8, ...: the frame lives in registers (C <= 256), is
// normalised there (two-pass fp32), and enters a running (max, sum, weighted-frame) triple - the frames are read
// once and the (B, T, C) normalised tensor, the (B, T) logits and the softmax never touch memory.  The eight warps'
// triples are merged in fixed order through shared memory; the classifier is two warp reductions.
#include "common.cuh"

namespace bimamba {

constexpr int kHeadWarps = 8;
constexpr int kHeadThreads = kHeadWarps * 32;

__device__ __forceinline__ float warp_sum_h(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

template <typename T, int NPL>
__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ w_att, const float* __restrict__ b_att, const float* __restrict__ w_cls,
                const float* __restrict__ b_cls, float* __restrict__ features, float* __restrict__ logits, int L, int C,
                int ncls, float eps) {
  __shared__ float s_m[kHeadWarps], s_l[kHeadWarps];
  __shared__ float s_acc[kHeadWarps][32 * NPL];
  __shared__ float s_feat[32 * NPL];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const T* xb = x + (int64_t)b * L * C;
  float g[NPL], be[NPL], wa[NPL], acc[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < C ? __ldg(gamma + c) : 0.f;
    be[i] = c < C ? __ldg(beta + c) : 0.f;
    wa[i] = c < C ? __ldg(w_att + c) : 0.f;
    acc[i] = 0.f;
  }
  const float ba = b_att ? __ldg(b_att) : 0.f;
  float m = -INFINITY, l = 0.f;
  for (int t = warp; t < L; t += kHeadWarps) {
    const T* xr = xb + (int64_t)t * C;
    float v[NPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? to_f(xr[c]) : 0.f;
      s += v[i];
    }
    const float mu = warp_sum_h(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      const float d = c < C ? v[i] - mu : 0.f;
      q = fmaf(d, d, q);
    }
    const float rs = rsqrtf(warp_sum_h(q) / C + eps);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? fmaf((v[i] - mu) * rs, g[i], be[i]) : 0.f;   // norm_f(frame)
      dot = fmaf(wa[i], v[i], dot);
    }
    const float sc = warp_sum_h(dot) + ba;                          // attention logit of this frame
    const float mn = fmaxf(m, sc);
    const float scale = ex2_approx((m - mn) * kLog2e);              // 0 on the first frame (m = -inf)
    const float pw = ex2_approx((sc - mn) * kLog2e);
    l = fmaf(l, scale, pw);
#pragma unroll
    for (int i = 0; i < NPL; ++i) acc[i] = fmaf(acc[i], scale, pw * v[i]);
    m = mn;
  }
  if (lane == 0) {
    s_m[warp] = m;
    s_l[warp] = l;
  }
#pragma unroll
  for (int i = 0; i < NPL; ++i) s_acc[warp][lane + 32 * i] = acc[i];
  __syncthreads();
  // merge the warps in fixed order
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < kHeadWarps; ++w) M = fmaxf(M, s_m[w]);
  float den = 0.f;
#pragma unroll
  for (int w = 0; w < kHeadWarps; ++w) den += s_m[w] == -INFINITY ? 0.f : s_l[w] * ex2_approx((s_m[w] - M) * kLog2e);
  for (int c = threadIdx.x; c < C; c += kHeadThreads) {
    float num = 0.f;
#pragma unroll
    for (int w = 0; w < kHeadWarps; ++w)
      num += s_m[w] == -INFINITY ? 0.f : s_acc[w][c] * ex2_approx((s_m[w] - M) * kLog2e);
    const float f = num / den;
    s_feat[c] = f;
    features[(int64_t)b * C + c] = f;
  }
  __syncthreads();
  for (int j = warp; j < ncls; j += kHeadWarps) {
    float d = 0.f;
    for (int c = lane; c < C; c += 32) d = fmaf(__ldg(w_cls + (int64_t)j * C + c), s_feat[c], d);
    d = warp_sum_h(d);
    if (lane == 0) logits[(int64_t)b * ncls + j] = d + (b_cls ? __ldg(b_cls + j) : 0.f);
  }
}

template <typename T>
static void launch_head(const void* x, const float* gamma, const float* beta, const float* w_att, const float* b_att,
                        const float* w_cls, const float* b_cls, float* features, float* logits, int batch, int L, int C,
                        int ncls, float eps, cudaStream_t st) {
  const T* xp = reinterpret_cast<const T*>(x);
  if (C <= 160)
    head_fwd_kernel<T, 5><<<batch, kHeadThreads, 0, st>>>(xp, gamma, beta, w_att, b_att, w_cls, b_cls, features, logits, L, C, ncls, eps);
  else
    head_fwd_kernel<T, 8><<<batch, kHeadThreads, 0, st>>>(xp, gamma, beta, w_att, b_att, w_cls, b_cls, features, logits, L, C, ncls, eps);
}

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_head_fwd(const void* x, const float* gamma, const float* beta, const float* w_att,
                                const float* b_att, const float* w_cls, const float* b_cls, float* features,
                                float* logits, int batch, int seqlen, int channels, int nclasses, float eps, int dtype,
                                bimamba_stream_t stream) {
  if (batch == 0) return 0;
  if (!x || !gamma || !beta || !w_att || !w_cls || !features || !logits) { set_err("head: null operand"); return -1; }
  if (batch < 0 || seqlen < 1 || channels < 1 || channels > 256 || nclasses < 1) { set_err("head: seqlen >= 1, channels 1..256"); return -3; }
  if (dtype < 0 || dtype > 2) { set_err("head: bad dtype"); return -6; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == BIMAMBA_F32) launch_head<float>(x, gamma, beta, w_att, b_att, w_cls, b_cls, features, logits, batch, seqlen, channels, nclasses, eps, st);
  else if (dtype == BIMAMBA_BF16) launch_head<__nv_bfloat16>(x, gamma, beta, w_att, b_att, w_cls, b_cls, features, logits, batch, seqlen, channels, nclasses, eps, st);
  else launch_head<__half>(x, gamma, beta, w_att, b_att, w_cls, b_cls, features, logits, batch, seqlen, channels, nclasses, eps, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
