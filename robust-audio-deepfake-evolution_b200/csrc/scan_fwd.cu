// Selective scan, forward, both time directions in one launch.  sm_100a.
//
// Math (reference: src/models/modules/mamba_block.py:80-120 and :61):
//   delta = softplus(Wdt . dtr[t] + bias)            (or softplus(delta_raw + bias) when given)
//   a[t,n] = exp(delta[t] * A[n]);  h[t,n] = a[t,n] h[t-1,n] + delta[t] B[t,n] u[t]
//   y[t] = sum_n C[t,n] h[t,n] + D u[t];  out[t] = y[t] * silu(z[t])
//
// Mapping (B200-first; not the upstream block-scan):
//   * activations are channel-last (batch, dir, time, channel).  One THREAD owns one channel of
//     one (batch, direction): its 16 states live in registers (fp32, packed as 8 float2 so the
//     recurrence issues as FMUL2 / FFMA2 - Blackwell's packed fp32 pipe) for the whole
//     sequence, so the n-sum needs no shuffles and the per-element transcendentals (softplus,
//     SiLU) are computed exactly once.  A warp reads 32 consecutive channels of a time row.
//   * time is walked in chunks of 16 steps.  The chunk's tiles - u, z (and delta when given)
//     [16 x G channels] and the per-(batch,time) projection rows [16 x (B|C|dt_r)], which are
//     shared by every channel - are staged by 16-byte cp.async into double-buffered shared
//     memory one chunk ahead of the compute; B/C/dt_r are then broadcast-read as float4.
//   * the dt projection (K = dt_rank = 9: too thin for tensor cores) is 9 FMAs per element
//     against the staged dt_r row, so delta is never written to or read from HBM.
//   * direction 1 walks the same storage back to front (t = L-1-step): flip(M(flip(x))) of
//     src/models/DualStreamSEMamba.py:476-478 with no flipped copy; both directions are
//     blockIdx.y of the same launch.
//   * training forward also writes the fp32 state entering every chunk ("checkpoints",
//     (B, dir, chunk, D, 16): 64 contiguous bytes per thread) and the pre-gate y; the backward
//     recomputes the states of a chunk from its checkpoint (no (B, L, D, N) tensor).
#include "common.cuh"

namespace bimamba {

constexpr int kFwdMaxThreads = 128;

template <typename T>
__global__ void __launch_bounds__(kFwdMaxThreads) scan_fwd_kernel(const bimamba_scan_desc p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int G = blockDim.x, tid = threadIdx.x;
  const int b = blockIdx.z, dir = blockIdx.y, d0 = blockIdx.x * G, d = d0 + tid;
  const bool ok = d < p.dim;
  const int L = p.seqlen, nck = (L + kT - 1) / kT;
  const bool has_z = p.z != nullptr, expl = p.delta != nullptr;
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;
  const int R = expl ? 0 : p.dt_rank;

  const T* gu = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds;
  const T* gz = has_z ? reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds : nullptr;
  const T* gd = expl ? reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds : nullptr;
  const T* gbc = reinterpret_cast<const T*>(p.bc) + (int64_t)b * p.bc_bs + (int64_t)dir * p.bc_ds;
  const T* gdtr = R ? reinterpret_cast<const T*>(p.dtr) + (int64_t)b * p.dtr_bs + (int64_t)dir * p.dtr_ds : nullptr;
  const int64_t obase = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds;
  T* gout = reinterpret_cast<T*>(p.out) + obase;
  T* gyp = p.ypre ? reinterpret_cast<T*>(p.ypre) + obase : nullptr;

  // shared memory: [2][nact][kT*G] T | [2][kT*kXW] T | [kT*kXW] float
  const int nact = 1 + (has_z ? 1 : 0) + (expl ? 1 : 0);
  T* s_act = reinterpret_cast<T*>(smem_raw);
  T* s_xr = s_act + 2 * nact * kT * G;
  float* s_xf = reinterpret_cast<float*>(s_xr + 2 * kT * kXW);

  constexpr int kV = 16 / sizeof(T);
  const bool dim_vec = (p.dim % kV) == 0;
  const bool vec_u = dim_vec && aligned16(gu + d0) && (p.u_ts % kV) == 0;
  const bool vec_z = has_z && dim_vec && aligned16(gz + d0) && (p.z_ts % kV) == 0;
  const bool vec_d = expl && dim_vec && aligned16(gd + d0) && (p.delta_ts % kV) == 0;
  const bool vec_bc = aligned16(gbc) && (p.bc_ts % kV) == 0;
  const bool vec_dtr = R && (p.flags & BIMAMBA_FLAG_DTR_PADDED) && aligned16(gdtr) && (p.dtr_ts % kV) == 0;

  auto stage = [&](int c0, int bf) {
    auto row_of = [&](int i) -> int64_t {
      const int tau = c0 * kT + i;
      return tau < L ? (int64_t)(dir ? (L - 1 - tau) : tau) : (int64_t)-1;
    };
    T* sa = s_act + bf * nact * kT * G;
    stage_tile(sa, G, gu, p.u_ts, kT, G, d0, p.dim, vec_u, row_of, tid, G);
    int k = 1;
    if (has_z) stage_tile(sa + (k++) * kT * G, G, gz, p.z_ts, kT, G, d0, p.dim, vec_z, row_of, tid, G);
    if (expl) stage_tile(sa + k * kT * G, G, gd, p.delta_ts, kT, G, d0, p.dim, vec_d, row_of, tid, G);
    T* sx = s_xr + bf * kT * kXW;
    stage_tile(sx, kXW, gbc, p.bc_ts, kT, 2 * kN, 0, 2 * kN, vec_bc, row_of, tid, G);
    if (R) {
      const int w = vec_dtr ? 16 : R;
      stage_tile(sx + 2 * kN, kXW, gdtr, p.dtr_ts, kT, w, 0, w, vec_dtr, row_of, tid, G);
    }
    cp_async_commit();
  };

  // per-channel constants
  float2 A2[kN / 2], h[kN / 2];
  float4 wdt[4];
  float bias = 0.f, Dd = 0.f;
#pragma unroll
  for (int j = 0; j < kN / 2; ++j) {
    h[j] = make_float2(0.f, 0.f);
    A2[j] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) wdt[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
#pragma unroll
    for (int j = 0; j < kN / 2; ++j) {
      A2[j].x = __ldg(p.A + (int64_t)d * kN + 2 * j) * kLog2e;
      A2[j].y = __ldg(p.A + (int64_t)d * kN + 2 * j + 1) * kLog2e;
    }
    if (p.delta_bias) bias = __ldg(p.delta_bias + d);
    if (p.D) Dd = __ldg(p.D + d);
    if (R) {
      float* w = reinterpret_cast<float*>(wdt);
#pragma unroll
      for (int r = 0; r < BIMAMBA_MAX_DT_RANK; ++r)
        if (r < R) w[r] = __ldg(p.Wdt + (int64_t)d * R + r);
    }
  }
  const int R4 = (R + 3) >> 2;

  if (nck > 0) stage(0, 0);
  for (int c0 = 0; c0 < nck; ++c0) {
    const int bf = c0 & 1;
    cp_async_wait<0>();
    __syncthreads();  // chunk c0 is visible; every thread is done with chunk c0-1's buffers
    if (c0 + 1 < nck) stage(c0 + 1, bf ^ 1);
    {
      const T* sx = s_xr + bf * kT * kXW;
      const int valid = 2 * kN + R;
      for (int e = tid; e < kT * kXW; e += G) {
        const int col = e % kXW;
        s_xf[e] = col < valid ? to_f(sx[e]) : 0.f;
      }
    }
    __syncthreads();
    if (ok) {
      if (p.ckpt) {
        float4* ck = reinterpret_cast<float4*>(
            p.ckpt + ((((int64_t)b * p.ndir + dir) * nck + c0) * p.dim + d) * kN);
#pragma unroll
        for (int q = 0; q < 4; ++q) ck[q] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
      }
      const T* su = s_act + bf * nact * kT * G + tid;
      const T* sz = su + kT * G;
      const T* sd = su + (has_z ? 2 : 1) * kT * G;
      const int nsteps = min(kT, L - c0 * kT);
#pragma unroll 4
      for (int i = 0; i < nsteps; ++i) {
        const float4* xr = reinterpret_cast<const float4*>(s_xf + i * kXW);
        const float u = to_f(su[i * G]);
        float draw = bias;
        if (expl) {
          draw += to_f(sd[i * G]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < R4) {
              const float4 x = xr[8 + q];
              draw = fmaf(wdt[q].x, x.x, draw);
              draw = fmaf(wdt[q].y, x.y, draw);
              draw = fmaf(wdt[q].z, x.z, draw);
              draw = fmaf(wdt[q].w, x.w, draw);
            }
          }
        }
        const float delta = softplus ? softplus_f(draw) : draw;
        const float du = delta * u;
        const float2 dd = make_float2(delta, delta), duu = make_float2(du, du);
        float2 y2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 Bq = xr[q], Cq = xr[4 + q];
          {
            const float2 x = __fmul2_rn(dd, A2[2 * q]);
            const float2 a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            h[2 * q] = __ffma2_rn(a, h[2 * q], __fmul2_rn(duu, make_float2(Bq.x, Bq.y)));
            y2 = __ffma2_rn(make_float2(Cq.x, Cq.y), h[2 * q], y2);
          }
          {
            const float2 x = __fmul2_rn(dd, A2[2 * q + 1]);
            const float2 a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            h[2 * q + 1] = __ffma2_rn(a, h[2 * q + 1], __fmul2_rn(duu, make_float2(Bq.z, Bq.w)));
            y2 = __ffma2_rn(make_float2(Cq.z, Cq.w), h[2 * q + 1], y2);
          }
        }
        float y = fmaf(Dd, u, y2.x + y2.y);
        const int tau = c0 * kT + i;
        const int64_t o = (int64_t)(dir ? (L - 1 - tau) : tau) * p.out_ts + d;
        if (gyp) gyp[o] = from_f<T>(y);
        if (has_z) {
          const float z = to_f(sz[i * G]);
          y *= z * sigmoid_f(z);
        }
        gout[o] = from_f<T>(y);
      }
    }
  }
}

static size_t fwd_smem_bytes(int G, int esize, bool has_z, bool expl) {
  const int nact = 1 + (has_z ? 1 : 0) + (expl ? 1 : 0);
  return (size_t)2 * nact * kT * G * esize + (size_t)2 * kT * kXW * esize + (size_t)kT * kXW * 4;
}

template <typename T>
static int launch_fwd(const bimamba_scan_desc* d, cudaStream_t st) {
  const int G = d->group_channels;
  const size_t smem = fwd_smem_bytes(G, (int)sizeof(T), d->z != nullptr, d->delta != nullptr);
  dim3 grid((d->dim + G - 1) / G, d->ndir, d->batch);
  if (smem > 48 * 1024) cudaFuncSetAttribute(scan_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  scan_fwd_kernel<T><<<grid, G, smem, st>>>(*d);
  return 0;
}

int check_desc(const bimamba_scan_desc* d, bool bwd);  // api.cu

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_selective_scan_fwd(const bimamba_scan_desc* d, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;  // empty: nothing to do (pointers may be null)
  int rc = check_desc(d, false);
  if (rc) return rc;
  const int G = d->group_channels;
  if (G < 32 || G > kFwdMaxThreads || (G & 31)) { set_err("forward group_channels must be 32, 64, 96 or 128 (use bimamba_scan_plan)"); return -5; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d->io_dtype) {
    case BIMAMBA_F32: launch_fwd<float>(d, st); break;
    case BIMAMBA_BF16: launch_fwd<__nv_bfloat16>(d, st); break;
    default: launch_fwd<__half>(d, st); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
