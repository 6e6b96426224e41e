// Selective scan, forward - the TWO-LANES-PER-CHANNEL variant used for small problems (see scan_fwd.cu for the
// math, the layout and the one-lane variant).  sm_100a.
//
// Math (reference: src/models/modules/mamba_block.py:80-120 and :61):
//   delta = softplus(Wdt . dtr[t] + bias)            (or softplus(delta_raw + bias) when given)
//   a[t,n] = exp(delta[t] * A[n]);  h[t,n] = a[t,n] h[t-1,n] + delta[t] B[t,n] u[t]
//   y[t] = sum_n C[t,n] h[t,n] + D u[t];  out[t] = y[t] * silu(z[t])
//
// Mapping (B200-first; not the upstream block-scan):
//   * activations are channel-last (batch, dir, time, channel).  A PAIR of lanes owns one channel of
//     one (batch, direction): each lane keeps 8 of its 16 states in registers (fp32, packed as 4
//     float2 so the recurrence issues as FMUL2 / FFMA2 - Blackwell's packed fp32 pipe) for the
//     whole sequence.  Steps are taken two at a time: lane s does the per-element work (dt
//     projection, softplus, gate) of step 2j+s, the pair exchanges delta / delta*u / partial y by
//     shuffle - so the transcendentals are still computed exactly once per element, the two lanes
//     run the same instruction stream, and there are twice as many warps to hide latency with
//     (the Phase-6 shapes only have 36 864 (batch, dir, channel) rows for 148 SMs).
//   * time is walked in chunks of 16 steps.  The chunk's tiles - u, z (and delta when given)
//     [16 x G channels] and the per-(batch,time) projection rows [16 x (B|C|dt_r)], which are
//     shared by every channel - are staged by 16-byte cp.async into double-buffered shared
//     memory one chunk ahead of the compute; B/C/dt_r are then broadcast-read as float4.
//   * the dt projection (K = dt_rank = 9: too thin for tensor cores) is 9 FMAs per element
//     against the staged dt_r row, so delta is never written to or read from HBM.
//   * direction 1 walks the same storage back to front (t = L-1-step): flip(M(flip(x))) of
//     src/models/DualStreamSEMamba.py:476-478 with no flipped copy; both directions are
//     blockIdx.y of the same launch.
//   * training forward also writes the fp32 state entering every 8-step chunk ("checkpoints",
//     (B, dir, chunk, D, 16): 64 contiguous bytes per thread) and the pre-gate y; the backward
//     recomputes the states of a chunk from its checkpoint (no (B, L, D, N) tensor).
#include "common.cuh"

namespace bimamba {

constexpr int kFwdPairMaxThreads = 256;

// kMode: 0 = delta given; 1 = fused dt projection with dt_rank <= 12; 2 = dt_rank <= 16.
template <typename T, int kMode, bool kGate>
__global__ void __launch_bounds__(kFwdPairMaxThreads) scan_fwd_pair_kernel(const bimamba_scan_desc p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool expl = kMode == 0;
  constexpr int R4 = kMode == 1 ? 3 : 4;
  const int NT = blockDim.x, G = NT >> 1, tid = threadIdx.x;
  const int cidx = tid >> 1, sh = tid & 1;  // channel within the CTA, state half (n = 8*sh .. 8*sh+7)
  const int b = blockIdx.z, dir = blockIdx.y, d0 = blockIdx.x * G, d = d0 + cidx;
  const bool ok = d < p.dim;
  const int L = p.seqlen, nck = (L + kT - 1) / kT, nckpt = (L + BIMAMBA_CKPT - 1) / BIMAMBA_CKPT;
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;
  const int R = expl ? 0 : p.dt_rank;

  const T* gu = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds;
  const T* gz = kGate ? reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds : nullptr;
  const T* gd = expl ? reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds : nullptr;
  const T* gbc = reinterpret_cast<const T*>(p.bc) + (int64_t)b * p.bc_bs + (int64_t)dir * p.bc_ds;
  const T* gdtr = expl ? nullptr : reinterpret_cast<const T*>(p.dtr) + (int64_t)b * p.dtr_bs + (int64_t)dir * p.dtr_ds;
  const int64_t obase = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds;

  // shared memory: [2][nact][kT*G] T | [2][kT*kXW] T | [kT*kXW] float
  constexpr int nact = 1 + (kGate ? 1 : 0) + (expl ? 1 : 0);
  T* s_act = reinterpret_cast<T*>(smem_raw);
  T* s_xr = s_act + 2 * nact * kT * G;
  float* s_xf = reinterpret_cast<float*>(s_xr + 2 * kT * kXW);

  constexpr int kV = 16 / sizeof(T);
  const bool dim_vec = (p.dim % kV) == 0;
  const bool vec_u = dim_vec && aligned16(gu + d0) && (p.u_ts % kV) == 0;
  const bool vec_z = kGate && dim_vec && aligned16(gz + d0) && (p.z_ts % kV) == 0;
  const bool vec_d = expl && dim_vec && aligned16(gd + d0) && (p.delta_ts % kV) == 0;
  const bool vec_bc = aligned16(gbc) && (p.bc_ts % kV) == 0;
  const bool vec_dtr = !expl && (p.flags & BIMAMBA_FLAG_DTR_PADDED) && aligned16(gdtr) && (p.dtr_ts % kV) == 0;

  auto stage = [&](int c0, int bf) {
    auto row_of = [&](int i) -> int64_t {
      const int tau = c0 * kT + i;
      return tau < L ? (int64_t)(dir ? (L - 1 - tau) : tau) : (int64_t)-1;
    };
    T* sa = s_act + bf * nact * kT * G;
    stage_tile(sa, G, gu, p.u_ts, kT, G, d0, p.dim, vec_u, row_of, tid, NT);
    if (kGate) stage_tile(sa + kT * G, G, gz, p.z_ts, kT, G, d0, p.dim, vec_z, row_of, tid, NT);
    if (expl) stage_tile(sa + (nact - 1) * kT * G, G, gd, p.delta_ts, kT, G, d0, p.dim, vec_d, row_of, tid, NT);
    T* sx = s_xr + bf * kT * kXW;
    stage_tile(sx, kXW, gbc, p.bc_ts, kT, 2 * kN, 0, 2 * kN, vec_bc, row_of, tid, NT);
    if (!expl) {
      const int w = vec_dtr ? 16 : R;
      stage_tile(sx + 2 * kN, kXW, gdtr, p.dtr_ts, kT, w, 0, w, vec_dtr, row_of, tid, NT);
    }
    cp_async_commit();
  };

  // per-thread constants: this lane's 8 states of channel d
  float2 A2[kN / 4], h[kN / 4];
  float2 wdt[2 * R4];
  float bias = 0.f, Dd = 0.f;
#pragma unroll
  for (int j = 0; j < kN / 4; ++j) {
    h[j] = make_float2(0.f, 0.f);
    A2[j] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < 2 * R4; ++q) wdt[q] = make_float2(0.f, 0.f);
  if (ok) {
#pragma unroll
    for (int j = 0; j < kN / 4; ++j) {
      A2[j].x = __ldg(p.A + (int64_t)d * kN + 8 * sh + 2 * j) * kLog2e;
      A2[j].y = __ldg(p.A + (int64_t)d * kN + 8 * sh + 2 * j + 1) * kLog2e;
    }
    if (p.delta_bias) bias = __ldg(p.delta_bias + d);
    if (p.D) Dd = __ldg(p.D + d);
    if (!expl) {
      float* w = reinterpret_cast<float*>(wdt);
#pragma unroll
      for (int r = 0; r < 4 * R4; ++r)
        if (r < R) w[r] = __ldg(p.Wdt + (int64_t)d * R + r);
    }
  }
  // output walks: this lane finishes steps sh, sh+2, sh+4, ... (element (t, d), t advancing by +-2 per pair)
  const int64_t ostep2 = 2 * (dir ? -p.out_ts : p.out_ts);
  int64_t opos = obase + (int64_t)(dir ? (L - 1 - sh) : sh) * p.out_ts + d;
  T* const gout = reinterpret_cast<T*>(p.out);
  T* const gyp = reinterpret_cast<T*>(p.ypre);
  const unsigned pair_src0 = (unsigned)(tid & 31 & ~1);  // lane of the pair that owns even steps

  if (nck > 0) stage(0, 0);
  for (int c0 = 0; c0 < nck; ++c0) {
    const int bf = c0 & 1;
    cp_async_wait<0>();
    __syncthreads();  // chunk c0 is visible; every thread is done with chunk c0-1's buffers
    if (c0 + 1 < nck) stage(c0 + 1, bf ^ 1);
    {
      const T* sx = s_xr + bf * kT * kXW;
      const int valid = 2 * kN + R;
      for (int e = tid; e < kT * kXW; e += NT) {
        const int col = e % kXW;
        s_xf[e] = col < valid ? to_f(sx[e]) : 0.f;
      }
    }
    __syncthreads();
    // Both lanes of a channel walk the chunk together.  Per pair of steps (2j, 2j+1) lane `sh` does the
    // per-ELEMENT work (dt projection, softplus, gate) of step 2j+sh, the pair exchanges delta and delta*u by
    // shuffle, each lane advances ITS 8 states through both steps, and the partial y of the other lane's step
    // is shuffled back: the two lanes execute identical instruction streams (no divergence, nothing duplicated).
    const T* su = s_act + bf * nact * kT * G + cidx;
    const int nvalid = L - c0 * kT;  // steps of this chunk that exist (>= 1)
#pragma unroll 2
    for (int jp = 0; jp < kT / 2; ++jp) {
      const int io = 2 * jp + sh;    // the step whose element this lane owns
      if ((jp & (BIMAMBA_CKPT / 2 - 1)) == 0 && p.ckpt && ok && 2 * jp < nvalid) {  // state entering this 8-step chunk
        float4* ck = reinterpret_cast<float4*>(
            p.ckpt + ((((int64_t)b * p.ndir + dir) * nckpt + (c0 * kT + 2 * jp) / BIMAMBA_CKPT) * p.dim + d) * kN + 8 * sh);
        ck[0] = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
        ck[1] = make_float4(h[2].x, h[2].y, h[3].x, h[3].y);
      }
      const float u = to_f(su[io * G]);
      float draw;
      if (expl) {
        draw = bias + to_f(su[((nact - 1) * kT + io) * G]);
      } else {
        const float4* xo = reinterpret_cast<const float4*>(s_xf + io * kXW + 2 * kN);
        float2 acc0 = make_float2(bias, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < R4; ++q) {
          const float4 x = xo[q];
          acc0 = __ffma2_rn(wdt[2 * q], make_float2(x.x, x.y), acc0);
          acc1 = __ffma2_rn(wdt[2 * q + 1], make_float2(x.z, x.w), acc1);
        }
        draw = (acc0.x + acc0.y) + (acc1.x + acc1.y);
      }
      const float delta_o = softplus ? softplus_f(draw) : draw;
      const float du_o = delta_o * u;
      float ypart[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float delta = __shfl_sync(kFull, delta_o, pair_src0 + e);
        const float du = __shfl_sync(kFull, du_o, pair_src0 + e);
        const float4* xr = reinterpret_cast<const float4*>(s_xf + (2 * jp + e) * kXW + 8 * sh);
        const float4 B0 = xr[0], B1 = xr[1], C0 = xr[4], C1 = xr[5];
        const float2 dd = make_float2(delta, delta), duu = make_float2(du, du);
        const float2 Bv[4] = {make_float2(B0.x, B0.y), make_float2(B0.z, B0.w), make_float2(B1.x, B1.y), make_float2(B1.z, B1.w)};
        const float2 Cv[4] = {make_float2(C0.x, C0.y), make_float2(C0.z, C0.w), make_float2(C1.x, C1.y), make_float2(C1.z, C1.w)};
        float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = __fmul2_rn(dd, A2[j]);
          const float2 a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          h[j] = __ffma2_rn(a, h[j], __fmul2_rn(duu, Bv[j]));
          if (j & 1) yb = __ffma2_rn(Cv[j], h[j], yb);
          else ya = __ffma2_rn(Cv[j], h[j], ya);
        }
        ypart[e] = (ya.x + ya.y) + (yb.x + yb.y);
      }
      // the other lane's partial of MY step comes back; mine of ITS step goes over
      const float mine = sh ? ypart[1] : ypart[0];
      const float send = sh ? ypart[0] : ypart[1];
      float y = fmaf(Dd, u, mine + __shfl_xor_sync(kFull, send, 1));
      if (ok && io < nvalid) {
        if (gyp) gyp[opos] = from_f<T>(y);
        if (kGate) {
          const float z = to_f(su[(kT + io) * G]);
          y *= z * sigmoid_f(z);
        }
        gout[opos] = from_f<T>(y);
      }
      opos += ostep2;
    }
  }
}

static size_t fwd_pair_smem_bytes(int G, int esize, bool has_z, bool expl) {
  const int nact = 1 + (has_z ? 1 : 0) + (expl ? 1 : 0);
  return (size_t)2 * nact * kT * G * esize + (size_t)2 * kT * kXW * esize + (size_t)kT * kXW * 4;
}

template <typename T, int kMode, bool kGate>
static void launch_fwd_pair2(const bimamba_scan_desc* d, cudaStream_t st) {
  const int G = d->group_channels;
  const size_t smem = fwd_pair_smem_bytes(G, (int)sizeof(T), kGate, kMode == 0);
  dim3 grid((d->dim + G - 1) / G, d->ndir, d->batch);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(scan_fwd_pair_kernel<T, kMode, kGate>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  scan_fwd_pair_kernel<T, kMode, kGate><<<grid, 2 * G, smem, st>>>(*d);
}

template <typename T>
static void launch_fwd_pair_t(const bimamba_scan_desc* d, cudaStream_t st) {
  const bool gate = d->z != nullptr;
  const int mode = d->delta ? 0 : (d->dt_rank <= 12 ? 1 : 2);
  if (gate) {
    if (mode == 0) launch_fwd_pair2<T, 0, true>(d, st);
    else if (mode == 1) launch_fwd_pair2<T, 1, true>(d, st);
    else launch_fwd_pair2<T, 2, true>(d, st);
  } else {
    if (mode == 0) launch_fwd_pair2<T, 0, false>(d, st);
    else if (mode == 1) launch_fwd_pair2<T, 1, false>(d, st);
    else launch_fwd_pair2<T, 2, false>(d, st);
  }
}


// Called by bimamba_selective_scan_fwd (scan_fwd.cu) when the problem has too few (batch, dir, channel) rows to
// fill the GPU with one lane per channel.
void launch_fwd_pair(const bimamba_scan_desc* d, cudaStream_t st) {
  switch (d->io_dtype) {
    case BIMAMBA_F32: launch_fwd_pair_t<float>(d, st); break;
    case BIMAMBA_BF16: launch_fwd_pair_t<__nv_bfloat16>(d, st); break;
    default: launch_fwd_pair_t<__half>(d, st); break;
  }
}

}  // namespace bimamba
