// Selective scan, forward, TIME-PARALLEL variant of the one-warp-per-CTA kernel of scan_fwd1.cu (same staging, same
// per-step math; reference: src/models/modules/mamba_block.py:80-120, :61).  sm_100a.
// Kept in its own translation unit so that the unsplit kernel of scan_fwd1.cu compiles exactly as it did before.
//
// TIME SPLIT (kSeg != 0; bimamba_selective_scan_fwd_split).  A long sequence at a small batch leaves most of the GPU
// idle (8192 steps x batch 64 = 3.9 warps of channel lanes per SM), so scan time is cut into nseg segments of seg_len
// steps (a multiple of the 16-step chunk) that run as separate CTAs, with the recurrence's carry between them:
//   kSeg = 1, carry pass (segments 0 .. nseg-2): the segment's recurrence from a zero state WITHOUT outputs (no C.h, no
//             gate, no stores) -> its end state e_s[16] and sum of step sizes S_s per channel;
//   kSeg = 2, output pass (all segments): state entering segment s by the associative rule of the scan applied to the
//             carries in order, h <- exp(A S_k) h + e_k for k < s  (prod_t exp(delta_t A) = exp(A sum_t delta_t)), then the
//             ordinary loop over the segment's chunks (outputs, ypre and checkpoints are indexed by absolute scan time,
//             so the backward needs no change).
#include "common.cuh"

namespace bimamba {

constexpr int kF1G = 32;

struct SegArgs {
  float* hend;   // (batch, ndir, nseg - 1, dim, 16) end state of segments 0 .. nseg-2 (zero initial state)
  float* sdel;   // (batch, ndir, nseg - 1, dim)     sum of delta over the segment
  int nseg, seg_len;
};

template <typename T, int kMode, bool kGate, int kSeg>
__global__ void __launch_bounds__(kF1G) scan_fwd_warp_seg_kernel(const bimamba_scan_desc p, const SegArgs sa_) {
  static_assert(kSeg == 1 || kSeg == 2, "1 = carry pass, 2 = output pass");
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr bool expl = kMode == 0;
  constexpr int R4 = kMode == 1 ? 3 : 4;
  constexpr int kV = 16 / sizeof(T);
  constexpr int G = kF1G;
  constexpr int IZ = 1, IDL = 2, kNAct = 3;
  const int tid = threadIdx.x;
  const int b = blockIdx.z, d0 = blockIdx.x * G, d = d0 + tid;
  const int dir = (int)blockIdx.y % p.ndir;
  const int sidx = (int)blockIdx.y / p.ndir;                            // time segment of this CTA
  const bool ok = d < p.dim;
  const int L = p.seqlen, nckpt = (L + BIMAMBA_CKPT - 1) / BIMAMBA_CKPT;
  // chunks [c_beg, nck) of scan time belong to this CTA
  const int c_beg = sidx * (sa_.seg_len / kT);
  const int nck = (min(L, (sidx + 1) * sa_.seg_len) + kT - 1) / kT;
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;
  const int R = expl ? 0 : p.dt_rank;

  const T* gu = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds;
  const T* gz = kGate ? reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds : nullptr;
  const T* gd = expl ? reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds : nullptr;
  const T* gbc = reinterpret_cast<const T*>(p.bc) + (int64_t)b * p.bc_bs + (int64_t)dir * p.bc_ds;
  const T* gdtr = expl ? nullptr : reinterpret_cast<const T*>(p.dtr) + (int64_t)b * p.dtr_bs + (int64_t)dir * p.dtr_ds;
  const int64_t obase = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds;

  float* s_xf = reinterpret_cast<float*>(smem_raw);                 // [16][kXW] rows as fp32
  T* s_xr = reinterpret_cast<T*>(s_xf + kT * kXW);                  // [2][16][kXW] rows as staged
  T* s_act = s_xr + 2 * kT * kXW;                                   // [2][3][16][G]

  const bool dim_vec = (p.dim % kV) == 0;
  const bool vec_u = dim_vec && aligned16(gu + d0) && (p.u_ts % kV) == 0;
  const bool vec_z = kGate && dim_vec && aligned16(gz + d0) && (p.z_ts % kV) == 0;
  const bool vec_d = expl && dim_vec && aligned16(gd + d0) && (p.delta_ts % kV) == 0;
  const bool vec_bc = aligned16(gbc) && (p.bc_ts % kV) == 0;
  const bool vec_dtr = !expl && (p.flags & BIMAMBA_FLAG_DTR_PADDED) && aligned16(gdtr) && (p.dtr_ts % kV) == 0;
  const bool fast = vec_u && (!kGate || vec_z) && (expl ? vec_d : vec_dtr) && vec_bc;
  constexpr int VPR = G / kV, VT = kT * VPR;
  constexpr int BV = 2 * kN / kV, DV = 16 / kV, RV = BV + (expl ? 0 : DV);

  auto stage = [&](int c0, int bf) {
    auto row_of = [&](int i) -> int64_t {
      const int tau = c0 * kT + i;
      return tau < L ? (int64_t)(dir ? (L - 1 - tau) : tau) : (int64_t)-1;
    };
    T* sa = s_act + bf * kNAct * kT * G;
    T* sx = s_xr + bf * kT * kXW;
    if (fast) {
#pragma unroll
      for (int k = 0; k < VT / G; ++k) {
        const int e = tid + k * G, i = e / VPR, v = e - i * VPR;
        const int64_t t = row_of(i);
        const int c = d0 + v * kV;
        const bool okv = t >= 0 && c < p.dim;
        const int so = i * G + v * kV;
        cp_async16(sa + so, okv ? gu + t * p.u_ts + c : gu, okv);
        if (kGate) cp_async16(sa + IZ * kT * G + so, okv ? gz + t * p.z_ts + c : gz, okv);
        if (expl) cp_async16(sa + IDL * kT * G + so, okv ? gd + t * p.delta_ts + c : gd, okv);
      }
#pragma unroll
      for (int k = 0; k < (kT * RV + G - 1) / G; ++k) {
        const int e = tid + k * G;
        if (e < kT * RV) {
          const int i = e / RV, v = e - i * RV;
          const int64_t t = row_of(i);
          const bool okv = t >= 0;
          const T* src = v < BV ? (gbc + t * p.bc_ts + v * kV) : (gdtr + t * p.dtr_ts + (v - BV) * kV);
          cp_async16(sx + i * kXW + v * kV, okv ? src : gbc, okv);
        }
      }
      cp_async_commit();
      return;
    }
    stage_tile(sa, G, gu, p.u_ts, kT, G, d0, p.dim, vec_u, row_of, tid, G);
    if (kGate) stage_tile(sa + IZ * kT * G, G, gz, p.z_ts, kT, G, d0, p.dim, vec_z, row_of, tid, G);
    if (expl) stage_tile(sa + IDL * kT * G, G, gd, p.delta_ts, kT, G, d0, p.dim, vec_d, row_of, tid, G);
    stage_tile(sx, kXW, gbc, p.bc_ts, kT, 2 * kN, 0, 2 * kN, vec_bc, row_of, tid, G);
    if (!expl) {
      const int w = vec_dtr ? 16 : R;
      stage_tile(sx + 2 * kN, kXW, gdtr, p.dtr_ts, kT, w, 0, w, vec_dtr, row_of, tid, G);
    }
    cp_async_commit();
  };

  float2 A2[kN / 2], h[kN / 2];
  float2 wdt[2 * R4];
  float bias = 0.f, Dd = 0.f;
#pragma unroll
  for (int j = 0; j < kN / 2; ++j) {
    h[j] = make_float2(0.f, 0.f);
    A2[j] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < 2 * R4; ++q) wdt[q] = make_float2(0.f, 0.f);
  if (ok) {
#pragma unroll
    for (int j = 0; j < kN / 2; ++j) {
      A2[j].x = __ldg(p.A + (int64_t)d * kN + 2 * j) * kLog2e;
      A2[j].y = __ldg(p.A + (int64_t)d * kN + 2 * j + 1) * kLog2e;
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
  [[maybe_unused]] float sdsum = 0.f;                                    // carry pass: sum of delta over the segment
  if constexpr (kSeg == 2) {
    if (ok) {   // state entering this segment from the carries of the segments before it, in order
      const int64_t cb = ((int64_t)b * p.ndir + dir) * (sa_.nseg - 1);
      for (int k = 0; k < sidx; ++k) {
        const float4* e = reinterpret_cast<const float4*>(sa_.hend + ((cb + k) * p.dim + d) * kN);
        const float sd = sa_.sdel[(cb + k) * p.dim + d];
        const float2 dd = make_float2(sd, sd);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ev = e[q];
          const float2 x0 = __fmul2_rn(dd, A2[2 * q]), x1 = __fmul2_rn(dd, A2[2 * q + 1]);
          h[2 * q] = __ffma2_rn(make_float2(ex2_approx(x0.x), ex2_approx(x0.y)), h[2 * q], make_float2(ev.x, ev.y));
          h[2 * q + 1] = __ffma2_rn(make_float2(ex2_approx(x1.x), ex2_approx(x1.y)), h[2 * q + 1], make_float2(ev.z, ev.w));
        }
      }
    }
  }
  const int ostep = (int)(dir ? -p.out_ts : p.out_ts);
  T* const gout = reinterpret_cast<T*>(p.out) + obase + d;
  T* const gyp = p.ypre ? reinterpret_cast<T*>(p.ypre) + obase + d : nullptr;
  float* const gck = p.ckpt ? p.ckpt + ((((int64_t)b * p.ndir + dir) * nckpt) * p.dim + d) * kN : nullptr;
  const int valid_cols = 2 * kN + R;

  if (nck > c_beg) stage(c_beg, 0);
  for (int c0 = c_beg; c0 < nck; ++c0) {
    const int bf = (c0 - c_beg) & 1;
    cp_async_wait<0>();
    __syncwarp();  // chunk c0 is visible; every lane is done with chunk c0-1's buffers
    if (c0 + 1 < nck) stage(c0 + 1, bf ^ 1);
    {
      const T* sx = s_xr + bf * kT * kXW;
#pragma unroll
      for (int k = 0; k < (kT * RV + G - 1) / G; ++k) {
        const int e = tid + k * G;
        if (e < kT * RV) {
          const int i = e / RV, v = e - i * RV;
          const int o = i * kXW + v * kV;
          T raw[kV];
          *reinterpret_cast<uint4*>(raw) = *reinterpret_cast<const uint4*>(sx + o);
          float f[kV];
#pragma unroll
          for (int x = 0; x < kV; ++x) f[x] = (v < BV || v * kV + x < valid_cols) ? to_f(raw[x]) : 0.f;
#pragma unroll
          for (int x = 0; x < kV; x += 4) *reinterpret_cast<float4*>(s_xf + o + x) = make_float4(f[x], f[x + 1], f[x + 2], f[x + 3]);
        }
      }
    }
    __syncwarp();
    const T* su = s_act + bf * kNAct * kT * G + tid;
    const int tau0 = c0 * kT;
    const int nvalid = L - tau0;  // steps of this chunk that exist (>= 1)
    const int64_t off0 = (int64_t)(dir ? (L - 1 - tau0) : tau0) * p.out_ts;
    [[maybe_unused]] float csum = 0.f;   // carry pass: this chunk's sum of delta (two-level summation)
#pragma unroll 8
    for (int i = 0; i < kT; ++i) {
      const bool live = ok && i < nvalid;
      if (kSeg != 1 && (i & (BIMAMBA_CKPT - 1)) == 0 && gck && live) {  // state entering this 8-step chunk
        float4* ck = reinterpret_cast<float4*>(gck + (int64_t)((tau0 + i) / BIMAMBA_CKPT) * p.dim * kN);
#pragma unroll
        for (int q = 0; q < 4; ++q) ck[q] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
      }
      const float4* xr = reinterpret_cast<const float4*>(s_xf + i * kXW);
      const float u = to_f(su[i * G]);
      float draw;
      if (expl) {
        draw = bias + to_f(su[(IDL * kT + i) * G]);
      } else {
        float2 acc0 = make_float2(bias, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < R4; ++q) {
          const float4 x = xr[8 + q];
          acc0 = __ffma2_rn(wdt[2 * q], make_float2(x.x, x.y), acc0);
          acc1 = __ffma2_rn(wdt[2 * q + 1], make_float2(x.z, x.w), acc1);
        }
        const float2 acc = __fadd2_rn(acc0, acc1);
        draw = acc.x + acc.y;
      }
      const float spl = softplus_f(draw);
      const float delta = softplus ? spl : draw;
      const float du = delta * u;
      const float2 dd = make_float2(delta, delta), duu = make_float2(du, du);
      if constexpr (kSeg == 1) {   // carry pass: the recurrence only
        csum += delta;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 Bq = xr[q];
          const float2 x0 = __fmul2_rn(dd, A2[2 * q]), x1 = __fmul2_rn(dd, A2[2 * q + 1]);
          h[2 * q] = __ffma2_rn(make_float2(ex2_approx(x0.x), ex2_approx(x0.y)), h[2 * q], __fmul2_rn(duu, make_float2(Bq.x, Bq.y)));
          h[2 * q + 1] = __ffma2_rn(make_float2(ex2_approx(x1.x), ex2_approx(x1.y)), h[2 * q + 1], __fmul2_rn(duu, make_float2(Bq.z, Bq.w)));
        }
        continue;
      }
      float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 Bq = xr[q], Cq = xr[4 + q];
        {
          const float2 x = __fmul2_rn(dd, A2[2 * q]);
          const float2 a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          h[2 * q] = __ffma2_rn(a, h[2 * q], __fmul2_rn(duu, make_float2(Bq.x, Bq.y)));
          ya = __ffma2_rn(make_float2(Cq.x, Cq.y), h[2 * q], ya);
        }
        {
          const float2 x = __fmul2_rn(dd, A2[2 * q + 1]);
          const float2 a = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          h[2 * q + 1] = __ffma2_rn(a, h[2 * q + 1], __fmul2_rn(duu, make_float2(Bq.z, Bq.w)));
          yb = __ffma2_rn(make_float2(Cq.z, Cq.w), h[2 * q + 1], yb);
        }
      }
      const float2 ys = __fadd2_rn(ya, yb);
      float y = fmaf(Dd, u, ys.x + ys.y);
      const int64_t off = off0 + i * ostep;
      if (gyp && live) gyp[off] = from_f<T>(y);
      if (kGate) {
        const float z = to_f(su[(IZ * kT + i) * G]);
        y *= z * sigmoid_f(z);
      }
      if (live) gout[off] = from_f<T>(y);
    }
    if constexpr (kSeg == 1) sdsum += csum;
  }
  if constexpr (kSeg == 1) {
    if (ok) {
      const int64_t ci = (((int64_t)b * p.ndir + dir) * (sa_.nseg - 1) + sidx) * p.dim + d;
      float4* e = reinterpret_cast<float4*>(sa_.hend + ci * kN);
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
      sa_.sdel[ci] = sdsum;
    }
  }
}

// ---- time split: carry pass over segments 0 .. nseg-2, then the output pass over all segments (two launches; the second
// starts with griddepcontrol.wait, so the carries are visible)
template <typename T, int kMode, bool kGate>
static void launch_split3(const bimamba_scan_desc* d, const SegArgs& sa, cudaStream_t st) {
  const size_t smem = (size_t)kT * kXW * 4 + (size_t)2 * kT * kXW * sizeof(T) + (size_t)2 * 3 * kT * kF1G * sizeof(T);
  const unsigned ng = (unsigned)((d->dim + kF1G - 1) / kF1G);
  launch_k(scan_fwd_warp_seg_kernel<T, kMode, false, 1>, dim3(ng, (unsigned)(d->ndir * (sa.nseg - 1)), (unsigned)d->batch), kF1G,
           smem, st, *d, sa);
  launch_k(scan_fwd_warp_seg_kernel<T, kMode, kGate, 2>, dim3(ng, (unsigned)(d->ndir * sa.nseg), (unsigned)d->batch), kF1G, smem,
           st, *d, sa);
}

template <typename T>
static void launch_split2(const bimamba_scan_desc* d, const SegArgs& sa, cudaStream_t st) {
  const bool gate = d->z != nullptr;
  const int mode = d->delta ? 0 : (d->dt_rank <= 12 ? 1 : 2);
  if (gate) {
    if (mode == 0) launch_split3<T, 0, true>(d, sa, st);
    else if (mode == 1) launch_split3<T, 1, true>(d, sa, st);
    else launch_split3<T, 2, true>(d, sa, st);
  } else {
    if (mode == 0) launch_split3<T, 0, false>(d, sa, st);
    else if (mode == 1) launch_split3<T, 1, false>(d, sa, st);
    else launch_split3<T, 2, false>(d, sa, st);
  }
}

int check_desc(const bimamba_scan_desc* d, bool bwd);  // api.cu

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_scan_fwd_split_plan(int batch, int ndir, int seqlen, int dim, int io_dtype, int* nseg, int* seg_len) {
  int ns = 1, sl = seqlen;
  const int force = g_tune[BIMAMBA_TUNE_SCAN_SPLIT];   // 0 = automatic, 1 = never, >= 2 = that many segments (parity tests)
  const int64_t warps = (int64_t)batch * ndir * ((dim + kF1G - 1) / kF1G);
  if (force >= 2) {
    ns = force;
  } else if (force == 0 && io_dtype != BIMAMBA_F32 && warps > 0 && warps < 148 * 6 && seqlen >= 2048) {
    // fewer than 6 warps of channel lanes per SM and a long walk: fill the ~12 warps per SM the kernel's registers allow
    // (one wave: 14 warps per SM fit at the output pass's 142 registers).  Measured on a B200 (profiles/r2_scan_split_ab.jsonl,
    // 8192 steps x batch 64, bf16): serial 2.30 ms, 2 / 3 / 4 / 6 segments 2.02 / 1.69 / 1.96 / 1.86 ms; at 4096 x 128
    // (7.8 warps per SM) every split is slower than the serial 1.37 ms, and with fp32 I/O the serial walk wins even at
    // 8192 x 64 (1.82 ms against 1.96 - 2.10 ms) - so fp32 never splits by itself.
    ns = (int)((148 * 12) / warps);
    if (ns > 8) ns = 8;
  }
  if (ns > 1 && seqlen > 0) {
    sl = ((seqlen + ns - 1) / ns + kT - 1) / kT * kT;   // whole 16-step chunks
    if (force == 0 && sl < 512) sl = 512;                // a segment must amortise its carry pass
    ns = (seqlen + sl - 1) / sl;                         // the last segment is not empty
  }
  if (ns < 2) { ns = 1; sl = seqlen; }
  if (nseg) *nseg = ns;
  if (seg_len) *seg_len = sl;
  return 0;
}

extern "C" size_t bimamba_scan_fwd_split_workspace_bytes(int batch, int ndir, int dim, int nseg) {
  if (batch <= 0 || ndir <= 0 || dim <= 0 || nseg < 2) return 0;
  return (size_t)batch * ndir * (nseg - 1) * dim * (kN + 1) * 4;   // end states (16) + sum of delta (1) per channel
}

extern "C" int bimamba_selective_scan_fwd_split(const bimamba_scan_desc* d, int nseg, int seg_len, float* carry,
                                                size_t carry_bytes, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;
  int rc = check_desc(d, false);
  if (rc) return rc;
  if (nseg < 2) return bimamba_selective_scan_fwd(d, stream);
  if (seg_len < kT || (seg_len % kT) != 0 || (int64_t)(nseg - 1) * seg_len >= d->seqlen || (int64_t)nseg * seg_len < d->seqlen) {
    set_err("scan_fwd_split: seg_len must be a multiple of 16 with (nseg-1)*seg_len < seqlen <= nseg*seg_len (use bimamba_scan_fwd_split_plan)");
    return -5;
  }
  if ((int64_t)d->ndir * nseg > 65535) { set_err("scan_fwd_split: too many segments"); return -3; }
  const size_t need = bimamba_scan_fwd_split_workspace_bytes(d->batch, d->ndir, d->dim, nseg);
  if (!carry || carry_bytes < need || !aligned16(carry)) { set_err("scan_fwd_split: carry workspace missing, too small or misaligned"); return -10; }
  SegArgs sa;
  sa.hend = carry;
  sa.sdel = carry + (size_t)d->batch * d->ndir * (nseg - 1) * d->dim * kN;
  sa.nseg = nseg;
  sa.seg_len = seg_len;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d->io_dtype) {
    case BIMAMBA_F32: launch_split2<float>(d, sa, st); break;
    case BIMAMBA_BF16: launch_split2<__nv_bfloat16>(d, sa, st); break;
    default: launch_split2<__half>(d, sa, st); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
