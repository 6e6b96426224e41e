// Selective scan, backward, both time directions in one launch.  sm_100a.
//
// Gradients (autograd of src/models/modules/mamba_block.py:80-120, :61):
//   g = dout * silu(z);  dz = dout * ypre * silu'(z)
//   dh[t] = g[t] C[t] + a[t+1] dh[t+1]                       (reverse-time recurrence)
//   ddelta[t] = sum_n dh a h[t-1] A + u sum_n dh B;  du[t] = g D + delta sum_n dh B
//   dB[t,n] = sum_d dh delta u;  dC[t,n] = sum_d g h;  dA[n] = sum_t dh a h[t-1] delta
//
// Mapping.  The backward needs h[t-1] and a[t] of every step while walking time in reverse, so a
// thread cannot keep all 16 states of a channel for a whole chunk.  Instead
//   * a thread owns a QUAD of states (2 x float2, so the recurrences issue as FMUL2 / FFMA2) of one
//     channel: lane = 4 * channel_in_warp + quad, a warp works on 8 channels, a CTA (4 warps) on a
//     "pass" of 32 channels and on `group_channels` = 32 * passes channels of one (batch, dir).
//   * time is walked in 8-step chunks from the last to the first (8 = the forward's checkpoint
//     interval).  Per (chunk, pass) the raw tiles (u, dout, z, ypre [8 x 32 channels], the chunk's
//     B|C|dt_r rows and the 32 x 16 checkpoint tile) are staged by 16-byte cp.async into double
//     buffers while the previous item computes: ONE block barrier per item.
//   * the four lanes of a quad also split the chunk's 8 ELEMENTS of their channel two each: a lane
//     computes delta (fused dt projection + softplus), delta*u, g and dz of its two elements once,
//     the quad exchanges them by width-4 shuffles, and after the recurrence the 4-lane
//     reduce-scatter of the n-sums lands exactly those two elements back on the lane that owns
//     them, which finishes du / ddelta / dD / dbias - no shared-memory round trip, no extra barrier.
//   * the chunk is re-run forward from its checkpoint keeping a[t], h[t] in 64 registers, then the
//     reverse recurrence runs over the same registers - no (B, L, D, N) tensor, 16 exps per element.
//   * dB/dC accumulate in registers over the passes of a chunk, then are summed over the CTA's 32
//     (warp, channel) lanes in fixed order through shared memory and written as per-group partials;
//     dA/dD/dbias are per-(batch, dir, channel) partials.  All cross-CTA sums are done in fixed order
//     by bimamba_reduce_partials: the whole backward is deterministic (no atomics).
#include "common.cuh"

namespace bimamba {

#ifndef BIMAMBA_BWD_MINB
#define BIMAMBA_BWD_MINB 2
#endif
constexpr int kBT = BIMAMBA_CKPT;          // steps per backward chunk == checkpoint interval (8)
constexpr int kBW = 4;                     // warps per CTA
constexpr int kBThreads = kBW * 32;
constexpr int kBC = 32;                    // channels per pass
constexpr int kMaxKP = 4;                  // passes per CTA: group_channels = 32 * passes
constexpr int kNRaw = 5;                   // raw tiles: u, dout, z, ypre, delta
constexpr int kRedStride = kBT * 2 * kN + 16;  // floats per (warp, channel) partial in the dB/dC reduction
static_assert(kBC == 2 * kN, "the dB/dC column sum maps one thread column to one [dB|dC] column");
static_assert(kBT == 8, "a quad of lanes owns 2 x 4 = 8 steps");

// Sum over the 4 lanes of a quad of v[0..7]; lane q returns the totals of steps 2q and 2q+1.
// Fixed tree -> deterministic.
__device__ __forceinline__ void reduce_scatter4(const float (&v)[8], int q, float& r0, float& r1) {
  const bool b1 = q & 2, b0 = q & 1;
  float w4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b1 ? v[i] : v[i + 4];
    const float keep = b1 ? v[i + 4] : v[i];
    w4[i] = keep + __shfl_xor_sync(kFull, send, 2);
  }
  float w2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b0 ? w4[i] : w4[i + 2];
    const float keep = b0 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(kFull, send, 1);
  }
  r0 = w2[0];
  r1 = w2[1];
}

template <typename T>
struct BwdSmem {
  static constexpr int kV = 16 / sizeof(T);
  static constexpr int RS = kBC + kV;      // padded row of a raw tile (bank-conflict-free strided reads)
  static constexpr size_t raw_elems = (size_t)kNRaw * kBT * RS;
  static constexpr size_t raw_bytes = 2 * raw_elems * sizeof(T);
  static constexpr size_t xr_bytes = (size_t)kBT * kXW * sizeof(T);
  static constexpr size_t f32_floats = 2 * kBC * kN /*ckpt tiles*/ + kBT * kXW + 32 * kRedStride +
                                       kMaxKP * kBC * (BIMAMBA_MAX_DT_RANK + 2 + 2 * kN + 2 * 4);
  static constexpr size_t total = raw_bytes + xr_bytes + f32_floats * sizeof(float);
};

template <typename T>
__global__ void __launch_bounds__(kBThreads, BIMAMBA_BWD_MINB) scan_bwd_kernel(const bimamba_scan_desc p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = BwdSmem<T>;
  constexpr int kV = SM::kV, RS = SM::RS;
  const int tid = threadIdx.x;
  const int b = blockIdx.z, dir = blockIdx.y, G = p.group_channels, g = blockIdx.x, d0 = g * G;
  const int KP = G / kBC;
  const int ngroups = gridDim.x;
  const int warp = tid >> 5, lane = tid & 31, cw = lane >> 2, q = lane & 3;
  const int rc = warp * 8 + cw;  // channel of this thread within a pass
  const int L = p.seqlen, nsub = (L + kBT - 1) / kBT;
  const bool gated = p.z != nullptr, expl = p.delta != nullptr;
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;
  const int R = expl ? 0 : p.dt_rank;
  const int64_t bd = (int64_t)b * p.ndir + dir;

  const T* gu = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds;
  const T* gz = gated ? reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds : nullptr;
  const T* gdl = expl ? reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds : nullptr;
  const T* gbc = reinterpret_cast<const T*>(p.bc) + (int64_t)b * p.bc_bs + (int64_t)dir * p.bc_ds;
  const T* gdtr = R ? reinterpret_cast<const T*>(p.dtr) + (int64_t)b * p.dtr_bs + (int64_t)dir * p.dtr_ds : nullptr;
  const T* gdo = reinterpret_cast<const T*>(p.dout) + (int64_t)b * p.dout_bs + (int64_t)dir * p.dout_ds;
  const int64_t obase = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds;
  const bool need_yp = gated && p.dz != nullptr;
  const T* gyp = need_yp ? reinterpret_cast<const T*>(p.ypre) + obase : nullptr;
  T* gdu = reinterpret_cast<T*>(p.du) + obase;
  T* gdd = reinterpret_cast<T*>(p.ddelta) + obase;
  T* gdz = need_yp ? reinterpret_cast<T*>(p.dz) + obase : nullptr;
  // partial layout (batch, ngroups, L, ndir, 32): reducing over ngroups leaves rows ordered (b, t, dir)
  const int64_t pb_ts = (int64_t)p.ndir * 2 * kN;
  float* partB = p.dbc_part + (((int64_t)b * ngroups + g) * L) * pb_ts + dir * 2 * kN;
  const float* gck = p.ckpt ? p.ckpt + bd * nsub * (int64_t)p.dim * kN : nullptr;

  // ---- shared memory carve
  T* s_raw = reinterpret_cast<T*>(smem_raw);                       // [2][5][kBT*RS]
  T* s_xr = reinterpret_cast<T*>(smem_raw + SM::raw_bytes);        // [kBT*kXW]
  float* s_ck = reinterpret_cast<float*>(smem_raw + SM::raw_bytes + SM::xr_bytes);  // [2][kBC*16]
  float* s_xf = s_ck + 2 * kBC * kN;                               // [kBT*kXW]
  float* s_red = s_xf + kBT * kXW;                                 // [32][kRedStride]
  float* s_wdt = s_red + 32 * kRedStride;                          // [G][16]
  float* s_bias = s_wdt + kMaxKP * kBC * BIMAMBA_MAX_DT_RANK;      // [G]
  float* s_D = s_bias + kMaxKP * kBC;                              // [G]
  float* s_m = s_D + kMaxKP * kBC;                                 // [G][16]  reverse carry a*dh
  float* s_dA = s_m + kMaxKP * kBC * kN;                           // [G][16]
  float* s_acc = s_dA + kMaxKP * kBC * kN;                         // [2][4][G]: dD / dbias partial of (quad lane, channel)

  const bool dim_vec = (p.dim % kV) == 0 && (d0 % kV) == 0;
  const bool vec_u = dim_vec && aligned16(gu + d0) && (p.u_ts % kV) == 0;
  const bool vec_z = gated && dim_vec && aligned16(gz + d0) && (p.z_ts % kV) == 0;
  const bool vec_do = dim_vec && aligned16(gdo + d0) && (p.dout_ts % kV) == 0;
  const bool vec_yp = need_yp && dim_vec && aligned16(gyp + d0) && (p.out_ts % kV) == 0;
  const bool vec_dl = expl && dim_vec && aligned16(gdl + d0) && (p.delta_ts % kV) == 0;
  const bool vec_bc = aligned16(gbc) && (p.bc_ts % kV) == 0;
  const bool vec_dtr = R && (p.flags & BIMAMBA_FLAG_DTR_PADDED) && aligned16(gdtr) && (p.dtr_ts % kV) == 0;
  const bool vec_ck = gck != nullptr && aligned16(gck);  // rows of dim * 16 floats
  const int col_end = min(p.dim, d0 + G);

  // Fast staging path (every tensor 16-byte friendly): a fixed (tensor, row, vector) assignment per
  // thread, so an item costs each thread a handful of address computations and cp.async's.
  constexpr int VPR = kBC / kV;        // vectors per tile row
  constexpr int VT = kBT * VPR;        // vectors per activation tile
  const bool fast = vec_u && vec_do && (!gated || vec_z) && (!need_yp || vec_yp) && (!expl || vec_dl) && vec_bc &&
                    (!R || vec_dtr) && (!gck || vec_ck);
  // stage the raw tiles of item (chunk c0, pass k) into buffer bf; with_rows also stages the chunk's rows
  auto stage = [&](int c0, int k, int bf, bool with_rows) {
    auto row_of = [&](int i) -> int64_t {
      const int tau = c0 * kBT + i;
      return tau < L ? (int64_t)(dir ? (L - 1 - tau) : tau) : (int64_t)-1;
    };
    const int c_lo = d0 + k * kBC;
    T* sr = s_raw + bf * SM::raw_elems;
    float* sc = s_ck + bf * kBC * kN;
    if (fast) {
#pragma unroll
      for (int it = 0; it < (kNRaw * VT + kBThreads - 1) / kBThreads; ++it) {
        const int e = tid + it * kBThreads;
        const int ts_ = e / VT;          // warp-uniform tensor index: 0 u, 1 dout, 2 z, 3 ypre, 4 delta
        if (ts_ < kNRaw) {
          const T* gp = ts_ == 0 ? gu : ts_ == 1 ? gdo : ts_ == 2 ? gz : ts_ == 3 ? gyp : gdl;
          const int64_t gts = ts_ == 0 ? p.u_ts : ts_ == 1 ? p.dout_ts : ts_ == 2 ? p.z_ts : ts_ == 3 ? p.out_ts : p.delta_ts;
          if (gp != nullptr) {
            const int r = e - ts_ * VT, i = r / VPR, v = r - i * VPR;
            const int64_t t = row_of(i);
            const int c = c_lo + v * kV;
            const bool ok = t >= 0 && c < col_end;
            cp_async16(sr + ts_ * kBT * RS + i * RS + v * kV, ok ? (gp + t * gts + c) : gp, ok);
          }
        }
      }
      if (gck) {  // 32 channels x 16 states = 128 float4
        const int c = c_lo + (tid >> 2);
        const bool ok = c < col_end;
        cp_async16(sc + tid * 4, ok ? (gck + ((int64_t)c0 * p.dim + c) * kN + (tid & 3) * 4) : gck, ok);
      }
      if (with_rows) {
        constexpr int BV = 2 * kN / kV, DV = 16 / kV;   // vectors per row: B|C and padded dt_r
        const int vpr = BV + (R ? DV : 0);
        if (tid < kBT * vpr) {
          const int i = tid / vpr, v = tid - i * vpr;
          const int64_t t = row_of(i);
          const bool ok = t >= 0;
          const T* src = v < BV ? (gbc + t * p.bc_ts + v * kV) : (gdtr + t * p.dtr_ts + (v - BV) * kV);
          cp_async16(s_xr + i * kXW + v * kV, ok ? src : gbc, ok);
        }
      }
      cp_async_commit();
      return;
    }
    stage_tile(sr, RS, gu, p.u_ts, kBT, kBC, c_lo, col_end, vec_u, row_of, tid, kBThreads);
    stage_tile(sr + kBT * RS, RS, gdo, p.dout_ts, kBT, kBC, c_lo, col_end, vec_do, row_of, tid, kBThreads);
    if (gated) stage_tile(sr + 2 * kBT * RS, RS, gz, p.z_ts, kBT, kBC, c_lo, col_end, vec_z, row_of, tid, kBThreads);
    if (need_yp) stage_tile(sr + 3 * kBT * RS, RS, gyp, p.out_ts, kBT, kBC, c_lo, col_end, vec_yp, row_of, tid, kBThreads);
    if (expl) stage_tile(sr + 4 * kBT * RS, RS, gdl, p.delta_ts, kBT, kBC, c_lo, col_end, vec_dl, row_of, tid, kBThreads);
    if (gck) {  // checkpoint tile: channels [c_lo, c_lo+32) x 16 states = one contiguous run of floats
      auto one = [&](int) -> int64_t { return (int64_t)c0; };
      stage_tile(sc, kBC * kN, gck, (int64_t)p.dim * kN, 1, kBC * kN, c_lo * kN, col_end * kN, vec_ck, one, tid, kBThreads);
    }
    if (with_rows) {
      stage_tile(s_xr, kXW, gbc, p.bc_ts, kBT, 2 * kN, 0, 2 * kN, vec_bc, row_of, tid, kBThreads);
      if (R) {
        const int w = vec_dtr ? 16 : R;
        stage_tile(s_xr + 2 * kN, kXW, gdtr, p.dtr_ts, kBT, w, 0, w, vec_dtr, row_of, tid, kBThreads);
      }
    }
    cp_async_commit();
  };

  if (nsub > 0) stage(nsub - 1, 0, 0, true);

  // per-CTA constants and accumulators
  for (int e = tid; e < G * BIMAMBA_MAX_DT_RANK; e += kBThreads) {
    const int cc = e / BIMAMBA_MAX_DT_RANK, r = e % BIMAMBA_MAX_DT_RANK;
    const int c = d0 + cc;
    s_wdt[e] = (c < p.dim && r < R) ? __ldg(p.Wdt + (int64_t)c * R + r) : 0.f;
  }
  for (int cc = tid; cc < G; cc += kBThreads) {
    const int c = d0 + cc;
    s_bias[cc] = (c < p.dim && p.delta_bias) ? __ldg(p.delta_bias + c) : 0.f;
    s_D[cc] = (c < p.dim && p.D) ? __ldg(p.D + c) : 0.f;
  }
  for (int e = tid; e < G * kN; e += kBThreads) {
    s_m[e] = 0.f;
    s_dA[e] = 0.f;
  }
  for (int e = tid; e < 2 * 4 * kMaxKP * kBC; e += kBThreads) s_acc[e] = 0.f;
  if (!gck) {
    for (int e = tid; e < 2 * kBC * kN; e += kBThreads) s_ck[e] = 0.f;  // single chunk: the start state is zero
  }

  float2 dBa[kBT][2], dCa[kBT][2];
#pragma unroll
  for (int i = 0; i < kBT; ++i) {
    dBa[i][0] = dBa[i][1] = make_float2(0.f, 0.f);
    dCa[i][0] = dCa[i][1] = make_float2(0.f, 0.f);
  }
  const int R4 = (R + 3) >> 2;
  const int pcc = tid & (kBC - 1);  // [dB|dC] column of this thread in the per-chunk column sum

  int item = 0;
  for (int c0 = nsub - 1; c0 >= 0; --c0) {
    const int tau0 = c0 * kBT;
#pragma unroll 1
    for (int k = 0; k < KP; ++k, ++item) {
      const int bf = item & 1;
      cp_async_wait<0>();
      __syncthreads();  // (1) this item's tiles are visible; every thread is done with the previous item
      if (k == 0) {
        const int valid = 2 * kN + R;
        for (int e = tid; e < kBT * kXW; e += kBThreads) {
          const int col = e % kXW;
          s_xf[e] = col < valid ? to_f(s_xr[e]) : 0.f;
        }
        __syncthreads();  // rows ready; the raw rows may be restaged
      }
      {  // prefetch the next item into the other buffers
        int nk = k + 1, nc = c0;
        if (nk >= KP) {
          nk = 0;
          nc = c0 - 1;
        }
        if (nc >= 0) stage(nc, nk, bf ^ 1, nk == 0);
      }

      const int cg = k * kBC + rc;  // channel within the group
      const int c = d0 + cg;
      const bool okc = c < p.dim && cg < G;
      const T* sr = s_raw + bf * SM::raw_elems + rc;
      const float4 hs = *reinterpret_cast<const float4*>(s_ck + bf * kBC * kN + rc * kN + 4 * q);

      // ---- this lane's two elements (steps 2q, 2q+1 of channel rc)
      float e_u[2], e_dl[2], e_dlu[2], e_g[2], e_sp[2];
      int64_t e_off[2];
      {
        const float bias = s_bias[cg];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int i = 2 * q + j;
          const int tau = tau0 + i;
          float dl = 0.f, dlu = 0.f, gg = 0.f, uu = 0.f, sp = 0.f;
          const int64_t t = dir ? (L - 1 - tau) : tau;
          e_off[j] = t * p.out_ts + c;
          if (okc && tau < L) {
            uu = to_f(sr[i * RS]);
            const float dov = to_f(sr[(kBT + i) * RS]);
            float draw = bias;
            if (expl) {
              draw += to_f(sr[(4 * kBT + i) * RS]);
            } else {
              const float4* xr = reinterpret_cast<const float4*>(s_xf + i * kXW + 2 * kN);
              const float4* wr = reinterpret_cast<const float4*>(s_wdt + cg * BIMAMBA_MAX_DT_RANK);
              float2 acc0 = make_float2(draw, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
              for (int r4 = 0; r4 < 4; ++r4) {
                if (r4 < R4) {
                  const float4 x = xr[r4], w = wr[r4];
                  acc0 = __ffma2_rn(make_float2(w.x, w.y), make_float2(x.x, x.y), acc0);
                  acc1 = __ffma2_rn(make_float2(w.z, w.w), make_float2(x.z, x.w), acc1);
                }
              }
              draw = (acc0.x + acc0.y) + (acc1.x + acc1.y);
            }
            if (softplus) {
              dl = softplus_f(draw);
              sp = draw > 20.f ? 1.f : sigmoid_f(draw);
            } else {
              dl = draw;
              sp = 1.f;
            }
            dlu = dl * uu;
            gg = dov;
            if (gated) {
              const float zz = to_f(sr[(2 * kBT + i) * RS]);
              const float sg = sigmoid_f(zz);
              gg = dov * zz * sg;
              if (need_yp) {
                const float yp = to_f(sr[(3 * kBT + i) * RS]);
                gdz[e_off[j]] = from_f<T>(dov * yp * sg * (1.f + zz * (1.f - sg)));
              }
            }
          }
          e_u[j] = uu;
          e_dl[j] = dl;
          e_dlu[j] = dlu;
          e_g[j] = gg;
          e_sp[j] = sp;
        }
      }
      // quad exchange: every lane gets delta, delta*u, g of all 8 steps of its channel
      float dq[kBT], uq[kBT], gq[kBT];
#pragma unroll
      for (int i = 0; i < kBT; ++i) {
        dq[i] = __shfl_sync(kFull, e_dl[i & 1], i >> 1, 4);
        uq[i] = __shfl_sync(kFull, e_dlu[i & 1], i >> 1, 4);
        gq[i] = __shfl_sync(kFull, e_g[i & 1], i >> 1, 4);
      }

      // ---- recurrence: this thread owns states 4q..4q+3 of channel rc of the pass
      float4 A4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (okc) A4 = __ldg(reinterpret_cast<const float4*>(p.A + (int64_t)c * kN) + q);
      const float2 A2a = make_float2(A4.x * kLog2e, A4.y * kLog2e), A2b = make_float2(A4.z * kLog2e, A4.w * kLog2e);
      // re-run the chunk forward from the checkpoint, keeping a[t], h[t]
      float2 a[kBT][2], hh[kBT][2];
      {
        float2 h0 = make_float2(hs.x, hs.y), h1 = make_float2(hs.z, hs.w);
#pragma unroll
        for (int i = 0; i < kBT; ++i) {
          const float4 B4 = *reinterpret_cast<const float4*>(s_xf + i * kXW + 4 * q);
          const float2 dd = make_float2(dq[i], dq[i]), uu = make_float2(uq[i], uq[i]);
          const float2 x0 = __fmul2_rn(dd, A2a), x1 = __fmul2_rn(dd, A2b);
          a[i][0] = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
          a[i][1] = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
          h0 = __ffma2_rn(a[i][0], h0, __fmul2_rn(uu, make_float2(B4.x, B4.y)));
          h1 = __ffma2_rn(a[i][1], h1, __fmul2_rn(uu, make_float2(B4.z, B4.w)));
          hh[i][0] = h0;
          hh[i][1] = h1;
        }
      }
      // reverse recurrence:  dh_i = g_i C_i + m_{i+1},  m_i = a_i dh_i
      float4* pm = reinterpret_cast<float4*>(s_m + cg * kN + 4 * q);
      const float4 m4 = *pm;
      float2 m0 = make_float2(m4.x, m4.y), m1 = make_float2(m4.z, m4.w);
      float2 dA0 = make_float2(0.f, 0.f), dA1 = make_float2(0.f, 0.f);
      float vA[kBT], vU[kBT];
#pragma unroll
      for (int i = kBT - 1; i >= 0; --i) {
        const float4 B4 = *reinterpret_cast<const float4*>(s_xf + i * kXW + 4 * q);
        const float4 C4 = *reinterpret_cast<const float4*>(s_xf + i * kXW + kN + 4 * q);
        const float2 gg = make_float2(gq[i], gq[i]), dd = make_float2(dq[i], dq[i]), uu = make_float2(uq[i], uq[i]);
        const float2 dh0 = __ffma2_rn(gg, make_float2(C4.x, C4.y), m0);
        const float2 dh1 = __ffma2_rn(gg, make_float2(C4.z, C4.w), m1);
        m0 = __fmul2_rn(a[i][0], dh0);
        m1 = __fmul2_rn(a[i][1], dh1);
        const float2 hp0 = (i == 0) ? make_float2(hs.x, hs.y) : hh[i == 0 ? 0 : i - 1][0];
        const float2 hp1 = (i == 0) ? make_float2(hs.z, hs.w) : hh[i == 0 ? 0 : i - 1][1];
        const float2 da0 = __fmul2_rn(m0, hp0), da1 = __fmul2_rn(m1, hp1);
        dA0 = __ffma2_rn(da0, dd, dA0);
        dA1 = __ffma2_rn(da1, dd, dA1);
        dBa[i][0] = __ffma2_rn(dh0, uu, dBa[i][0]);
        dBa[i][1] = __ffma2_rn(dh1, uu, dBa[i][1]);
        dCa[i][0] = __ffma2_rn(gg, hh[i][0], dCa[i][0]);
        dCa[i][1] = __ffma2_rn(gg, hh[i][1], dCa[i][1]);
        const float2 ta = __ffma2_rn(da1, A2b, __fmul2_rn(da0, A2a));
        const float2 tu = __ffma2_rn(dh1, make_float2(B4.z, B4.w), __fmul2_rn(dh0, make_float2(B4.x, B4.y)));
        vA[i] = ta.x + ta.y;   // sum_n dh a h[t-1] A (x log2e; scaled back below)
        vU[i] = tu.x + tu.y;   // sum_n dh B
      }
      *pm = make_float4(m0.x, m0.y, m1.x, m1.y);
      {
        float4* pa = reinterpret_cast<float4*>(s_dA + cg * kN + 4 * q);
        float4 acc = *pa;
        acc.x += dA0.x;
        acc.y += dA0.y;
        acc.z += dA1.x;
        acc.w += dA1.y;
        *pa = acc;
      }
      float rA[2], rU[2];
      reduce_scatter4(vA, q, rA[0], rA[1]);
      reduce_scatter4(vU, q, rU[0], rU[1]);

      // ---- finish this lane's two elements: du, ddelta, dD, dbias
      {
        const float Dd = s_D[cg];
        float dDl = 0.f, dbl = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (okc && tau0 + 2 * q + j < L) {
            dDl = fmaf(e_g[j], e_u[j], dDl);
            const float duv = fmaf(e_g[j], Dd, e_dl[j] * rU[j]);
            const float ddl = fmaf(e_u[j], rU[j], rA[j] * kLn2) * e_sp[j];
            dbl += ddl;
            gdu[e_off[j]] = from_f<T>(duv);
            gdd[e_off[j]] = from_f<T>(ddl);
          }
        }
        s_acc[q * (kMaxKP * kBC) + cg] += dDl;   // this thread is the only writer of these two slots
        s_acc[(4 + q) * (kMaxKP * kBC) + cg] += dbl;
      }
    }  // passes

    // ---- dB/dC of this chunk: sum over the CTA's 32 (warp, channel) lanes in fixed order
    {
      float* my = s_red + (warp * 8 + cw) * kRedStride;
#pragma unroll
      for (int i = 0; i < kBT; ++i) {
        *reinterpret_cast<float4*>(my + i * 2 * kN + 4 * q) = make_float4(dBa[i][0].x, dBa[i][0].y, dBa[i][1].x, dBa[i][1].y);
        *reinterpret_cast<float4*>(my + i * 2 * kN + kN + 4 * q) = make_float4(dCa[i][0].x, dCa[i][0].y, dCa[i][1].x, dCa[i][1].y);
        dBa[i][0] = dBa[i][1] = make_float2(0.f, 0.f);
        dCa[i][0] = dCa[i][1] = make_float2(0.f, 0.f);
      }
    }
    __syncthreads();  // (2)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i = (tid >> 5) + j * kBW;
      const int tau = tau0 + i;
      if (tau < L) {
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) s += s_red[w * kRedStride + i * 2 * kN + pcc];
        const int64_t t = dir ? (L - 1 - tau) : tau;
        partB[t * pb_ts + pcc] = s;
      }
    }
    // the next item's barrier (1) orders these reads before s_red is rewritten
  }

  // ---- per-channel partials
  __syncthreads();
  for (int e = tid; e < G * kN; e += kBThreads) {
    const int c = d0 + e / kN;
    if (c < p.dim) p.dA_part[(bd * p.dim + c) * kN + (e % kN)] = s_dA[e];
  }
  // dD / dbias: the 4 lanes of a quad share a channel; combine in fixed order
  for (int cg = tid; cg < G; cg += kBThreads) {
    const int c = d0 + cg;
    if (c < p.dim) {
      float sD = 0.f, sb = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        sD += s_acc[w * (kMaxKP * kBC) + cg];
        sb += s_acc[(4 + w) * (kMaxKP * kBC) + cg];
      }
      if (p.dD_part) p.dD_part[bd * p.dim + c] = sD;
      if (p.dbias_part) p.dbias_part[bd * p.dim + c] = sb;
    }
  }
}

template <typename T>
static void launch_bwd(const bimamba_scan_desc* d, cudaStream_t st) {
  cudaFuncSetAttribute(scan_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BwdSmem<T>::total);
  const int G = d->group_channels;
  dim3 grid((d->dim + G - 1) / G, d->ndir, d->batch);
  scan_bwd_kernel<T><<<grid, kBThreads, BwdSmem<T>::total, st>>>(*d);
}

int check_desc(const bimamba_scan_desc* d, bool bwd);  // api.cu

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_selective_scan_bwd(const bimamba_scan_desc* d, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;
  int rc = check_desc(d, true);
  if (rc) return rc;
  const int G = d->group_channels;
  if (G < kBC || G > kMaxKP * kBC || (G % kBC)) { set_err("backward group_channels must be 32, 64, 96 or 128 (use bimamba_scan_plan)"); return -5; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d->io_dtype) {
    case BIMAMBA_F32: launch_bwd<float>(d, st); break;
    case BIMAMBA_BF16: launch_bwd<__nv_bfloat16>(d, st); break;
    default: launch_bwd<__half>(d, st); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
