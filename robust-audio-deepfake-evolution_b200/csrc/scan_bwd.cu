// Selective scan, backward, STATE-PAIR variant (both time directions in one launch).  sm_100a.
// The default backward kernel is the one-lane-per-channel kernel of scan_bwd1.cu (2.5x fewer instructions); this one
// is the earlier design, kept as BIMAMBA_BWD_KERNEL=pair and for group widths other than 32, and holds the C entry
// point that dispatches between the two.
//
// Gradients (autograd of src/models/modules/mamba_block.py:80-120, :61):
//   g = dout * silu(z);  dz = dout * ypre * silu'(z)
//   dh[t] = g[t] C[t] + a[t+1] dh[t+1]                       (reverse-time recurrence)
//   ddelta[t] = sum_n dh a h[t-1] A + u sum_n dh B;  du[t] = g D + delta sum_n dh B
//   dB[t,n] = sum_d dh delta u;  dC[t,n] = sum_d g h;  dA[n] = sum_t dh a h[t-1] delta
//
// Mapping.  The backward needs h[t-1] and a[t] of every step while walking time in reverse; to stay at 128 registers
// (16 warps per SM) a thread here does not keep all 16 states of a channel for a whole chunk.  Instead
//   * a thread owns a PAIR of states (one float2, so the recurrences issue as FMUL2 / FFMA2) of one
//     channel: lane = 8 * channel_in_warp + pair, a warp works on 4 channels, a CTA (8 warps) on a
//     "pass" of 32 channels and on `group_channels` = 32 * passes channels of one (batch, dir).
//     The per-thread history (a[t], h[t]) and dB/dC accumulators are 4 x 8 float2 = 64 registers, which
//     keeps the kernel under 128 registers: 16 warps per SM.
//   * time is walked in 8-step chunks from the last to the first (8 = the forward's checkpoint
//     interval).  Per (chunk, pass) the raw tiles (u, dout, z, ypre [8 x 32 channels], the chunk's
//     B|C|dt_r rows and the 32 x 16 checkpoint tile) are staged by 16-byte cp.async into double
//     buffers while the previous item computes: ONE block barrier per item.
//   * the eight lanes of a channel also own the chunk's 8 ELEMENTS of that channel, one each: a lane
//     computes delta (fused dt projection + softplus), delta*u, g and dz of its element once, the
//     octet exchanges them by width-8 shuffles, and after the recurrence the 8-lane reduce-scatter of
//     the n-sums lands exactly that element back on the lane that owns it, which finishes du / ddelta
//     / dD / dbias - no shared-memory round trip, no extra barrier.
//   * the chunk is re-run forward from its checkpoint keeping a[t], h[t] in registers, then the
//     reverse recurrence runs over the same registers - no (B, L, D, N) tensor, 16 exps per element.
//   * dB/dC accumulate in registers over the passes of a chunk, then are summed over the CTA's 32
//     (warp, channel) lanes in fixed order through shared memory and written as per-group partials;
//     dA/dD/dbias are per-(batch, dir, channel) partials.  All cross-CTA sums are done in fixed order
//     by bimamba_reduce_partials: the whole backward is deterministic (no atomics).
#include <cstdlib>

#include "common.cuh"

namespace bimamba {

#ifndef BIMAMBA_BWD_MINB
#define BIMAMBA_BWD_MINB 2
#endif
constexpr int kBT = BIMAMBA_CKPT;          // steps per backward chunk == checkpoint interval (8)
constexpr int kBW = 8;                     // warps per CTA
constexpr int kBThreads = kBW * 32;
constexpr int kBC = 32;                    // channels per pass (4 per warp)
constexpr int kMaxKP = 4;                  // passes per CTA: group_channels = 32 * passes
constexpr int kNRaw = 5;                   // raw tiles: u, dout, z, ypre, delta
constexpr int kRedStride = kBT * 2 * kN + 16;  // floats per (warp, channel) partial in the dB/dC reduction
static_assert(kBC == 2 * kN, "the dB/dC column sum maps one thread column to one [dB|dC] column");
static_assert(kBT == 8, "the 8 lanes of a channel own the 8 steps of a chunk");
static_assert(kBThreads == kBT * 2 * kN, "one thread per [dB|dC] entry of a chunk in the column sum");

// Sum over the 8 lanes of an octet of v[0..7]; lane p returns the total of step p.  Fixed tree -> deterministic.
__device__ __forceinline__ float reduce_scatter8(const float (&v)[8], int p) {
  const bool b2 = p & 4, b1 = p & 2, b0 = p & 1;
  float w4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b2 ? v[i] : v[i + 4];
    const float keep = b2 ? v[i + 4] : v[i];
    w4[i] = keep + __shfl_xor_sync(kFull, send, 4);
  }
  float w2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b1 ? w4[i] : w4[i + 2];
    const float keep = b1 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(kFull, send, 2);
  }
  const float send = b0 ? w2[0] : w2[1];
  const float keep = b0 ? w2[1] : w2[0];
  return keep + __shfl_xor_sync(kFull, send, 1);
}

template <typename T>
struct BwdSmem {
  static constexpr int kV = 16 / sizeof(T);
  static constexpr int RS = kBC + kV;      // padded row of a raw tile (bank-conflict-free strided reads)
  static constexpr size_t raw_elems = (size_t)kNRaw * kBT * RS;
  static constexpr size_t raw_bytes = 2 * raw_elems * sizeof(T);
  static constexpr size_t xr_bytes = (size_t)kBT * kXW * sizeof(T);
  static constexpr size_t f32_floats = 2 * kBC * kN /*ckpt tiles*/ + kBT * kXW + 32 * kRedStride +
                                       kMaxKP * kBC * (BIMAMBA_MAX_DT_RANK + 2 + 2 * kN + 2 * 8) + 3 * kBC * kBT;
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
  const int warp = tid >> 5, lane = tid & 31, cw = lane >> 3, pr = lane & 7;
  const int rc = warp * 4 + cw;  // channel of this thread within a pass
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
  float* s_acc = s_dA + kMaxKP * kBC * kN;                         // [2][G][8]: dD / dbias partial of (channel, octet lane)
  float* s_el = s_acc + 2 * 8 * kMaxKP * kBC;                      // [3][kBC][8]: delta, delta*u, g of the item (step-major per channel)

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
      if (gck && tid < kBC * kN / 4) {  // 32 channels x 16 states = 128 float4
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
  for (int e = tid; e < 2 * 8 * kMaxKP * kBC; e += kBThreads) s_acc[e] = 0.f;
  if (!gck) {
    for (int e = tid; e < 2 * kBC * kN; e += kBThreads) s_ck[e] = 0.f;  // single chunk: the start state is zero
  }

  float2 dBa[kBT], dCa[kBT];
#pragma unroll
  for (int i = 0; i < kBT; ++i) {
    dBa[i] = make_float2(0.f, 0.f);
    dCa[i] = make_float2(0.f, 0.f);
  }
  const int R4 = (R + 3) >> 2;

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
      const float2 hs = *reinterpret_cast<const float2*>(s_ck + bf * kBC * kN + rc * kN + 2 * pr);

      // ---- this lane's element: step pr of channel rc
      float e_u = 0.f, e_dl = 0.f, e_dlu = 0.f, e_g = 0.f, e_sp = 0.f;
      const int tau_e = tau0 + pr;
      const bool live = okc && tau_e < L;
      const int64_t e_off = (int64_t)(dir ? (L - 1 - tau_e) : tau_e) * p.out_ts + c;
      if (live) {
        const T* sr = s_raw + bf * SM::raw_elems + pr * RS + rc;
        e_u = to_f(sr[0]);
        const float dov = to_f(sr[kBT * RS]);
        float draw = s_bias[cg];
        if (expl) {
          draw += to_f(sr[4 * kBT * RS]);
        } else {
          const float4* xr = reinterpret_cast<const float4*>(s_xf + pr * kXW + 2 * kN);
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
          e_dl = softplus_f(draw);
          e_sp = draw > 20.f ? 1.f : sigmoid_f(draw);
        } else {
          e_dl = draw;
          e_sp = 1.f;
        }
        e_dlu = e_dl * e_u;
        e_g = dov;
        if (gated) {
          const float zz = to_f(sr[2 * kBT * RS]);
          const float sg = sigmoid_f(zz);
          e_g = dov * zz * sg;
          if (need_yp) {
            const float yp = to_f(sr[3 * kBT * RS]);
            gdz[e_off] = from_f<T>(dov * yp * sg * (1.f + zz * (1.f - sg)));
          }
        }
      }

      // ---- recurrence: this thread owns states 2*pr, 2*pr+1 of channel rc of the pass; delta, delta*u and g of
      // the channel's 8 steps are exchanged inside the warp through 3 x 32 floats of shared memory
      // (one STS each, then LDS.128 broadcasts: cheaper on the shared-memory pipe than 40 shuffles)
      s_el[rc * kBT + pr] = e_dl;
      s_el[(kBC + rc) * kBT + pr] = e_dlu;
      s_el[(2 * kBC + rc) * kBT + pr] = e_g;
      __syncwarp();
      const float4* pdl = reinterpret_cast<const float4*>(s_el + rc * kBT);
      const float4* pdu = reinterpret_cast<const float4*>(s_el + (kBC + rc) * kBT);
      const float4* pgg = reinterpret_cast<const float4*>(s_el + (2 * kBC + rc) * kBT);
      float2 A2 = make_float2(0.f, 0.f);
      if (okc) {
        const float2 Av = __ldg(reinterpret_cast<const float2*>(p.A + (int64_t)c * kN) + pr);
        A2 = make_float2(Av.x * kLog2e, Av.y * kLog2e);
      }
      // re-run the chunk forward from the checkpoint, keeping a[t], h[t]
      float2 a[kBT], hh[kBT];
      {
        float2 h = hs;
        const float4 d0v = pdl[0], d1v = pdl[1], u0v = pdu[0], u1v = pdu[1];
        const float dq[kBT] = {d0v.x, d0v.y, d0v.z, d0v.w, d1v.x, d1v.y, d1v.z, d1v.w};
        const float uq[kBT] = {u0v.x, u0v.y, u0v.z, u0v.w, u1v.x, u1v.y, u1v.z, u1v.w};
#pragma unroll
        for (int i = 0; i < kBT; ++i) {
          const float di = dq[i], ui = uq[i];
          const float2 B2 = *reinterpret_cast<const float2*>(s_xf + i * kXW + 2 * pr);
          const float2 x = __fmul2_rn(make_float2(di, di), A2);
          a[i] = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          h = __ffma2_rn(a[i], h, __fmul2_rn(make_float2(ui, ui), B2));
          hh[i] = h;
        }
      }
      // reverse recurrence:  dh_i = g_i C_i + m_{i+1},  m_i = a_i dh_i
      float2* pm = reinterpret_cast<float2*>(s_m + cg * kN + 2 * pr);
      float2 m = *pm;
      float2 dA2 = make_float2(0.f, 0.f);
      float vA[kBT], vU[kBT];
      const float4 d0r = pdl[0], d1r = pdl[1], u0r = pdu[0], u1r = pdu[1], g0r = pgg[0], g1r = pgg[1];
      const float dr[kBT] = {d0r.x, d0r.y, d0r.z, d0r.w, d1r.x, d1r.y, d1r.z, d1r.w};
      const float ur[kBT] = {u0r.x, u0r.y, u0r.z, u0r.w, u1r.x, u1r.y, u1r.z, u1r.w};
      const float gr[kBT] = {g0r.x, g0r.y, g0r.z, g0r.w, g1r.x, g1r.y, g1r.z, g1r.w};
#pragma unroll
      for (int i = kBT - 1; i >= 0; --i) {
        const float di = dr[i], ui = ur[i], gi = gr[i];
        const float2 B2 = *reinterpret_cast<const float2*>(s_xf + i * kXW + 2 * pr);
        const float2 C2 = *reinterpret_cast<const float2*>(s_xf + i * kXW + kN + 2 * pr);
        const float2 gg = make_float2(gi, gi);
        const float2 dh = __ffma2_rn(gg, C2, m);
        m = __fmul2_rn(a[i], dh);
        const float2 hp = (i == 0) ? hs : hh[i == 0 ? 0 : i - 1];
        const float2 da = __fmul2_rn(m, hp);
        dA2 = __ffma2_rn(da, make_float2(di, di), dA2);
        dBa[i] = __ffma2_rn(dh, make_float2(ui, ui), dBa[i]);
        dCa[i] = __ffma2_rn(gg, hh[i], dCa[i]);
        const float2 ta = __fmul2_rn(da, A2);
        const float2 tu = __fmul2_rn(dh, B2);
        vA[i] = ta.x + ta.y;   // sum_n dh a h[t-1] A (x log2e; scaled back below)
        vU[i] = tu.x + tu.y;   // sum_n dh B
      }
      *pm = m;
      {
        float2* pa = reinterpret_cast<float2*>(s_dA + cg * kN + 2 * pr);
        float2 acc = *pa;
        acc.x += dA2.x;
        acc.y += dA2.y;
        *pa = acc;
      }
      const float rA = reduce_scatter8(vA, pr);
      const float rU = reduce_scatter8(vU, pr);

      // ---- finish this lane's element: du, ddelta, dD, dbias
      {
        float dDl = 0.f, dbl = 0.f;
        if (live) {
          dDl = e_g * e_u;
          const float duv = fmaf(e_g, s_D[cg], e_dl * rU);
          dbl = fmaf(e_u, rU, rA * kLn2) * e_sp;
          gdu[e_off] = from_f<T>(duv);
          gdd[e_off] = from_f<T>(dbl);
        }
        s_acc[cg * 8 + pr] += dDl;   // this thread is the only writer of these two slots
        s_acc[(kMaxKP * kBC + cg) * 8 + pr] += dbl;
      }
    }  // passes

    // ---- dB/dC of this chunk: sum over the CTA's 32 (warp, channel) lanes in fixed order
    {
      float* my = s_red + (warp * 4 + cw) * kRedStride;
#pragma unroll
      for (int i = 0; i < kBT; ++i) {
        *reinterpret_cast<float2*>(my + i * 2 * kN + 2 * pr) = dBa[i];
        *reinterpret_cast<float2*>(my + i * 2 * kN + kN + 2 * pr) = dCa[i];
        dBa[i] = make_float2(0.f, 0.f);
        dCa[i] = make_float2(0.f, 0.f);
      }
    }
    __syncthreads();  // (2)
    {
      const int i = tid >> 5, col = tid & 31;   // one thread per [dB|dC] entry of the chunk
      const int tau = tau0 + i;
      if (tau < L) {
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) s += s_red[w * kRedStride + i * 2 * kN + col];
        const int64_t t = dir ? (L - 1 - tau) : tau;
        partB[t * pb_ts + col] = s;
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
  // dD / dbias: the 8 lanes of an octet share a channel; combine in fixed order
  for (int cg = tid; cg < G; cg += kBThreads) {
    const int c = d0 + cg;
    if (c < p.dim) {
      float sD = 0.f, sb = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        sD += s_acc[cg * 8 + w];
        sb += s_acc[(kMaxKP * kBC + cg) * 8 + w];
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
void launch_bwd_lane(const bimamba_scan_desc* d, cudaStream_t st);  // scan_bwd1.cu

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_selective_scan_bwd(const bimamba_scan_desc* d, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;
  int rc = check_desc(d, true);
  if (rc) return rc;
  const int G = d->group_channels;
  if (G < kBC || G > kMaxKP * kBC || (G % kBC)) { set_err("backward group_channels must be 32, 64, 96 or 128 (use bimamba_scan_plan)"); return -5; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const char* force = getenv("BIMAMBA_BWD_KERNEL");   // "lane" | "pair": tuning experiments and the parity tests of both
  const bool lane = (force ? force[0] == 'l' : true) && G == 32;   // the lane kernel is built for one warp per CTA
  if (lane) {
    launch_bwd_lane(d, st);
  } else {
    switch (d->io_dtype) {
      case BIMAMBA_F32: launch_bwd<float>(d, st); break;
      case BIMAMBA_BF16: launch_bwd<__nv_bfloat16>(d, st); break;
      default: launch_bwd<__half>(d, st); break;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
