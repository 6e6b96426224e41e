// Selective scan, backward, both time directions in one launch.  sm_100a.
//
// Gradients (autograd of src/models/modules/mamba_block.py:80-120, :61):
//   g = dout * silu(z);  dz = dout * ypre * silu'(z)
//   dh[t] = g[t] C[t] + a[t+1] dh[t+1]                       (reverse-time recurrence)
//   ddelta[t] = sum_n dh a h[t-1] A + u sum_n dh B;  du[t] = g D + delta sum_n dh B
//   dB[t,n] = sum_d dh delta u;  dC[t,n] = sum_d g h;  dA[n] = sum_t dh a h[t-1] delta
//
// Mapping: the backward needs h[t-1] and a[t] of every step while walking time in reverse, so a
// thread cannot own all 16 states of a channel (16 steps x 16 states x 2 values).  Instead
//   * lane = state: a half-warp owns one channel, lane n holds state n; a warp works on two
//     channels at a time and a CTA (4 warps) on a group of <= 32 channels of one (batch, dir).
//   * time is walked in 16-step chunks from the last to the first.  Per chunk the raw tiles
//     (u, z, dout, ypre [16 x 32 channels] and the B|C|dt_r rows) are staged by 16-byte cp.async
//     one chunk ahead; a per-ELEMENT pre-pass computes delta (fused dt projection + softplus),
//     delta*u, g, dz (stored straight away, coalesced) once and leaves them in shared memory,
//     from where the recurrence broadcast-reads them four steps per LDS.128.
//   * the chunk is re-run forward from its checkpoint keeping a[t], h[t] in 32 registers, then
//     the reverse recurrence runs over the same registers - no (B, L, D, N) tensor, 16 exps
//     per element.  Sums over n (ddelta, du) are 16-lane shuffle reduce-scatters that land
//     step j on lane j; du / ddelta go through a shared-memory tile so the global stores are
//     channel-contiguous.
//   * dB/dC accumulate in registers over the warp's channels, are combined over the CTA's warps
//     in fixed order through shared memory and written as per-group partials; dA/dD/dbias are
//     per-(batch, dir, channel) partials.  All cross-CTA sums are done in fixed order by
//     bimamba_reduce_partials: the whole backward is deterministic (no atomics).
#include "common.cuh"

namespace bimamba {

constexpr int kBW = 4;                     // warps per CTA
constexpr int kBThreads = kBW * 32;
constexpr int kBG = 32;                    // channels per CTA (tile width)
constexpr int kBKP = kBG / (2 * kBW);      // channel pairs per warp
constexpr int kDS = 20;                    // floats per channel row of the derived arrays (16 steps + pad)
constexpr int kOS = kBG + 1;               // row stride of the output tiles [step][channel]
static_assert(kBG == 2 * kN, "the post-pass maps one thread column to one [dB|dC] column");
constexpr int kNDer = 5;                   // derived arrays: delta, delta*u, g, u, d(delta)/d(raw)

// Sum over the 16 lanes of a half-warp of v[0..15]; lane j (within its half) returns sum of v[j].
// Fixed tree -> deterministic.
__device__ __forceinline__ float reduce_scatter16(float (&v)[16], int r) {
  const bool b3 = r & 8, b2 = r & 4, b1 = r & 2, b0 = r & 1;
  float w8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b3 ? v[i] : v[i + 8];
    const float keep = b3 ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(kFull, send, 8);
  }
  float w4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b2 ? w8[i] : w8[i + 4];
    const float keep = b2 ? w8[i + 4] : w8[i];
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

__device__ __forceinline__ float half_sum(float v) {  // all-reduce over the 16 lanes of a half-warp
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

template <typename T>
struct BwdSmem {
  static constexpr int kMaxRaw = 5;  // u, dout, z, ypre, delta
  static constexpr size_t raw_bytes = (size_t)2 * kMaxRaw * kT * kBG * sizeof(T);
  static constexpr size_t xr_bytes = (size_t)2 * kT * kXW * sizeof(T);
  static constexpr size_t f32_floats = kT * kXW + kNDer * kBG * kDS + 2 * kT * kOS + kBW * 2 * kT * kN +
                                       kBG * BIMAMBA_MAX_DT_RANK + 2 * kBG;
  static constexpr size_t total = raw_bytes + xr_bytes + f32_floats * sizeof(float);
};

template <typename T>
__global__ void __launch_bounds__(kBThreads, 3) scan_bwd_kernel(const bimamba_scan_desc p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int b = blockIdx.z, dir = blockIdx.y, G = p.group_channels, g = blockIdx.x, d0 = g * G;
  const int ngroups = gridDim.x;
  const int warp = tid >> 5, lane = tid & 31, half = lane >> 4, n = lane & 15;
  const int L = p.seqlen, nck = (L + kT - 1) / kT;
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

  // ---- shared memory carve
  using SM = BwdSmem<T>;
  T* s_raw = reinterpret_cast<T*>(smem_raw);                       // [2][5][kT*kBG]
  T* s_xr = reinterpret_cast<T*>(smem_raw + SM::raw_bytes);        // [2][kT*kXW]
  float* s_xf = reinterpret_cast<float*>(smem_raw + SM::raw_bytes + SM::xr_bytes);  // [kT*kXW]
  float* s_der = s_xf + kT * kXW;                                  // [5][kBG*kDS]
  float* s_out = s_der + kNDer * kBG * kDS;                        // [2][kT*kOS]
  float* s_red = s_out + 2 * kT * kOS;                             // [kBW][kT*32]
  float* s_wdt = s_red + kBW * 2 * kT * kN;                        // [kBG][16]
  float* s_bias = s_wdt + kBG * BIMAMBA_MAX_DT_RANK;               // [kBG]
  float* s_D = s_bias + kBG;                                       // [kBG]
  float* s_dl = s_der, *s_du = s_der + kBG * kDS, *s_g = s_der + 2 * kBG * kDS, *s_u = s_der + 3 * kBG * kDS,
        *s_sp = s_der + 4 * kBG * kDS;

  constexpr int kV = 16 / sizeof(T);
  const bool dim_vec = (p.dim % kV) == 0 && (d0 % kV) == 0;
  const bool vec_u = dim_vec && aligned16(gu + d0) && (p.u_ts % kV) == 0;
  const bool vec_z = gated && dim_vec && aligned16(gz + d0) && (p.z_ts % kV) == 0;
  const bool vec_do = dim_vec && aligned16(gdo + d0) && (p.dout_ts % kV) == 0;
  const bool vec_yp = need_yp && dim_vec && aligned16(gyp + d0) && (p.out_ts % kV) == 0;
  const bool vec_dl = expl && dim_vec && aligned16(gdl + d0) && (p.delta_ts % kV) == 0;
  const bool vec_bc = aligned16(gbc) && (p.bc_ts % kV) == 0;
  const bool vec_dtr = R && (p.flags & BIMAMBA_FLAG_DTR_PADDED) && aligned16(gdtr) && (p.dtr_ts % kV) == 0;
  const int col_end = min(p.dim, d0 + G);

  auto stage = [&](int c0, int bf) {
    auto row_of = [&](int i) -> int64_t {
      const int tau = c0 * kT + i;
      return tau < L ? (int64_t)(dir ? (L - 1 - tau) : tau) : (int64_t)-1;
    };
    T* sa = s_raw + bf * SM::kMaxRaw * kT * kBG;
    stage_tile(sa, kBG, gu, p.u_ts, kT, kBG, d0, col_end, vec_u, row_of, tid, kBThreads);
    stage_tile(sa + kT * kBG, kBG, gdo, p.dout_ts, kT, kBG, d0, col_end, vec_do, row_of, tid, kBThreads);
    if (gated) stage_tile(sa + 2 * kT * kBG, kBG, gz, p.z_ts, kT, kBG, d0, col_end, vec_z, row_of, tid, kBThreads);
    if (need_yp) stage_tile(sa + 3 * kT * kBG, kBG, gyp, p.out_ts, kT, kBG, d0, col_end, vec_yp, row_of, tid, kBThreads);
    if (expl) stage_tile(sa + 4 * kT * kBG, kBG, gdl, p.delta_ts, kT, kBG, d0, col_end, vec_dl, row_of, tid, kBThreads);
    T* sx = s_xr + bf * kT * kXW;
    stage_tile(sx, kXW, gbc, p.bc_ts, kT, 2 * kN, 0, 2 * kN, vec_bc, row_of, tid, kBThreads);
    if (R) {
      const int w = vec_dtr ? 16 : R;
      stage_tile(sx + 2 * kN, kXW, gdtr, p.dtr_ts, kT, w, 0, w, vec_dtr, row_of, tid, kBThreads);
    }
    cp_async_commit();
  };

  if (nck > 0) stage(nck - 1, (nck - 1) & 1);

  // per-CTA constants
  for (int e = tid; e < kBG * BIMAMBA_MAX_DT_RANK; e += kBThreads) {
    const int cc = e / BIMAMBA_MAX_DT_RANK, r = e % BIMAMBA_MAX_DT_RANK;
    const int c = d0 + cc;
    s_wdt[e] = (cc < G && c < p.dim && r < R) ? __ldg(p.Wdt + (int64_t)c * R + r) : 0.f;
  }
  for (int cc = tid; cc < kBG; cc += kBThreads) {
    const int c = d0 + cc;
    const bool okc = cc < G && c < p.dim;
    s_bias[cc] = (okc && p.delta_bias) ? __ldg(p.delta_bias + c) : 0.f;
    s_D[cc] = (okc && p.D) ? __ldg(p.D + c) : 0.f;
  }

  float A2[kBKP], Dd[kBKP], dDacc[kBKP], dbacc[kBKP], mcar[kBKP], dAacc[kBKP];
  int ch[kBKP];
#pragma unroll
  for (int k = 0; k < kBKP; ++k) {
    const int cl = 2 * (warp + kBW * k) + half;
    const int c = d0 + cl;
    const bool ok = cl < G && c < p.dim;
    ch[k] = ok ? c : -1;
    A2[k] = ok ? __ldg(p.A + (int64_t)c * kN + n) * kLog2e : 0.f;
    Dd[k] = (ok && p.D) ? __ldg(p.D + c) : 0.f;
    dDacc[k] = 0.f;
    dbacc[k] = 0.f;
    mcar[k] = 0.f;
    dAacc[k] = 0.f;
  }
  const int R4 = (R + 3) >> 2;

  for (int c0 = nck - 1; c0 >= 0; --c0) {
    const int bf = c0 & 1;
    const int tau0 = c0 * kT;
    cp_async_wait<0>();
    __syncthreads();  // (1) chunk c0 tiles visible; previous chunk's post-pass done
    if (c0 > 0) stage(c0 - 1, bf ^ 1);

    // ---- pre-pass: one element (step i, channel cc) per thread-iteration
    {
      const T* sx = s_xr + bf * kT * kXW;
      const int valid = 2 * kN + R;
      for (int e = tid; e < kT * kXW; e += kBThreads) {
        const int col = e % kXW;
        s_xf[e] = col < valid ? to_f(sx[e]) : 0.f;
      }
    }
    __syncthreads();  // s_xf ready (dt_r rows are read below)
    {
      const T* sa = s_raw + bf * SM::kMaxRaw * kT * kBG;
      const int cc = tid & (kBG - 1);
      const int c = d0 + cc;
      const bool okc = cc < G && c < p.dim;
      const float bias = s_bias[cc];
#pragma unroll
      for (int j = 0; j < kT * kBG / kBThreads; ++j) {
        const int i = (tid >> 5) + j * (kBThreads / kBG);
        const int tau = tau0 + i;
        float dl = 0.f, dlu = 0.f, gg = 0.f, uu = 0.f, sp = 0.f;
        if (okc && tau < L) {
          const int e = i * kBG + cc;
          uu = to_f(sa[e]);
          const float dov = to_f(sa[kT * kBG + e]);
          float draw = bias;
          if (expl) {
            draw += to_f(sa[4 * kT * kBG + e]);
          } else {
            const float4* xr = reinterpret_cast<const float4*>(s_xf + i * kXW + 2 * kN);
            const float4* wr = reinterpret_cast<const float4*>(s_wdt + cc * BIMAMBA_MAX_DT_RANK);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < R4) {
                const float4 x = xr[q], w = wr[q];
                draw = fmaf(w.x, x.x, draw);
                draw = fmaf(w.y, x.y, draw);
                draw = fmaf(w.z, x.z, draw);
                draw = fmaf(w.w, x.w, draw);
              }
            }
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
            const float zz = to_f(sa[2 * kT * kBG + e]);
            const float sg = sigmoid_f(zz);
            gg = dov * zz * sg;
            if (need_yp) {
              const float yp = to_f(sa[3 * kT * kBG + e]);
              const int64_t t = dir ? (L - 1 - tau) : tau;
              gdz[t * p.out_ts + c] = from_f<T>(dov * yp * sg * (1.f + zz * (1.f - sg)));
            }
          }
        }
        const int o = cc * kDS + i;
        s_dl[o] = dl;
        s_du[o] = dlu;
        s_g[o] = gg;
        s_u[o] = uu;
        s_sp[o] = sp;
      }
    }
    __syncthreads();  // (2) derived arrays ready

    // ---- recurrence
    float Bv[kT], Cv[kT], dBa[kT], dCa[kT];
#pragma unroll
    for (int i = 0; i < kT; ++i) {
      Bv[i] = s_xf[i * kXW + n];
      Cv[i] = s_xf[i * kXW + kN + n];
      dBa[i] = 0.f;
      dCa[i] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < kBKP; ++k) {
      // warp-uniform skip: the pair index is out of range for the whole warp
      if (2 * (warp + kBW * k) >= G || d0 + 2 * (warp + kBW * k) >= p.dim) continue;
      const int c = ch[k];
      const bool ok = c >= 0;
      const int cl = 2 * (warp + kBW * k) + half;
      const float hstart = (ok && c0 > 0) ? p.ckpt[(((bd * nck + c0) * p.dim) + c) * kN + n] : 0.f;
      const float4* pd = reinterpret_cast<const float4*>(s_dl + cl * kDS);
      const float4* pu = reinterpret_cast<const float4*>(s_du + cl * kDS);
      const float4* pg = reinterpret_cast<const float4*>(s_g + cl * kDS);

      // re-run the chunk forward, keeping a[t], h[t]
      float a[kT], hh[kT];
      const float a2 = A2[k];
      {
        float hk = hstart;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 d4 = pd[q], u4 = pu[q];
          const float dq[4] = {d4.x, d4.y, d4.z, d4.w}, uq[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * q + e;
            a[i] = ex2_approx(dq[e] * a2);
            hk = fmaf(a[i], hk, uq[e] * Bv[i]);
            hh[i] = hk;
          }
        }
      }
      // reverse recurrence:  dh_i = g_i C_i + m_{i+1},  m_i = a_i dh_i
      float m = mcar[k];
      float dAl = 0.f;
#pragma unroll
      for (int q = 3; q >= 0; --q) {
        const float4 d4 = pd[q], u4 = pu[q], g4 = pg[q];
        const float dq[4] = {d4.x, d4.y, d4.z, d4.w}, uq[4] = {u4.x, u4.y, u4.z, u4.w},
                    gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int e = 3; e >= 0; --e) {
          const int i = 4 * q + e;
          const float dh = fmaf(gq[e], Cv[i], m);
          m = a[i] * dh;
          const float hp = (i == 0) ? hstart : hh[i - 1];
          const float daa = m * hp;
          dAl = fmaf(daa, dq[e], dAl);
          dBa[i] = fmaf(dh, uq[e], dBa[i]);
          dCa[i] = fmaf(gq[e], hh[i], dCa[i]);
          a[i] = daa * a2;      // a[i] is dead: reuse as the d(delta) partial (x ln2 later)
          hh[i] = dh * Bv[i];   // hh[i] is dead for the remaining steps: the d(delta*u) partial
        }
      }
      mcar[k] = m;
      dAacc[k] += dAl;
      const float rA = reduce_scatter16(a, n);
      const float rU = reduce_scatter16(hh, n);
      // epilogue: lane n owns step n of this channel
      {
        const int o = cl * kDS + n;
        const float uj = s_u[o], dj = s_dl[o], gj = s_g[o], sp = s_sp[o];
        dDacc[k] = fmaf(gj, uj, dDacc[k]);
        const float duv = fmaf(gj, Dd[k], dj * rU);
        const float ddl = fmaf(uj, rU, rA * kLn2) * sp;
        dbacc[k] += ddl;
        s_out[n * kOS + cl] = duv;
        s_out[kT * kOS + n * kOS + cl] = ddl;
      }
    }
    // ---- dB/dC of this chunk: halves by shuffle, warps through shared memory (fixed order)
    {
      float* myRed = s_red + warp * (2 * kT * kN);
#pragma unroll
      for (int i = 0; i < kT; ++i) {
        const float vb = dBa[i] + __shfl_xor_sync(kFull, dBa[i], 16);
        const float vc = dCa[i] + __shfl_xor_sync(kFull, dCa[i], 16);
        myRed[i * 2 * kN + lane] = half ? vc : vb;  // row i: [dB_0..15 | dC_0..15]
      }
    }
    __syncthreads();  // (3) output tiles and per-warp dB/dC tiles complete
#pragma unroll
    for (int j = 0; j < kT * kBG / kBThreads; ++j) {
      const int i = (tid >> 5) + j * (kBThreads / kBG);
      const int cc = tid & (kBG - 1);
      const int tau = tau0 + i;
      if (tau < L) {
        const int64_t t = dir ? (L - 1 - tau) : tau;
        if (cc < G && d0 + cc < p.dim) {
          gdu[t * p.out_ts + d0 + cc] = from_f<T>(s_out[i * kOS + cc]);
          gdd[t * p.out_ts + d0 + cc] = from_f<T>(s_out[kT * kOS + i * kOS + cc]);
        }
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kBW; ++w) s += s_red[w * (2 * kT * kN) + i * 2 * kN + cc];
        partB[t * pb_ts + cc] = s;
      }
    }
    // the next iteration's barrier (1) orders these reads before the tiles are rewritten
  }

#pragma unroll
  for (int k = 0; k < kBKP; ++k) {
    const float sD = half_sum(dDacc[k]);
    const float sb = half_sum(dbacc[k]);
    const int c = ch[k];
    if (c >= 0) {
      p.dA_part[(bd * p.dim + c) * kN + n] = dAacc[k];
      if (n == 0) {
        if (p.dD_part) p.dD_part[bd * p.dim + c] = sD;
        if (p.dbias_part) p.dbias_part[bd * p.dim + c] = sb;
      }
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
  if (G < 2 || G > kBG || (G & 1)) { set_err("backward group_channels must be even, 2..32 (use bimamba_scan_plan)"); return -5; }
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
