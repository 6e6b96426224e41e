// Selective scan (forward + backward), both time directions in one launch.  sm_100a.
//
// Math (reference: src/models/modules/mamba_block.py:80-120 and :61):
//   delta = softplus(delta_raw + bias);  a[t,n] = exp(delta[t] * A[n])
//   h[t,n] = a[t,n] h[t-1,n] + delta[t] B[t,n] u[t];   y[t] = sum_n C[t,n] h[t,n] + D u[t]
//   out[t] = y[t] * silu(z[t])
//
// Mapping (B200-first, not the upstream block-scan):
//   * operands are channel-first (batch, dir, dim, L).  A warp owns TWO channels; lane = 16*c + n
//     holds state n of channel c in a register for the whole sequence (fp32).
//   * time is walked in chunks of T = 16 steps.  For each chunk lane n first acts as the owner
//     of ELEMENT t0+n of its channel: it loads u/delta/z (coalesced), applies softplus / SiLU once
//     per element, and the 16 lanes broadcast their element to each other by warp shuffle
//     while the recurrence runs.  The n-sum  y[t] = sum_n C h  is a 16-lane shuffle
//     reduce-scatter that lands y[t0+j] back on lane j, which gates it and stores it coalesced.
//     So every transcendental other than the 16 decay exps per element is computed once.
//   * B[t,n], C[t,n] of a chunk sit in registers of lane n and are reused by all the channels the
//     warp handles (chunk-outer, channel-inner loop).
//   * direction 1 walks the same storage back to front (t = L-1-step): flip(M(flip(x))) of
//     src/models/DualStreamSEMamba.py:476-478 with no flipped copy; both directions are grid.y
//     of the same launch.
//   * training forward also writes the fp32 state entering every chunk ("checkpoints") and the
//     pre-gate y.  Backward walks the chunks last to first: it re-runs the 16 steps of a chunk
//     from its checkpoint keeping a[t], h[t] in registers, then runs the reverse recurrence for
//     dh over the same registers - no (B, L, D, N) tensor, 16 exps per element in each pass.
//     dB/dC accumulate in registers over the warp's channels, are combined over the CTA's warps
//     in fixed order through shared memory and written as per-group partials; dA/dD/dbias are
//     per-(batch, dir, channel) partials.  All cross-CTA sums are done in fixed order by
//     bimamba_reduce_partials: the whole backward is deterministic (no atomics).
#include "common.cuh"

namespace bimamba {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kT = 16;       // steps per chunk == checkpoint interval
constexpr int kN = 16;       // d_state handled by these kernels
constexpr int kMaxG = 32;    // channels per CTA
constexpr int kMaxKP = kMaxG / (2 * kWarps);  // channel pairs per warp

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

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) scan_fwd_kernel(const bimamba_scan_desc p) {
  const int b = blockIdx.z, dir = blockIdx.y, G = p.group_channels, d0 = blockIdx.x * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, n = lane & 15;
  const int L = p.seqlen, dt = p.io_dtype;
  const int nck = (L + kT - 1) / kT;
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;
  const int64_t bd = (int64_t)b * p.ndir + dir;
  const int64_t bcb = (int64_t)b * p.bc_bs + (int64_t)dir * p.bc_ds + (int64_t)n * p.bc_rs;

  float h[kMaxKP], A2[kMaxKP], bias[kMaxKP], Dd[kMaxKP];
  int ch[kMaxKP];
#pragma unroll
  for (int k = 0; k < kMaxKP; ++k) {
    const int cl = 2 * (warp + kWarps * k) + half;
    const int c = d0 + cl;
    const bool ok = cl < G && c < p.dim;
    ch[k] = ok ? c : -1;
    h[k] = 0.f;
    A2[k] = ok ? __ldg(p.A + (int64_t)c * kN + n) * kLog2e : 0.f;
    bias[k] = (ok && p.delta_bias) ? __ldg(p.delta_bias + c) : 0.f;
    Dd[k] = (ok && p.D) ? __ldg(p.D + c) : 0.f;
  }

  for (int c0 = 0; c0 < nck; ++c0) {
    const int tau0 = c0 * kT;
    float Bv[kT], Cv[kT];
#pragma unroll
    for (int i = 0; i < kT; ++i) {
      const int tau = tau0 + i;
      Bv[i] = 0.f;
      Cv[i] = 0.f;
      if (tau < L) {
        const int t = dir ? (L - 1 - tau) : tau;
        Bv[i] = ld_f(p.Bm, bcb + t, p.bc_dtype);
        Cv[i] = ld_f(p.Cm, bcb + t, p.bc_dtype);
      }
    }
    const int tauj = tau0 + n;  // the element this lane owns in this chunk
    const int tj = dir ? (L - 1 - tauj) : tauj;

#pragma unroll
    for (int k = 0; k < kMaxKP; ++k) {
      // warp-uniform skip: the pair index is out of range for the whole warp
      if (2 * (warp + kWarps * k) >= G || d0 + 2 * (warp + kWarps * k) >= p.dim) continue;
      const int c = ch[k];
      const bool ok = c >= 0;
      const int cc = ok ? c : 0;
      const int64_t ub = (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds + (int64_t)cc * p.u_rs;
      const int64_t db = (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds + (int64_t)cc * p.delta_rs;
      const int64_t zb = (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds + (int64_t)cc * p.z_rs;
      const int64_t ob = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds + (int64_t)cc * p.out_rs;

      if (p.ckpt && ok) p.ckpt[((bd * p.dim + c) * nck + c0) * kN + n] = h[k];

      const bool live = ok && tauj < L;
      float uj = 0.f, dj = 0.f, zj = 0.f;
      if (live) {
        uj = ld_f(p.u, ub + tj, dt);
        dj = ld_f(p.delta, db + tj, dt) + bias[k];
        if (softplus) dj = softplus_f(dj);
        if (p.z) zj = ld_f(p.z, zb + tj, dt);
      }
      const float duj = dj * uj;

      float pr[kT];
      float hk = h[k];
      const float a2 = A2[k];
#pragma unroll
      for (int i = 0; i < kT; ++i) {
        const float di = __shfl_sync(kFull, dj, i, 16);
        const float dui = __shfl_sync(kFull, duj, i, 16);
        const float a = ex2_approx(di * a2);
        hk = fmaf(a, hk, dui * Bv[i]);
        pr[i] = Cv[i] * hk;
      }
      h[k] = hk;
      float y = reduce_scatter16(pr, n);
      y = fmaf(Dd[k], uj, y);
      if (live) {
        if (p.ypre) st_f(p.ypre, ob + tj, y, dt);
        if (p.z) y *= zj * sigmoid_f(zj);
        st_f(p.out, ob + tj, y, dt);
      }
      if (c0 == 0 && ok) {
        for (int t = L + n; t < p.pad_to; t += 16) {
          st_f(p.out, ob + t, 0.f, dt);
          if (p.ypre) st_f(p.ypre, ob + t, 0.f, dt);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) scan_bwd_kernel(const bimamba_scan_desc p) {
  __shared__ float sCarry[kMaxG * kN];             // m = a*dh flowing to earlier steps, per (channel, n)
  __shared__ float sdA[kMaxG * kN];
  __shared__ float sRed[kWarps * 2 * kT * kN];     // per-warp dB/dC chunk tiles

  const int b = blockIdx.z, dir = blockIdx.y, G = p.group_channels, g = blockIdx.x, d0 = g * G;
  const int ngroups = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, n = lane & 15;
  const int L = p.seqlen, dt = p.io_dtype;
  const int nck = (L + kT - 1) / kT;
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;
  const bool gated = p.z != nullptr;
  const int64_t bd = (int64_t)b * p.ndir + dir;
  const int64_t bcb = (int64_t)b * p.bc_bs + (int64_t)dir * p.bc_ds + (int64_t)n * p.bc_rs;

  for (int i = threadIdx.x; i < kMaxG * kN; i += kThreads) {
    sCarry[i] = 0.f;
    sdA[i] = 0.f;
  }
  float* partB = p.dBC_part + ((bd * ngroups + g) * 2) * (int64_t)kN * p.dbc_rs;
  {  // zero the padding columns so the ordered reduction can run over whole rows
    const int padw = (int)(p.dbc_rs - L);
    for (int idx = threadIdx.x; idx < 2 * kN * padw; idx += kThreads)
      partB[(int64_t)(idx / padw) * p.dbc_rs + L + (idx % padw)] = 0.f;
  }

  float A2[kMaxKP], bias[kMaxKP], Dd[kMaxKP], dDacc[kMaxKP], dbacc[kMaxKP];
  int ch[kMaxKP];
#pragma unroll
  for (int k = 0; k < kMaxKP; ++k) {
    const int cl = 2 * (warp + kWarps * k) + half;
    const int c = d0 + cl;
    const bool ok = cl < G && c < p.dim;
    ch[k] = ok ? c : -1;
    A2[k] = ok ? __ldg(p.A + (int64_t)c * kN + n) * kLog2e : 0.f;
    bias[k] = (ok && p.delta_bias) ? __ldg(p.delta_bias + c) : 0.f;
    Dd[k] = (ok && p.D) ? __ldg(p.D + c) : 0.f;
    dDacc[k] = 0.f;
    dbacc[k] = 0.f;
  }
  __syncthreads();

  for (int c0 = nck - 1; c0 >= 0; --c0) {
    const int tau0 = c0 * kT;
    float Bv[kT], Cv[kT], dBa[kT], dCa[kT];
#pragma unroll
    for (int i = 0; i < kT; ++i) {
      const int tau = tau0 + i;
      Bv[i] = 0.f;
      Cv[i] = 0.f;
      dBa[i] = 0.f;
      dCa[i] = 0.f;
      if (tau < L) {
        const int t = dir ? (L - 1 - tau) : tau;
        Bv[i] = ld_f(p.Bm, bcb + t, p.bc_dtype);
        Cv[i] = ld_f(p.Cm, bcb + t, p.bc_dtype);
      }
    }
    const int tauj = tau0 + n;
    const int tj = dir ? (L - 1 - tauj) : tauj;

#pragma unroll
    for (int k = 0; k < kMaxKP; ++k) {
      if (2 * (warp + kWarps * k) >= G || d0 + 2 * (warp + kWarps * k) >= p.dim) continue;
      const int c = ch[k];
      const bool ok = c >= 0;
      const int cc = ok ? c : 0;
      const int cl = 2 * (warp + kWarps * k) + half;
      const int64_t ub = (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds + (int64_t)cc * p.u_rs;
      const int64_t db = (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds + (int64_t)cc * p.delta_rs;
      const int64_t zb = (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds + (int64_t)cc * p.z_rs;
      const int64_t ob = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds + (int64_t)cc * p.out_rs;
      const int64_t dzb = (int64_t)b * p.dz_bs + (int64_t)dir * p.dz_ds + (int64_t)cc * p.dz_rs;
      const int64_t yb = (int64_t)b * p.ypre_bs + (int64_t)dir * p.ypre_ds + (int64_t)cc * p.ypre_rs;

      const bool live = ok && tauj < L;
      float uj = 0.f, dj = 0.f, zj = 0.f, doj = 0.f, sgj = 0.f, gj = 0.f;
      if (live) {
        uj = ld_f(p.u, ub + tj, dt);
        dj = ld_f(p.delta, db + tj, dt) + bias[k];
        if (softplus) dj = softplus_f(dj);
        doj = ld_f(p.dout, ob + tj, dt);
        gj = doj;
        if (gated) {
          zj = ld_f(p.z, zb + tj, dt);
          sgj = sigmoid_f(zj);
          gj = doj * zj * sgj;
        }
      }
      const float duj = dj * uj;
      const float hstart = (ok && c0 > 0) ? p.ckpt[((bd * p.dim + c) * nck + c0) * kN + n] : 0.f;

      // ---- re-run the chunk forward, keeping a[t], h[t] ----
      float a[kT], hh[kT];
      const float a2 = A2[k];
      {
        float hk = hstart;
#pragma unroll
        for (int i = 0; i < kT; ++i) {
          const float di = __shfl_sync(kFull, dj, i, 16);
          const float dui = __shfl_sync(kFull, duj, i, 16);
          a[i] = ex2_approx(di * a2);
          hk = fmaf(a[i], hk, dui * Bv[i]);
          hh[i] = hk;
        }
      }
      // ---- reverse recurrence:  dh_i = g_i C_i + m_{i+1},  m_i = a_i dh_i ----
      float m = sCarry[cl * kN + n];
      float dAl = 0.f;
#pragma unroll
      for (int i = kT - 1; i >= 0; --i) {
        const float gi = __shfl_sync(kFull, gj, i, 16);
        const float di = __shfl_sync(kFull, dj, i, 16);
        const float dui = __shfl_sync(kFull, duj, i, 16);
        const float dh = fmaf(gi, Cv[i], m);
        m = a[i] * dh;
        const float hp = (i == 0) ? hstart : hh[i - 1];
        const float daa = m * hp;
        dAl = fmaf(daa, di, dAl);
        dBa[i] = fmaf(dh, dui, dBa[i]);
        dCa[i] = fmaf(gi, hh[i], dCa[i]);
        a[i] = daa * a2;       // a[i] is dead: reuse as the d(delta) partial (x ln2 later)
        hh[i] = dh * Bv[i];    // hh[i] is dead for the remaining steps: reuse as the d(delta*u) partial
      }
      sCarry[cl * kN + n] = m;
      sdA[cl * kN + n] += dAl;
      const float rA = reduce_scatter16(a, n);
      const float rU = reduce_scatter16(hh, n);

      if (live) {
        dDacc[k] = fmaf(gj, uj, dDacc[k]);
        const float duv = fmaf(gj, Dd[k], dj * rU);
        float ddl = fmaf(uj, rU, rA * kLn2);
        if (softplus) ddl *= (1.f - expf(-dj));  // sigmoid(raw) == 1 - exp(-softplus(raw))
        dbacc[k] += ddl;
        st_f(p.du, ub + tj, duv, dt);
        st_f(p.ddelta, db + tj, ddl, dt);
        if (gated && p.dz) {
          const float yj = ld_f(p.ypre, yb + tj, dt);
          st_f(p.dz, dzb + tj, doj * yj * sgj * (1.f + zj * (1.f - sgj)), dt);
        }
      }
      if (c0 == 0 && ok) {
        for (int t = L + n; t < p.pad_to; t += 16) {
          st_f(p.du, ub + t, 0.f, dt);
          st_f(p.ddelta, db + t, 0.f, dt);
          if (gated && p.dz) st_f(p.dz, dzb + t, 0.f, dt);
        }
      }
    }

    // ---- combine dB/dC of this chunk: halves by shuffle, warps through shared memory (fixed order) ----
    float* myRed = sRed + warp * (2 * kT * kN);
#pragma unroll
    for (int i = 0; i < kT; ++i) {
      const float vb = dBa[i] + __shfl_xor_sync(kFull, dBa[i], 16);
      const float vc = dCa[i] + __shfl_xor_sync(kFull, dCa[i], 16);
      if (half == 0) {
        myRed[n * kT + i] = vb;  // [row n][step i]
        myRed[kT * kN + n * kT + i] = vc;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * kN * kT; idx += kThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += sRed[w * (2 * kT * kN) + idx];
      const int row = idx / kT;  // 0..15 dB rows, 16..31 dC rows
      const int tau = tau0 + (idx % kT);
      if (tau < L) {
        const int t = dir ? (L - 1 - tau) : tau;
        partB[(int64_t)row * p.dbc_rs + t] = s;
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int k = 0; k < kMaxKP; ++k) {
    const float sD = half_sum(dDacc[k]);
    const float sb = half_sum(dbacc[k]);
    const int c = ch[k];
    if (c >= 0) {
      const int cl = 2 * (warp + kWarps * k) + half;
      p.dA_part[(bd * p.dim + c) * kN + n] = sdA[cl * kN + n];
      if (n == 0) {
        if (p.dD_part) p.dD_part[bd * p.dim + c] = sD;
        if (p.dbias_part) p.dbias_part[bd * p.dim + c] = sb;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
thread_local char g_err[512] = "";
void set_err(const char* msg) {
  size_t i = 0;
  for (; msg[i] && i + 1 < sizeof(g_err); ++i) g_err[i] = msg[i];
  g_err[i] = 0;
}

int check_desc(const bimamba_scan_desc* d, bool bwd) {
  if (!d) { set_err("null descriptor"); return -1; }
  if (d->dstate != kN) { set_err("dstate must be 16"); return -2; }
  if (d->batch < 0 || d->ndir < 1 || d->ndir > 2 || d->dim < 1 || d->seqlen < 0) { set_err("bad sizes"); return -3; }
  if (d->batch > 65535) { set_err("batch > 65535 not supported by this launch geometry"); return -3; }
  if (d->chunk_items != kT) { set_err("chunk_items must be 16 (use bimamba_scan_plan)"); return -4; }
  if (d->group_channels < 2 || d->group_channels > kMaxG || (d->group_channels & 1)) { set_err("group_channels must be even, 2..32"); return -5; }
  if (d->io_dtype < 0 || d->io_dtype > 2 || d->bc_dtype < 0 || d->bc_dtype > 2) { set_err("bad dtype"); return -6; }
  if (!d->u || !d->delta || !d->A || !d->Bm || !d->Cm) { set_err("null operand"); return -7; }
  if (!bwd && !d->out) { set_err("null out"); return -7; }
  if (bwd) {
    if (!d->dout || !d->du || !d->ddelta || !d->dBC_part || !d->dA_part) { set_err("null backward operand"); return -8; }
    if (d->seqlen > kT && !d->ckpt) { set_err("backward needs the forward checkpoints"); return -9; }
    if (d->dbc_rs < d->seqlen) { set_err("dbc_rs < seqlen"); return -10; }
    if (d->z && d->dz && !d->ypre) { set_err("gated backward needs ypre saved by the forward"); return -11; }
  }
  return 0;
}

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_abi_version(void) { return BIMAMBA_ABI_VERSION; }
extern "C" const char* bimamba_last_error(void) { return g_err; }

extern "C" int bimamba_scan_plan(int seqlen, int dim, int rows, int backward, int* chunk_items, int* group_channels) {
  (void)backward;
  int G = kMaxG;
  // keep at least ~3 CTAs per SM worth of blocks when the batch is small
  while (G > 2 * kWarps && (int64_t)rows * ((dim + G - 1) / G) < 3 * 148) G /= 2;
  if (chunk_items) *chunk_items = kT;
  if (group_channels) *group_channels = G;
  return seqlen > 0 ? (seqlen + kT - 1) / kT : 1;
}

extern "C" int bimamba_selective_scan_fwd(const bimamba_scan_desc* d, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;  // empty: nothing to do (pointers may be null)
  int rc = check_desc(d, false);
  if (rc) return rc;
  dim3 grid((d->dim + d->group_channels - 1) / d->group_channels, d->ndir, d->batch);
  scan_fwd_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int bimamba_selective_scan_bwd(const bimamba_scan_desc* d, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;
  int rc = check_desc(d, true);
  if (rc) return rc;
  dim3 grid((d->dim + d->group_channels - 1) / d->group_channels, d->ndir, d->batch);
  scan_bwd_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}
