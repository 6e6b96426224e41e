// Selective scan (forward + backward), both time directions in one launch.  sm_100a.
//
// Math (reference: src/models/modules/mamba_block.py:80-120 and :61):
//   delta = softplus(delta_raw + bias);  a[t,n] = exp(delta[t] * A[n])
//   h[t,n] = a[t,n] h[t-1,n] + delta[t] B[t,n] u[t];   y[t] = sum_n C[t,n] h[t,n] + D u[t]
//   out[t] = y[t] * silu(z[t])
//
// Mapping (B200-first, not the upstream block-scan):
//   * operands are channel-first (batch, dir, dim, L): one warp owns one (batch, dir, channel)
//     row at a time; lane l owns a STRIP of I consecutive scan steps, so a chunk is 32*I steps.
//   * per state n the lane runs its strip recurrence from zero, the 32 strip summaries
//     (P = prod a, H = strip-end state) are composed with a 5-step warp-shuffle scan of the
//     affine maps (P2,H2)o(P1,H1) = (P2 P1, P2 H1 + H2), and the strip is re-run from the
//     correct incoming state.  The 16 exps per element are computed ONCE and kept in registers
//     between the two passes - no recompute, no (B,L,D,N) tensor.
//   * chunk-to-chunk carries live in registers (lane n keeps state n of each of its channels).
//   * B[t,:], C[t,:] of the chunk are staged once per CTA in shared memory (they are shared by
//     every channel of the sample); strips are padded to an odd stride so lane-strided reads
//     are bank-conflict free.
//   * direction 1 walks the same storage back to front (t = L-1-step): flip(M(flip(x))) of
//     src/models/DualStreamSEMamba.py:476-478 with no flipped copy; both directions are grid.y
//     of the same launch.
//   * backward: forward states are recomputed per chunk from the fp32 chunk-boundary
//     checkpoints written by the forward kernel; dh runs as the mirrored (shfl_down) scan with
//     the same a[] registers; dB/dC are reduced over the CTA's channels in shared memory and
//     written as per-group partials; dA/dD/dbias as per-(batch,dir,channel) partials.  The
//     cross-CTA sums are done in fixed order by bimamba_reduce_partials.
#include "common.cuh"

namespace bimamba {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxCpw = 4;  // channels per warp (group_channels <= kMaxCpw * kWarps)
constexpr int kN = 16;      // d_state handled by these kernels

template <int I>
struct Geo {
  static constexpr int IS = I | 1;     // odd strip stride in shared memory
  static constexpr int ROW = 32 * IS;  // floats per staged row
  static constexpr int TC = 32 * I;    // scan steps per chunk
};

// position of chunk-local step tau in a staged row
template <int I>
__device__ __forceinline__ int spos(int tau) {
  return (tau / I) * Geo<I>::IS + (tau % I);
}

template <int I>
__device__ __forceinline__ void stage_bc(float* __restrict__ sB, float* __restrict__ sC,
                                         const bimamba_scan_desc& p, int b, int dir, int chunk) {
  constexpr int TC = Geo<I>::TC, ROW = Geo<I>::ROW;
  const int L = p.seqlen;
  const int64_t base = (int64_t)b * p.bc_bs + (int64_t)dir * p.bc_ds;
  for (int idx = threadIdx.x; idx < 2 * kN * TC; idx += kThreads) {
    const int row = idx / TC;
    const int tau = idx - row * TC;
    const int tg = chunk * TC + tau;
    const int n = row & (kN - 1);
    float v = 0.f;
    if (tg < L) {
      const int t = dir ? (L - 1 - tg) : tg;
      v = ld_f(row < kN ? p.Bm : p.Cm, base + (int64_t)n * p.bc_rs + t, p.bc_dtype);
    }
    (row < kN ? sB : sC)[n * ROW + spos<I>(tau)] = v;
  }
}

// inclusive scan over lanes of the affine maps (P, H), composing left-to-right (lower lanes first)
__device__ __forceinline__ void warp_scan_up(float& P, float& H, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float Pp = __shfl_up_sync(kFull, P, off);
    const float Hp = __shfl_up_sync(kFull, H, off);
    if (lane >= off) {
      H = fmaf(P, Hp, H);
      P *= Pp;
    }
  }
}
// mirrored: composes right-to-left (higher lanes first)
__device__ __forceinline__ void warp_scan_down(float& P, float& H, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float Pp = __shfl_down_sync(kFull, P, off);
    const float Hp = __shfl_down_sync(kFull, H, off);
    if (lane + off < 32) {
      H = fmaf(P, Hp, H);
      P *= Pp;
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int I, bool MULTI>
__global__ void __launch_bounds__(kThreads) scan_fwd_kernel(const bimamba_scan_desc p) {
  constexpr int IS = Geo<I>::IS, ROW = Geo<I>::ROW, TC = Geo<I>::TC;
  extern __shared__ float smem[];
  float* sB = smem;
  float* sC = sB + kN * ROW;
  float* sA = sC + kN * ROW;  // [group_channels][16], A * log2(e)

  const int b = blockIdx.z, dir = blockIdx.y, G = p.group_channels, d0 = blockIdx.x * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.seqlen, dt = p.io_dtype;
  const int nchunks = MULTI ? (L + TC - 1) / TC : 1;
  const int gch = min(G, p.dim - d0);
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;

  for (int idx = threadIdx.x; idx < gch * kN; idx += kThreads) sA[idx] = p.A[(int64_t)d0 * kN + idx] * kLog2e;

  float carry[kMaxCpw];
#pragma unroll
  for (int k = 0; k < kMaxCpw; ++k) carry[k] = 0.f;

  for (int c = 0; c < nchunks; ++c) {
    __syncthreads();
    stage_bc<I>(sB, sC, p, b, dir, c);
    __syncthreads();

#pragma unroll
    for (int k = 0; k < kMaxCpw; ++k) {
      const int cl = warp + k * kWarps;
      if (cl < gch) {
        const int d = d0 + cl;
        const int64_t ub = (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds + (int64_t)d * p.u_rs;
        const int64_t db = (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds + (int64_t)d * p.delta_rs;
        const int64_t zb = (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds + (int64_t)d * p.z_rs;
        const int64_t ob = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds + (int64_t)d * p.out_rs;
        const float bias = p.delta_bias ? __ldg(p.delta_bias + d) : 0.f;
        const float Dd = p.D ? __ldg(p.D + d) : 0.f;

        if (MULTI && p.ckpt && lane < kN)
          p.ckpt[((((int64_t)b * p.ndir + dir) * p.dim + d) * nchunks + c) * kN + lane] = carry[k];
        if (c == 0)
          for (int t = L + lane; t < p.pad_to; t += 32) st_f(p.out, ob + t, 0.f, dt);

        float dl[I], dlu[I], y[I], zv[I];
        float sumd = 0.f;
        const int tg0 = c * TC + lane * I;
#pragma unroll
        for (int i = 0; i < I; ++i) {
          const int tg = tg0 + i;
          float uv = 0.f, dv = 0.f;
          zv[i] = 0.f;
          if (tg < L) {
            const int t = dir ? (L - 1 - tg) : tg;
            uv = ld_f(p.u, ub + t, dt);
            dv = ld_f(p.delta, db + t, dt) + bias;
            if (softplus) dv = softplus_f(dv);
            if (p.z) zv[i] = ld_f(p.z, zb + t, dt);
          }
          dl[i] = dv;
          dlu[i] = dv * uv;
          y[i] = Dd * uv;
          sumd += dv;
        }

        const float* sAd = sA + cl * kN;
        const float* sBl = sB + lane * IS;
        const float* sCl = sC + lane * IS;
#pragma unroll 2
        for (int n = 0; n < kN; ++n) {
          const float A2 = sAd[n];
          float a[I], bb[I];
          float H = 0.f;
#pragma unroll
          for (int i = 0; i < I; ++i) {
            a[i] = ex2_approx(dl[i] * A2);
            bb[i] = dlu[i] * sBl[n * ROW + i];
            H = fmaf(a[i], H, bb[i]);
          }
          float P = ex2_approx(sumd * A2);
          warp_scan_up(P, H, lane);
          float Hin = __shfl_up_sync(kFull, H, 1);
          float Pin = __shfl_up_sync(kFull, P, 1);
          if (lane == 0) {
            Hin = 0.f;
            Pin = 1.f;
          }
          float h = Hin;
          if (MULTI) {
            const float cin = __shfl_sync(kFull, carry[k], n);
            h = fmaf(Pin, cin, Hin);
            const float P31 = __shfl_sync(kFull, P, 31);
            const float H31 = __shfl_sync(kFull, H, 31);
            if (lane == n) carry[k] = fmaf(P31, cin, H31);
          }
#pragma unroll
          for (int i = 0; i < I; ++i) {
            h = fmaf(a[i], h, bb[i]);
            y[i] = fmaf(sCl[n * ROW + i], h, y[i]);
          }
        }

#pragma unroll
        for (int i = 0; i < I; ++i) {
          const int tg = tg0 + i;
          if (tg < L) {
            const int t = dir ? (L - 1 - tg) : tg;
            float o = y[i];
            if (p.z) o *= zv[i] * sigmoid_f(zv[i]);
            st_f(p.out, ob + t, o, dt);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int I, bool MULTI>
__global__ void __launch_bounds__(kThreads) scan_bwd_kernel(const bimamba_scan_desc p) {
  constexpr int IS = Geo<I>::IS, ROW = Geo<I>::ROW, TC = Geo<I>::TC;
  extern __shared__ float smem[];
  float* sB = smem;
  float* sC = sB + kN * ROW;
  float* sdB = sC + kN * ROW;
  float* sdC = sdB + kN * ROW;
  float* sA = sdC + kN * ROW;              // [G][16]
  float* sRed = sA + p.group_channels * kN;  // [kWarps][16][33]

  const int b = blockIdx.z, dir = blockIdx.y, G = p.group_channels, g = blockIdx.x, d0 = g * G;
  const int ngroups = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.seqlen, dt = p.io_dtype;
  const int nchunks = MULTI ? (L + TC - 1) / TC : 1;
  const int gch = min(G, p.dim - d0);
  const bool softplus = (p.flags & BIMAMBA_FLAG_SOFTPLUS) != 0;
  const int64_t bd = (int64_t)b * p.ndir + dir;
  float* myRed = sRed + warp * (kN * 33);

  for (int idx = threadIdx.x; idx < gch * kN; idx += kThreads) sA[idx] = p.A[(int64_t)d0 * kN + idx] * kLog2e;

  float* partB = p.dBC_part + ((bd * ngroups + g) * 2) * (int64_t)kN * p.dbc_rs;
  {  // zero the padding columns so the ordered reduction can run over whole rows
    const int padw = (int)(p.dbc_rs - L);
    for (int idx = threadIdx.x; idx < 2 * kN * padw; idx += kThreads)
      partB[(int64_t)(idx / padw) * p.dbc_rs + L + (idx % padw)] = 0.f;
  }

  float carryR[kMaxCpw], dAacc[kMaxCpw], dDacc[kMaxCpw], dbacc[kMaxCpw];
#pragma unroll
  for (int k = 0; k < kMaxCpw; ++k) carryR[k] = dAacc[k] = dDacc[k] = dbacc[k] = 0.f;

  for (int c = nchunks - 1; c >= 0; --c) {
    __syncthreads();
    stage_bc<I>(sB, sC, p, b, dir, c);
    for (int idx = threadIdx.x; idx < 2 * kN * ROW; idx += kThreads) sdB[idx] = 0.f;  // sdB and sdC are adjacent
    __syncthreads();

#pragma unroll
    for (int k = 0; k < kMaxCpw; ++k) {
      const int cl = warp + k * kWarps;
      if (cl < gch) {
        const int d = d0 + cl;
        const int64_t ub = (int64_t)b * p.u_bs + (int64_t)dir * p.u_ds + (int64_t)d * p.u_rs;
        const int64_t db = (int64_t)b * p.delta_bs + (int64_t)dir * p.delta_ds + (int64_t)d * p.delta_rs;
        const int64_t zb = (int64_t)b * p.z_bs + (int64_t)dir * p.z_ds + (int64_t)d * p.z_rs;
        const int64_t ob = (int64_t)b * p.out_bs + (int64_t)dir * p.out_ds + (int64_t)d * p.out_rs;
        const int64_t dzb = (int64_t)b * p.dz_bs + (int64_t)dir * p.dz_ds + (int64_t)d * p.dz_rs;
        const float bias = p.delta_bias ? __ldg(p.delta_bias + d) : 0.f;
        const float Dd = p.D ? __ldg(p.D + d) : 0.f;

        if (c == 0) {
          for (int t = L + lane; t < p.pad_to; t += 32) {
            st_f(p.du, ub + t, 0.f, dt);
            st_f(p.ddelta, db + t, 0.f, dt);
            if (p.z && p.dz) st_f(p.dz, dzb + t, 0.f, dt);
          }
        }
        float cfw = 0.f;  // lane n: forward state n entering this chunk
        if (MULTI && c > 0 && lane < kN) cfw = p.ckpt[(((bd * p.dim + d) * nchunks) + c) * kN + lane];

        float uu[I], dl[I], dlu[I], gg[I], y[I], zv[I], dov[I], ddA[I], ddu[I];
        float sumd = 0.f;
        const int tg0 = c * TC + lane * I;
#pragma unroll
        for (int i = 0; i < I; ++i) {
          const int tg = tg0 + i;
          float uv = 0.f, dv = 0.f, zz = 0.f, dy = 0.f;
          if (tg < L) {
            const int t = dir ? (L - 1 - tg) : tg;
            uv = ld_f(p.u, ub + t, dt);
            dv = ld_f(p.delta, db + t, dt) + bias;
            if (softplus) dv = softplus_f(dv);
            dy = ld_f(p.dout, ob + t, dt);
            if (p.z) zz = ld_f(p.z, zb + t, dt);
          }
          uu[i] = uv;
          dl[i] = dv;
          dlu[i] = dv * uv;
          zv[i] = zz;
          dov[i] = dy;
          gg[i] = p.z ? dy * zz * sigmoid_f(zz) : dy;
          y[i] = Dd * uv;
          ddA[i] = 0.f;
          ddu[i] = 0.f;
          sumd += dv;
        }

        const float* sAd = sA + cl * kN;
        const int so = lane * IS;
#pragma unroll 1
        for (int n = 0; n < kN; ++n) {
          const float A2 = sAd[n];
          const float* sBn = sB + n * ROW + so;
          const float* sCn = sC + n * ROW + so;
          float a[I], h[I], Bv[I], Cv[I];
          float H = 0.f;
#pragma unroll
          for (int i = 0; i < I; ++i) {
            Bv[i] = sBn[i];
            Cv[i] = sCn[i];
            a[i] = ex2_approx(dl[i] * A2);
            h[i] = dlu[i] * Bv[i];  // holds b[i] until pass 2
            H = fmaf(a[i], H, h[i]);
          }
          const float Pstrip = ex2_approx(sumd * A2);
          float P = Pstrip;
          warp_scan_up(P, H, lane);
          float Hin = __shfl_up_sync(kFull, H, 1);
          float Pin = __shfl_up_sync(kFull, P, 1);
          if (lane == 0) {
            Hin = 0.f;
            Pin = 1.f;
          }
          float hin = Hin;
          if (MULTI) {
            const float cin = __shfl_sync(kFull, cfw, n);
            hin = fmaf(Pin, cin, Hin);
          }
          {
            float hh = hin;
#pragma unroll
            for (int i = 0; i < I; ++i) {
              hh = fmaf(a[i], hh, h[i]);
              h[i] = hh;
              y[i] = fmaf(Cv[i], hh, y[i]);
            }
          }
          // ---- reverse: m_i = a_i * dh_i, dh_i = g_i C_i + m_{i+1} ----
          float M = 0.f;
#pragma unroll
          for (int i = I - 1; i >= 0; --i) M = a[i] * fmaf(gg[i], Cv[i], M);
          float Q = Pstrip;
          warp_scan_down(Q, M, lane);
          float Min = __shfl_down_sync(kFull, M, 1);
          float Qin = __shfl_down_sync(kFull, Q, 1);
          if (lane == 31) {
            Min = 0.f;
            Qin = 1.f;
          }
          float m = Min;
          if (MULTI) {
            const float rin = __shfl_sync(kFull, carryR[k], n);
            m = fmaf(Qin, rin, Min);
            const float Q0 = __shfl_sync(kFull, Q, 0);
            const float M0 = __shfl_sync(kFull, M, 0);
            if (lane == n) carryR[k] = fmaf(Q0, rin, M0);
          }
          float dAl = 0.f;
#pragma unroll
          for (int i = I - 1; i >= 0; --i) {
            const float dh = fmaf(gg[i], Cv[i], m);
            m = a[i] * dh;
            const float hp = (i == 0) ? hin : h[i - 1];
            const float daa = m * hp;
            dAl = fmaf(daa, dl[i], dAl);
            ddA[i] = fmaf(daa, A2, ddA[i]);
            ddu[i] = fmaf(dh, Bv[i], ddu[i]);
            atomicAdd(sdB + n * ROW + so + i, dh * dlu[i]);
            atomicAdd(sdC + n * ROW + so + i, gg[i] * h[i]);
          }
          myRed[n * 33 + lane] = dAl;
        }
        __syncwarp();
        if (lane < kN) {
          float s = 0.f;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) s += myRed[lane * 33 + j];
          dAacc[k] += s;
        }
        __syncwarp();

        float dDl = 0.f, dbl = 0.f;
#pragma unroll
        for (int i = 0; i < I; ++i) {
          const int tg = tg0 + i;
          if (tg < L) {
            const int t = dir ? (L - 1 - tg) : tg;
            dDl = fmaf(gg[i], uu[i], dDl);
            const float duv = fmaf(gg[i], Dd, dl[i] * ddu[i]);
            float ddl = fmaf(uu[i], ddu[i], ddA[i] * kLn2);
            if (softplus) ddl *= (1.f - expf(-dl[i]));  // sigmoid(raw) == 1 - exp(-softplus(raw))
            dbl += ddl;
            st_f(p.du, ub + t, duv, dt);
            st_f(p.ddelta, db + t, ddl, dt);
            if (p.z && p.dz) {
              const float sg = sigmoid_f(zv[i]);
              st_f(p.dz, dzb + t, dov[i] * y[i] * sg * (1.f + zv[i] * (1.f - sg)), dt);
            }
          }
        }
        dDacc[k] += warp_sum(dDl);
        dbacc[k] += warp_sum(dbl);
      }
    }

    __syncthreads();
    // write this chunk's dB / dC partial tile (natural time order)
    for (int idx = threadIdx.x; idx < 2 * kN * TC; idx += kThreads) {
      const int row = idx / TC;
      const int tau = idx - row * TC;
      const int tg = c * TC + tau;
      if (tg < L) {
        const int t = dir ? (L - 1 - tg) : tg;
        partB[(int64_t)row * p.dbc_rs + t] = sdB[row * ROW + spos<I>(tau)];
      }
    }
  }

#pragma unroll
  for (int k = 0; k < kMaxCpw; ++k) {
    const int cl = warp + k * kWarps;
    if (cl < gch) {
      const int d = d0 + cl;
      if (lane < kN) p.dA_part[(bd * p.dim + d) * kN + lane] = dAacc[k];
      if (lane == 0) {
        if (p.dD_part) p.dD_part[bd * p.dim + d] = dDacc[k];
        if (p.dbias_part) p.dbias_part[bd * p.dim + d] = dbacc[k];
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

template <int I>
size_t fwd_smem(int G) { return sizeof(float) * (2 * kN * Geo<I>::ROW + G * kN); }
template <int I>
size_t bwd_smem(int G) { return sizeof(float) * (4 * kN * Geo<I>::ROW + G * kN + kWarps * kN * 33); }

template <int I, bool MULTI>
int launch_fwd(const bimamba_scan_desc& d, cudaStream_t st) {
  const size_t smem = fwd_smem<I>(d.group_channels);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(scan_fwd_kernel<I, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  dim3 grid((d.dim + d.group_channels - 1) / d.group_channels, d.ndir, d.batch);
  scan_fwd_kernel<I, MULTI><<<grid, kThreads, smem, st>>>(d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

template <int I, bool MULTI>
int launch_bwd(const bimamba_scan_desc& d, cudaStream_t st) {
  const size_t smem = bwd_smem<I>(d.group_channels);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(scan_bwd_kernel<I, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  dim3 grid((d.dim + d.group_channels - 1) / d.group_channels, d.ndir, d.batch);
  scan_bwd_kernel<I, MULTI><<<grid, kThreads, smem, st>>>(d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int check_desc(const bimamba_scan_desc* d, bool bwd) {
  if (!d) { set_err("null descriptor"); return -1; }
  if (d->dstate != kN) { set_err("dstate must be 16"); return -2; }
  if (d->batch < 0 || d->ndir < 1 || d->ndir > 2 || d->dim < 1 || d->seqlen < 0) { set_err("bad sizes"); return -3; }
  if (d->batch > 65535) { set_err("batch > 65535 not supported by this launch geometry"); return -3; }
  if (d->chunk_items < 1 || d->chunk_items > 8) { set_err("chunk_items must be 1..8"); return -4; }
  if (d->group_channels < 1 || d->group_channels > kMaxCpw * kWarps) { set_err("group_channels must be 1..32"); return -5; }
  if (d->io_dtype < 0 || d->io_dtype > 2 || d->bc_dtype < 0 || d->bc_dtype > 2) { set_err("bad dtype"); return -6; }
  if (!d->u || !d->delta || !d->A || !d->Bm || !d->Cm) { set_err("null operand"); return -7; }
  if (!bwd && !d->out) { set_err("null out"); return -7; }
  const int nchunks = (d->seqlen + 32 * d->chunk_items - 1) / (32 * d->chunk_items);
  if (bwd) {
    if (!d->dout || !d->du || !d->ddelta || !d->dBC_part || !d->dA_part) { set_err("null backward operand"); return -8; }
    if (nchunks > 1 && !d->ckpt) { set_err("backward over several chunks needs the forward checkpoints"); return -9; }
    if (d->dbc_rs < d->seqlen) { set_err("dbc_rs < seqlen"); return -10; }
  }
  return 0;
}

#define BIMAMBA_DISPATCH_I(FN, d, st)                                         \
  switch ((d).chunk_items) {                                                  \
    case 1: return multi ? FN<1, true>(d, st) : FN<1, false>(d, st);          \
    case 2: return multi ? FN<2, true>(d, st) : FN<2, false>(d, st);          \
    case 3: return multi ? FN<3, true>(d, st) : FN<3, false>(d, st);          \
    case 4: return multi ? FN<4, true>(d, st) : FN<4, false>(d, st);          \
    case 5: return multi ? FN<5, true>(d, st) : FN<5, false>(d, st);          \
    case 6: return multi ? FN<6, true>(d, st) : FN<6, false>(d, st);          \
    case 7: return multi ? FN<7, true>(d, st) : FN<7, false>(d, st);          \
    default: return multi ? FN<8, true>(d, st) : FN<8, false>(d, st);         \
  }

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_abi_version(void) { return BIMAMBA_ABI_VERSION; }
extern "C" const char* bimamba_last_error(void) { return g_err; }

extern "C" int bimamba_scan_plan(int seqlen, int dim, int rows, int backward, int* chunk_items, int* group_channels) {
  (void)backward;
  int I = (seqlen + 31) / 32;
  if (I < 1) I = 1;
  if (I > 8) I = 8;
  int G = kMaxCpw * kWarps;
  // keep at least ~2 CTAs per SM worth of blocks when the batch is small
  while (G > kWarps && (int64_t)rows * ((dim + G - 1) / G) < 2 * 148) G /= 2;
  if (chunk_items) *chunk_items = I;
  if (group_channels) *group_channels = G;
  const int tc = 32 * I;
  return seqlen > 0 ? (seqlen + tc - 1) / tc : 1;
}

extern "C" int bimamba_selective_scan_fwd(const bimamba_scan_desc* d, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;  // empty: nothing to do (pointers may be null)
  int rc = check_desc(d, false);
  if (rc) return rc;
  const bool multi = d->seqlen > 32 * d->chunk_items;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  BIMAMBA_DISPATCH_I(launch_fwd, *d, st)
}

extern "C" int bimamba_selective_scan_bwd(const bimamba_scan_desc* d, bimamba_stream_t stream) {
  if (d && (d->batch == 0 || d->seqlen == 0)) return 0;
  int rc = check_desc(d, true);
  if (rc) return rc;
  const bool multi = d->seqlen > 32 * d->chunk_items;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  BIMAMBA_DISPATCH_I(launch_bwd, *d, st)
}
