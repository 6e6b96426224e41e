// Host-side pieces of the C ABI shared by the kernels: error string, descriptor checks, launch plan.
#include "common.cuh"

namespace bimamba {

thread_local char g_err[512] = "";
void set_err(const char* msg) {
  size_t i = 0;
  for (; msg[i] && i + 1 < sizeof(g_err); ++i) g_err[i] = msg[i];
  g_err[i] = 0;
}

// Tuning knobs (tests and tuning experiments force kernel variants through bimamba_set_tuning; launches read these
// process-wide integers - no getenv on any launch path).  0 = automatic choice.
int g_tune[BIMAMBA_TUNE_COUNT] = {0, 0, 0, 0, 0, 0, 0};

int check_desc(const bimamba_scan_desc* d, bool bwd) {
  if (!d) { set_err("null descriptor"); return -1; }
  if (d->dstate != kN) { set_err("dstate must be 16"); return -2; }
  if (d->batch < 0 || d->ndir < 1 || d->ndir > 2 || d->dim < 1 || d->seqlen < 0) { set_err("bad sizes"); return -3; }
  if (d->batch > 65535) { set_err("batch > 65535 not supported by this launch geometry"); return -3; }
  if (d->io_dtype < 0 || d->io_dtype > 2) { set_err("bad dtype"); return -6; }
  if (!d->u || !d->A || !d->bc) { set_err("null operand"); return -7; }
  if (!d->delta) {
    if (!d->dtr || !d->Wdt) { set_err("either delta or (dtr, Wdt) must be given"); return -7; }
    if (d->dt_rank < 1 || d->dt_rank > BIMAMBA_MAX_DT_RANK) { set_err("dt_rank must be 1..16"); return -4; }
  }
  if (!bwd && !d->out) { set_err("null out"); return -7; }
  if (bwd) {
    if (!d->dout || !d->du || !d->ddelta || !d->dbc_part || !d->dA_part) { set_err("null backward operand"); return -8; }
    if (d->seqlen > BIMAMBA_CKPT && !d->ckpt) { set_err("backward needs the forward checkpoints"); return -9; }
    if (d->z && d->dz && !d->ypre) { set_err("gated backward needs ypre saved by the forward"); return -11; }
  }
  return 0;
}

}  // namespace bimamba

using namespace bimamba;

extern "C" int bimamba_abi_version(void) { return BIMAMBA_ABI_VERSION; }
extern "C" const char* bimamba_last_error(void) { return g_err; }

extern "C" int bimamba_set_tuning(int knob, int value) {
  if (knob < 0 || knob >= BIMAMBA_TUNE_COUNT) { set_err("set_tuning: unknown knob"); return -1; }
  g_tune[knob] = value;
  return 0;
}

extern "C" int bimamba_get_tuning(int knob) { return (knob < 0 || knob >= BIMAMBA_TUNE_COUNT) ? 0 : g_tune[knob]; }

/* Workspace sizes a non-Python host needs to allocate before the backward calls: the partial-sum buffers
 * of bimamba_selective_scan_bwd and the checkpoint / ypre tensors of bimamba_selective_scan_fwd, in BYTES. */
extern "C" size_t bimamba_scan_fwd_workspace_bytes(int batch, int ndir, int seqlen, int dim, int io_dtype, int want_ckpt) {
  if (batch <= 0 || ndir <= 0 || seqlen <= 0 || dim <= 0 || !want_ckpt) return 0;
  const size_t nck = (size_t)(seqlen + BIMAMBA_CKPT - 1) / BIMAMBA_CKPT;
  const size_t es = io_dtype == BIMAMBA_F32 ? 4 : 2;
  const size_t ckpt = nck > 1 ? (size_t)batch * ndir * nck * dim * kN * 4 : 0;
  return ckpt + (size_t)batch * ndir * seqlen * dim * es;       /* ckpt (fp32) followed by ypre (io dtype) */
}

extern "C" size_t bimamba_scan_bwd_workspace_bytes(int batch, int ndir, int seqlen, int dim) {
  if (batch <= 0 || ndir <= 0 || seqlen <= 0 || dim <= 0) return 0;
  int G = 0, ng = 0;
  bimamba_scan_plan(seqlen, dim, batch * ndir, 1, &G, &ng);
  const size_t dbc = (size_t)batch * ng * seqlen * ndir * 2 * kN * 4;
  const size_t dA = (size_t)batch * ndir * dim * kN * 4;
  const size_t dDb = (size_t)batch * ndir * dim * 4;
  return dbc + dA + 2 * dDb;                                     /* dbc_part, dA_part, dD_part, dbias_part */
}

extern "C" int bimamba_scan_plan(int seqlen, int dim, int rows, int backward, int* group_channels, int* ngroups) {
  int G;
  if (backward) {
    // one warp (32 channels) per CTA: no block barrier anywhere in the kernel.  Measured faster than 96-channel CTAs
    // at both ends (batch 64 x 201 frames and 2048 x 256: 3.6 vs 4.6 ms) despite three times the dB/dC partial rows.
    G = 32;
  } else {
    // one thread per channel: wide groups share the staged B|C|dt_r rows, narrow ones fill the GPU
    G = 32;
    const int cand[3] = {128, 96, 64};
    for (int i = 0; i < 3; ++i) {
      const int g = cand[i];
      if (dim % g == 0 && (int64_t)rows * (dim / g) >= 4 * 148) { G = g; break; }
    }
  }
  if (group_channels) *group_channels = G;
  if (ngroups) *ngroups = (dim + G - 1) / G;
  return seqlen > 0 ? (seqlen + BIMAMBA_CKPT - 1) / BIMAMBA_CKPT : 1;
}
