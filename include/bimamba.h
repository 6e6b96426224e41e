/* bimamba.h - C ABI of libbimamba_sm100.so (B200 / sm_100a Bi-Mamba hot path).
 *
 * Every entry point is `extern "C"`, takes plain pointers, sizes and element strides
 * and a CUDA stream handle; no torch / C++ types cross this boundary.
 *
 * What each entry point replaces in the reference (lux-liang/Robust-Audio-Deepfake-
 * Evolution).  The reference reaches this arithmetic through the third-party
 * `mamba_ssm` package (import at src/models/DualStreamSEMamba.py:43, constructed at
 * :455, called at :473 and :477); the in-repo statement of the same math is
 * src/models/modules/mamba_block.py.  The FFI the reference's path binds upstream is
 * `selective_scan_cuda.{fwd,bwd}` and `causal_conv1d_cuda.causal_conv1d_{fwd,bwd}`;
 * the functions below are their drop-in equivalents (same operand meaning, channel-
 * first (B, D, L) operands), extended with a direction axis so that the forward scan
 * and the flipped scan of DualStreamSEMamba.py:473-481 share ONE launch.
 *
 *   bimamba_causal_conv1d_fwd/bwd    <- mamba_block.py:24-31,52-55 (depthwise causal
 *                                       Conv1d k=4, crop to L, SiLU) and its autograd
 *   bimamba_selective_scan_fwd/bwd   <- mamba_block.py:80 (softplus), :82 (A),
 *                                       :92-117 (scan), :120 (D skip), :61 (z gate)
 *                                       and its autograd
 *   bimamba_reduce_partials          <- the sum over batch/time/direction that
 *                                       autograd performs for parameter gradients
 *
 * Ownership: the library never allocates or frees device memory.  All tensors and
 * workspaces are caller-allocated; pointers are borrowed for the duration of the
 * enqueue.  Calls only enqueue work on `stream` (no device sync, CUDA-graph safe).
 * Errors: 0 = success; negative = argument error (see bimamba_last_error());
 * positive = cudaError_t from the launch.  Nothing throws across this boundary.
 */
#ifndef BIMAMBA_H_
#define BIMAMBA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BIMAMBA_ABI_VERSION 2

/* element types of activation operands */
#define BIMAMBA_F32 0
#define BIMAMBA_BF16 1
#define BIMAMBA_F16 2

/* flags */
#define BIMAMBA_FLAG_SOFTPLUS 1 /* delta = softplus(delta + delta_bias) (mamba_block.py:80) */
#define BIMAMBA_FLAG_SILU 1     /* conv: apply SiLU (mamba_block.py:55) */

typedef void* bimamba_stream_t; /* a cudaStream_t */

/* Strides are in ELEMENTS.  "dir" is the direction axis: dir 0 scans t = 0..L-1,
 * dir 1 scans t = L-1..0 on the same storage order (natural time order in memory), i.e.
 * dir 1 computes flip(op(flip(.))) of DualStreamSEMamba.py:476-478 without any flipped copy. */
typedef struct bimamba_scan_desc {
  /* activations, element type io_dtype */
  const void* u;     /* (batch, ndir, dim, L)   conv output xc          */
  const void* delta; /* (batch, ndir, dim, L)   dt_proj output, pre-softplus */
  const void* z;     /* (batch, ndir, dim, L) or NULL (gate input; z_ds may be 0 to share) */
  const void* Bm;    /* (batch, ndir, dstate, L), element type bc_dtype */
  const void* Cm;    /* (batch, ndir, dstate, L), element type bc_dtype */
  const float* A;    /* (dim, dstate) fp32, = -exp(A_log)               */
  const float* D;    /* (dim) fp32 or NULL                              */
  const float* delta_bias; /* (dim) fp32 or NULL                        */
  void* out;         /* fwd: (batch, ndir, dim, L) io_dtype              */
  float* ckpt;       /* (batch, ndir, dim, nchunks, dstate) fp32 state entering each 16-step chunk;
                        written by fwd (may be NULL when nchunks == 1 or no backward is
                        needed), read by bwd when nchunks > 1             */
  void* ypre;        /* (batch, ndir, dim, L) io_dtype, strides ypre_*: y before the z gate.
                        Written by fwd when non-NULL; read by bwd to form dz (required there
                        when z and dz are given)                          */
  /* backward only */
  const void* dout;  /* (batch, ndir, dim, L) io_dtype, strides = out's  */
  void* du;          /* (batch, ndir, dim, L) io_dtype, strides = u's    */
  void* ddelta;      /* same, strides = delta's                          */
  void* dz;          /* same or NULL; strides dz_bs/dz_ds/dz_rs          */
  float* dBC_part;   /* (batch, ndir, ngroups, 2, dstate, dbc_rs) fp32 partial sums over each
                        channel group; reduce over ngroups with bimamba_reduce_partials */
  float* dA_part;    /* (batch, ndir, dim, dstate) fp32                  */
  float* dD_part;    /* (batch, ndir, dim) fp32 or NULL                  */
  float* dbias_part; /* (batch, ndir, dim) fp32 or NULL                  */

  int32_t batch, ndir, dim, seqlen, dstate;
  int32_t io_dtype, bc_dtype, flags;
  int32_t chunk_items;   /* steps per chunk = checkpoint interval (16);
                            nchunks = ceil(seqlen / chunk_items).  Use bimamba_scan_plan(). */
  int32_t group_channels;/* channels per CTA (even, 2..32); ngroups = ceil(dim / group_channels) */
  int32_t pad_to;        /* if > seqlen: columns [seqlen, pad_to) of every activation output
                            (out; du, ddelta, dz) are written as zeros, so padded rows can be
                            fed to GEMMs that contract over time                */
  int32_t reserved0;
  int64_t u_bs, u_ds, u_rs;
  int64_t delta_bs, delta_ds, delta_rs;
  int64_t z_bs, z_ds, z_rs;
  int64_t bc_bs, bc_ds, bc_rs;
  int64_t out_bs, out_ds, out_rs;
  int64_t dz_bs, dz_ds, dz_rs;
  int64_t dbc_rs;        /* row stride of dBC_part (>= seqlen); columns [seqlen, dbc_rs) are zeroed */
  int64_t ypre_bs, ypre_ds, ypre_rs;
} bimamba_scan_desc;

int bimamba_abi_version(void);
const char* bimamba_last_error(void);

/* Fills chunk_items / group_channels for (seqlen, dim, batch*ndir); returns nchunks. */
int bimamba_scan_plan(int seqlen, int dim, int rows, int backward, int* chunk_items, int* group_channels);

int bimamba_selective_scan_fwd(const bimamba_scan_desc* d, bimamba_stream_t stream);
int bimamba_selective_scan_bwd(const bimamba_scan_desc* d, bimamba_stream_t stream);

/* Depthwise causal conv (dir 0: taps t-(K-1)..t; dir 1: taps t..t+(K-1), i.e. the causal conv
 * of the time-reversed sequence), K in {2,3,4}.  x: (batch, dim, L); out: (batch, ndir, dim, L).
 * Output columns [seqlen, pad_to) are zero-filled when pad_to > seqlen. */
int bimamba_causal_conv1d_fwd(const void* x, const float* weight /*(dim,K)*/, const float* bias /*(dim) or NULL*/,
                              void* out, int batch, int ndir, int dim, int seqlen, int pad_to, int width,
                              int64_t x_bs, int64_t x_rs, int64_t out_bs, int64_t out_ds, int64_t out_rs,
                              int dtype, int flags, bimamba_stream_t stream);

/* dout: (batch, ndir, dim, L) grads w.r.t. the conv output of each direction.
 * dx: (batch, dim, L) (sum over directions).  dwb_part: (batch, dim, K+1) fp32 partials
 * [dw_0..dw_{K-1}, dbias] per (batch, channel); reduce over batch with bimamba_reduce_partials. */
int bimamba_causal_conv1d_bwd(const void* x, const float* weight, const float* bias, const void* dout,
                              void* dx, float* dwb_part, int batch, int ndir, int dim, int seqlen, int pad_to, int width,
                              int64_t x_bs, int64_t x_rs, int64_t dout_bs, int64_t dout_ds, int64_t dout_rs,
                              int64_t dx_bs, int64_t dx_rs, int dtype, int flags, bimamba_stream_t stream);

/* For every g < groups:  out[g*out_gs + j] (+)= sum_{i < rows} part[g*part_gs + i*row_stride + j],
 * j < cols, summed in fixed order (deterministic).  out element type out_dtype;
 * accumulate != 0 adds to the existing contents. */
int bimamba_reduce_partials(const float* part, void* out, int64_t groups, int64_t rows, int64_t cols,
                            int64_t part_gs, int64_t row_stride, int64_t out_gs,
                            int out_dtype, int accumulate, bimamba_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BIMAMBA_H_ */
