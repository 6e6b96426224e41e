"""Autograd Functions and functional API of the Bi-Mamba hot path.

Public names mirror what the reference's path binds upstream:
  * ``selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus)``  - mamba_ssm's op;
    reference math at src/models/modules/mamba_block.py:80-120 and :61
  * ``causal_conv1d_fn(x, weight, bias, activation)``                           - causal-conv1d's op;
    reference math at src/models/modules/mamba_block.py:52-55
  * ``bimamba_inner_fn(...)``  - the fused block: in_proj, conv+SiLU, x_proj, dt_proj, scan, gate,
    out_proj in BOTH time directions with shared weights
    (src/models/DualStreamSEMamba.py:473-481 around mamba_block.py:41-63)

All arithmetic runs in the CUDA library (ops on CPU tensors raise: there is no fallback).
Activations are channel-first (batch, dim, L) like the upstream ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ScanDesc

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}
D_STATE = 16


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("Bi-Mamba ops run on CUDA tensors only (no CPU fallback)")


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; use float32, bfloat16 or float16") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _s3(t: torch.Tensor) -> Tuple[int, int, int]:
    """(batch, dir, row) element strides of a (batch, ndir, rows, L) tensor with unit inner stride."""
    assert t.dim() == 4
    if t.size(3) > 1 and t.stride(3) != 1:
        raise ValueError("time must be the contiguous (innermost) axis")
    return t.stride(0), (t.stride(1) if t.size(1) > 1 else 0), t.stride(2)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def _timed(name: str):
    t = _lib.kernel_timer
    return _NULL if t is None else t(name)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().to(torch.float32).contiguous()


# ----------------------------------------------------------------------------------------
# raw kernels (no autograd)
# ----------------------------------------------------------------------------------------
def conv_fwd_raw(x, weight, bias, out, seqlen: int, silu: bool):
    """x (B, D, >=L) strided; weight (D, K) fp32; out (B, ndir, D, Lp) strided."""
    lib = _lib.load()
    B, D = x.shape[0], x.shape[1]
    ndir, Lp = out.shape[1], out.shape[3]
    obs, ods, ors = _s3(out)
    rc = lib.bimamba_causal_conv1d_fwd(
        _ptr(x), _ptr(weight), _ptr(bias), _ptr(out), B, ndir, D, seqlen, Lp, weight.shape[1],
        x.stride(0), x.stride(1), obs, ods, ors, _dt(x), _lib.FLAG_SILU if silu else 0, _stream())
    _lib.check(rc, "bimamba_causal_conv1d_fwd")


def conv_bwd_raw(x, weight, bias, dout, dx, seqlen: int, silu: bool):
    """dout (B, ndir, D, Lp); dx (B, D, Lp) strided.  Returns dwb (D, K+1) fp32 [dw | dbias]."""
    lib = _lib.load()
    B, D = x.shape[0], x.shape[1]
    ndir = dout.shape[1]
    K = weight.shape[1]
    gbs, gds, grs = _s3(dout)
    part = torch.empty((B, D, K + 1), device=x.device, dtype=torch.float32)
    rc = lib.bimamba_causal_conv1d_bwd(
        _ptr(x), _ptr(weight), _ptr(bias), _ptr(dout), _ptr(dx), _ptr(part), B, ndir, D, seqlen, dx.shape[2], K,
        x.stride(0), x.stride(1), gbs, gds, grs, dx.stride(0), dx.stride(1), _dt(x),
        _lib.FLAG_SILU if silu else 0, _stream())
    _lib.check(rc, "bimamba_causal_conv1d_bwd")
    dwb = torch.empty((D, K + 1), device=x.device, dtype=torch.float32)
    reduce_raw(part, dwb, groups=1, rows=B, cols=D * (K + 1), part_gs=0, row_stride=D * (K + 1), out_gs=0)
    return dwb


def reduce_raw(part, out, groups, rows, cols, part_gs, row_stride, out_gs, accumulate=False):
    lib = _lib.load()
    rc = lib.bimamba_reduce_partials(_ptr(part), _ptr(out), groups, rows, cols, part_gs, row_stride, out_gs,
                                     _dt(out), int(accumulate), _stream())
    _lib.check(rc, "bimamba_reduce_partials")


def _fill_desc(u, delta, A, Bm, Cm, D, z, delta_bias, softplus, seqlen, plan) -> ScanDesc:
    d = ScanDesc()
    Bsz, ndir, dim, Lp = u.shape
    d.u, d.delta, d.z = _ptr(u), _ptr(delta), _ptr(z)
    d.Bm, d.Cm, d.A, d.D, d.delta_bias = _ptr(Bm), _ptr(Cm), _ptr(A), _ptr(D), _ptr(delta_bias)
    d.batch, d.ndir, d.dim, d.seqlen, d.dstate = Bsz, ndir, dim, seqlen, A.shape[1]
    d.io_dtype, d.bc_dtype = _dt(u), _dt(Bm)
    d.flags = _lib.FLAG_SOFTPLUS if softplus else 0
    d.chunk_items, d.group_channels = plan[0], plan[1]
    d.pad_to = Lp
    d.u_bs, d.u_ds, d.u_rs = _s3(u)
    d.delta_bs, d.delta_ds, d.delta_rs = _s3(delta)
    if z is not None:
        d.z_bs, d.z_ds, d.z_rs = _s3(z)
    d.bc_bs, d.bc_ds, d.bc_rs = _s3(Bm)
    if _s3(Cm) != _s3(Bm):
        raise ValueError("B and C must share strides")
    return d


def scan_fwd_raw(u, delta, A, Bm, Cm, D, z, delta_bias, softplus: bool, seqlen: int, want_ckpt: bool):
    """All activations 4-D (B, ndir, rows, Lp) with unit inner stride.  Returns (out, ckpt, plan)."""
    lib = _lib.load()
    Bsz, ndir, dim, Lp = u.shape
    plan = _lib.scan_plan(seqlen, dim, Bsz * ndir, want_ckpt)
    nchunks = plan[2]
    out = torch.empty((Bsz, ndir, dim, Lp), device=u.device, dtype=u.dtype)
    ckpt = ypre = None
    if want_ckpt and nchunks > 1:
        ckpt = torch.empty((Bsz, ndir, dim, nchunks, A.shape[1]), device=u.device, dtype=torch.float32)
    if want_ckpt and z is not None:
        ypre = torch.empty((Bsz, ndir, dim, Lp), device=u.device, dtype=u.dtype)
    d = _fill_desc(u, delta, A, Bm, Cm, D, z, delta_bias, softplus, seqlen, plan)
    d.out = _ptr(out)
    d.out_bs, d.out_ds, d.out_rs = _s3(out)
    d.ckpt = _ptr(ckpt)
    d.ypre = _ptr(ypre)            # written with out's strides
    with _timed("scan_fwd"):
        _lib.check(lib.bimamba_selective_scan_fwd(C.byref(d), _stream()), "bimamba_selective_scan_fwd")
    return out, ckpt, ypre, plan


def scan_bwd_raw(u, delta, A, Bm, Cm, D, z, delta_bias, softplus: bool, seqlen: int, dout, ckpt, ypre, plan,
                 dz_out=None, bc_out_dtype=None):
    """Returns (du, ddelta, dz, dBC (B, ndir, 2N, Lp), dA (dim, N), dD (dim), dbias (dim))."""
    lib = _lib.load()
    Bsz, ndir, dim, Lp = u.shape
    N = A.shape[1]
    G = plan[1]
    ngroups = (dim + G - 1) // G
    dev = u.device
    du = torch.empty_like(u)
    ddelta = torch.empty((Bsz, ndir, dim, Lp), device=dev, dtype=u.dtype)
    dz = None
    if z is not None:
        dz = dz_out if dz_out is not None else torch.empty((Bsz, ndir, dim, Lp), device=dev, dtype=u.dtype)
    dBC_part = torch.empty((Bsz, ndir, ngroups, 2 * N, Lp), device=dev, dtype=torch.float32)
    dA_part = torch.empty((Bsz * ndir, dim, N), device=dev, dtype=torch.float32)
    dD_part = torch.empty((Bsz * ndir, dim), device=dev, dtype=torch.float32) if D is not None else None
    db_part = torch.empty((Bsz * ndir, dim), device=dev, dtype=torch.float32) if delta_bias is not None else None

    d = _fill_desc(u, delta, A, Bm, Cm, D, z, delta_bias, softplus, seqlen, plan)
    d.dout = _ptr(dout)
    d.out_bs, d.out_ds, d.out_rs = _s3(dout)
    d.ckpt = _ptr(ckpt)
    d.ypre = _ptr(ypre)
    if ypre is not None:
        d.ypre_bs, d.ypre_ds, d.ypre_rs = _s3(ypre)
    d.du, d.ddelta, d.dz = _ptr(du), _ptr(ddelta), _ptr(dz)
    if _s3(du) != _s3(u) or _s3(ddelta) != _s3(delta):
        # du / ddelta are written with u's / delta's strides
        raise ValueError("u and delta must be dense (B, ndir, dim, Lp) tensors for backward")
    if dz is not None:
        d.dz_bs, d.dz_ds, d.dz_rs = _s3(dz)
    d.dBC_part, d.dA_part, d.dD_part, d.dbias_part = _ptr(dBC_part), _ptr(dA_part), _ptr(dD_part), _ptr(db_part)
    d.dbc_rs = Lp
    with _timed("scan_bwd"):
        _lib.check(lib.bimamba_selective_scan_bwd(C.byref(d), _stream()), "bimamba_selective_scan_bwd")

    dBC = torch.empty((Bsz, ndir, 2 * N, Lp), device=dev, dtype=bc_out_dtype or Bm.dtype)
    cols = 2 * N * Lp
    reduce_raw(dBC_part, dBC, groups=Bsz * ndir, rows=ngroups, cols=cols, part_gs=ngroups * cols,
               row_stride=cols, out_gs=cols)
    dA = torch.empty((dim, N), device=dev, dtype=torch.float32)
    reduce_raw(dA_part, dA, 1, Bsz * ndir, dim * N, 0, dim * N, 0)
    dD = dbias = None
    if D is not None:
        dD = torch.empty((dim,), device=dev, dtype=torch.float32)
        reduce_raw(dD_part, dD, 1, Bsz * ndir, dim, 0, dim, 0)
    if delta_bias is not None:
        dbias = torch.empty((dim,), device=dev, dtype=torch.float32)
        reduce_raw(db_part, dbias, 1, Bsz * ndir, dim, 0, dim, 0)
    return du, ddelta, dz, dBC, dA, dD, dbias


# ----------------------------------------------------------------------------------------
# op-level autograd Functions (single direction, upstream signatures)
# ----------------------------------------------------------------------------------------
class CausalConv1dFn(torch.autograd.Function):
    """causal_conv1d_fn: x (B, D, L), weight (D, K), bias (D) -> (B, D, L).  mamba_block.py:52-55."""

    @staticmethod
    def forward(ctx, x, weight, bias, activation):
        if activation not in (None, "silu", "swish"):
            raise NotImplementedError("activation must be None, silu, or swish")
        _require_cuda(x, weight, bias)
        silu = activation is not None
        if x.stride(2) != 1:
            x = x.contiguous()
        w32, b32 = _f32c(weight), _f32c(bias)
        Bsz, D, L = x.shape
        out = torch.empty((Bsz, D, L), device=x.device, dtype=x.dtype)
        conv_fwd_raw(x, w32, b32, out.unsqueeze(1), L, silu)
        ctx.save_for_backward(x, w32, b32 if b32 is not None else torch.empty(0))
        ctx.silu, ctx.has_bias = silu, bias is not None
        ctx.wdtype = weight.dtype
        ctx.bdtype = bias.dtype if bias is not None else None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w32, b32 = ctx.saved_tensors
        b32 = b32 if ctx.has_bias else None
        dout = dout.to(x.dtype)
        if dout.stride(2) != 1:
            dout = dout.contiguous()
        dx = torch.empty_like(x, memory_format=torch.contiguous_format)
        dwb = conv_bwd_raw(x, w32, b32, dout.unsqueeze(1), dx, x.shape[2], ctx.silu)
        K = w32.shape[1]
        dw = dwb[:, :K].to(ctx.wdtype)
        db = dwb[:, K].to(ctx.bdtype) if ctx.has_bias else None
        return dx, dw, db, None


def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                     final_states_out=None, activation=None):
    if seq_idx is not None or initial_states is not None or return_final_states or final_states_out is not None:
        raise NotImplementedError("seq_idx / initial_states / final_states are not used by the reference path")
    return CausalConv1dFn.apply(x, weight, bias, activation)


def _as_bnl(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() == 4:
        if t.shape[1] != 1:
            raise NotImplementedError(f"{name}: grouped B/C (G > 1) is not used by the reference path")
        t = t[:, 0]
    if t.dim() != 3:
        raise NotImplementedError(f"{name} must be input-dependent: (batch, dstate, L)")
    return t


class SelectiveScanFn(torch.autograd.Function):
    """selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus).
    u, delta, z (B, D, L); A (D, N); B, C (B, N, L).  mamba_block.py:80-120, :61."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, z, delta_bias, delta_softplus):
        _require_cuda(u, delta, A, B, C, D, z, delta_bias)
        Bshape, Cshape = B.shape, C.shape
        Bm, Cm = _as_bnl(B, "B"), _as_bnl(C, "C")
        if A.shape[1] != D_STATE:
            raise NotImplementedError("d_state must be 16 (the Phase-6 configuration)")
        u = u.contiguous()
        delta = delta.to(u.dtype).contiguous()
        zc = z.to(u.dtype).contiguous() if z is not None else None
        Bm = Bm.contiguous()
        Cm = Cm.to(Bm.dtype).contiguous()
        A32, D32, b32 = _f32c(A), _f32c(D), _f32c(delta_bias)
        L = u.shape[2]
        needs_bwd = any(t is not None and t.requires_grad for t in (u, delta, A, B, C, D, z, delta_bias))
        out, ckpt, ypre, plan = scan_fwd_raw(u.unsqueeze(1), delta.unsqueeze(1), A32, Bm.unsqueeze(1),
                                             Cm.unsqueeze(1), D32, None if zc is None else zc.unsqueeze(1), b32,
                                             bool(delta_softplus), L, needs_bwd)
        ctx.save_for_backward(u, delta, A32, Bm, Cm, D32 if D32 is not None else torch.empty(0),
                              zc if zc is not None else torch.empty(0), b32 if b32 is not None else torch.empty(0),
                              ckpt if ckpt is not None else torch.empty(0),
                              ypre if ypre is not None else torch.empty(0))
        ctx.meta = (D is not None, z is not None, delta_bias is not None, bool(delta_softplus), plan,
                    A.dtype, None if D is None else D.dtype, None if delta_bias is None else delta_bias.dtype,
                    B.dtype, C.dtype, Bshape, Cshape, None if z is None else z.dtype)
        return out[:, 0]

    @staticmethod
    def backward(ctx, dout):
        u, delta, A32, Bm, Cm, D32, zc, b32, ckpt, ypre = ctx.saved_tensors
        (hasD, hasz, hasb, softplus, plan, Adt, Ddt, bdt, Bdt, Cdt, Bshape, Cshape, zdt) = ctx.meta
        D32 = D32 if hasD else None
        zc = zc if hasz else None
        b32 = b32 if hasb else None
        ckpt = ckpt if ckpt.numel() else None
        ypre = ypre if ypre.numel() else None
        dout = dout.to(u.dtype).contiguous()
        L = u.shape[2]
        du, ddelta, dz, dBC, dA, dD, dbias = scan_bwd_raw(
            u.unsqueeze(1), delta.unsqueeze(1), A32, Bm.unsqueeze(1), Cm.unsqueeze(1), D32,
            None if zc is None else zc.unsqueeze(1), b32, softplus, L, dout.unsqueeze(1), ckpt, ypre, plan,
            bc_out_dtype=torch.float32)
        N = A32.shape[1]
        dB = dBC[:, 0, :N].to(Bdt).reshape(Bshape)
        dC = dBC[:, 0, N:].to(Cdt).reshape(Cshape)
        return (du[:, 0], ddelta[:, 0], dA.to(Adt), dB, dC,
                dD.to(Ddt) if hasD else None, dz[:, 0].to(zdt) if hasz else None,
                dbias.to(bdt) if hasb else None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    if return_last_state:
        raise NotImplementedError("return_last_state is not used by the reference path")
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus)


# ----------------------------------------------------------------------------------------
# the fused block: both directions, shared weights
# ----------------------------------------------------------------------------------------
class BiMambaInnerFn(torch.autograd.Function):
    """out = M(x) [+ flip(M(flip(x)))] for one Mamba block M with shared weights.

    Reference: mamba_block.py:41-63 for M, DualStreamSEMamba.py:473-481 for the two directions.
    Uses: in_proj(flip x) = flip(in_proj x) so xz is computed once; the reverse
    direction reads the same x, z back to front; out_proj is applied once to y_fwd + y_rev.
    GEMMs are library calls in this version (cuBLAS through torch.matmul); conv, scan and all
    reductions are this repository's CUDA kernels.
    """

    @staticmethod
    def forward(ctx, x, W_in, conv_w, conv_b, W_x, W_dt, b_dt, A_log, Dp, W_out, bidirectional, cdtype):
        _require_cuda(x, W_in, conv_w, conv_b, W_x, W_dt, b_dt, A_log, Dp, W_out)
        with torch.autocast("cuda", enabled=False):
            Bsz, L, dm = x.shape
            D = W_in.shape[0] // 2
            N = A_log.shape[1]
            R = W_dt.shape[1]
            if N != D_STATE:
                raise NotImplementedError("d_state must be 16 (the Phase-6 configuration)")
            ndir = 2 if bidirectional else 1
            Lp = round_up(max(L, 1), 8)
            dev = x.device
            Wi, Wx, Wd, Wo = (w.detach().to(cdtype) for w in (W_in, W_x, W_dt, W_out))
            cw32 = _f32c(conv_w).reshape(D, -1)
            cb32 = _f32c(conv_b)
            A32 = -torch.exp(A_log.detach().float())
            D32, bdt32 = _f32c(Dp), _f32c(b_dt)

            x_pad = torch.zeros((Bsz, Lp, dm), device=dev, dtype=cdtype)
            x_pad[:, :L] = x.detach()
            xz = torch.matmul(Wi, x_pad.transpose(1, 2))                      # (B, 2D, Lp)   mamba_block.py:48
            xs, z = xz[:, :D], xz[:, D:]                                      # :49 (views)
            xc = torch.empty((Bsz, ndir, D, Lp), device=dev, dtype=cdtype)
            conv_fwd_raw(xs, cw32, cb32, xc, L, True)                         # :52-55, both directions
            x_dbl = torch.matmul(Wx, xc)                                      # (B, ndir, R+2N, Lp)   :73
            delta = torch.matmul(Wd, x_dbl[:, :, :R])                         # (B, ndir, D, Lp)      :80 (pre-bias)
            needs_bwd = any(ctx.needs_input_grad)
            y, ckpt, ypre, plan = scan_fwd_raw(xc, delta, A32, x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], D32,
                                               z.unsqueeze(1), bdt32, True, L, needs_bwd)   # :82-120, :61
            ysum = y[:, 0] + y[:, 1] if ndir == 2 else y[:, 0]                # DualStreamSEMamba.py:481 (before out_proj)
            out = torch.matmul(ysum.transpose(1, 2)[:, :L], Wo.t())           # (B, L, dm)   mamba_block.py:62
            if needs_bwd:
                ctx.save_for_backward(x_pad, xz, xc, x_dbl, delta, ysum, Wi, Wx, Wd, Wo, cw32, cb32, A32, D32, bdt32,
                                      ckpt if ckpt is not None else torch.empty(0), ypre)
                ctx.meta = (L, ndir, plan, x.dtype,
                            tuple(t.dtype for t in (W_in, conv_w, conv_b, W_x, W_dt, b_dt, A_log, Dp, W_out)),
                            tuple(conv_w.shape))
            return out

    @staticmethod
    def backward(ctx, dout):
        (x_pad, xz, xc, x_dbl, delta, ysum, Wi, Wx, Wd, Wo, cw32, cb32, A32, D32, bdt32, ckpt, ypre) = ctx.saved_tensors
        L, ndir, plan, xdt, pdt, cw_shape = ctx.meta
        ckpt = ckpt if ckpt.numel() else None
        with torch.autocast("cuda", enabled=False):
            cd = xz.dtype
            Bsz, Lp, dm = x_pad.shape
            D = xz.shape[1] // 2
            N = A32.shape[1]
            R = Wd.shape[1]
            dev = xz.device
            xs, z = xz[:, :D], xz[:, D:]

            dout_pad = torch.zeros((Bsz, Lp, dm), device=dev, dtype=cd)
            dout_pad[:, :L] = dout
            # out_proj
            dy = torch.matmul(Wo.t(), dout_pad.transpose(1, 2))               # (B, D, Lp), shared by both directions
            dW_out = torch.bmm(dout_pad.transpose(1, 2), ysum.transpose(1, 2)).sum(0)   # (dm, D)
            # scan (both directions in one launch); dy is broadcast over the direction axis
            dxz = torch.empty_like(xz)
            dz2 = torch.empty((Bsz, ndir, D, Lp), device=dev, dtype=cd)
            dyb = dy.unsqueeze(1).expand(Bsz, ndir, D, Lp)
            du, ddelta, dz2, dBC, dA, dD, dbdt = scan_bwd_raw(
                xc, delta, A32, x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], D32, z.unsqueeze(1), bdt32, True, L,
                dyb, ckpt, ypre, plan, dz_out=dz2, bc_out_dtype=cd)
            if ndir == 2:
                torch.add(dz2[:, 0], dz2[:, 1], out=dxz[:, D:])
            else:
                dxz[:, D:].copy_(dz2[:, 0])
            # dt_proj
            ddtr = torch.matmul(Wd.t(), ddelta)                               # (B, ndir, R, Lp)
            dW_dt = torch.bmm(ddelta.flatten(0, 1), x_dbl[:, :, :R].flatten(0, 1).transpose(1, 2)).sum(0)   # (D, R)
            # x_proj
            dxdbl = torch.cat([ddtr, dBC], dim=2)                             # (B, ndir, R+2N, Lp)
            dW_x = torch.bmm(dxdbl.flatten(0, 1), xc.flatten(0, 1).transpose(1, 2)).sum(0)                  # (R+2N, D)
            dxc = torch.baddbmm(du.flatten(0, 1), Wx.t().unsqueeze(0).expand(Bsz * ndir, D, R + 2 * N),
                                dxdbl.flatten(0, 1)).view(Bsz, ndir, D, Lp)
            # conv (writes dx into the x half of dxz)
            dwb = conv_bwd_raw(xs, cw32, cb32, dxc, dxz[:, :D], L, True)
            K = cw32.shape[1]
            # in_proj
            dW_in = torch.bmm(dxz, x_pad).sum(0)                              # (2D, dm)
            dx = torch.matmul(dxz.transpose(1, 2)[:, :L], Wi)                 # (B, L, dm)
            dA_log = dA * A32                                                 # A = -exp(A_log)
        return (dx.to(xdt), dW_in.to(pdt[0]), dwb[:, :K].reshape(cw_shape).to(pdt[1]), dwb[:, K].to(pdt[2]),
                dW_x.to(pdt[3]), dW_dt.to(pdt[4]), dbdt.to(pdt[5]), dA_log.to(pdt[6]), dD.to(pdt[7]),
                dW_out.to(pdt[8]), None, None)


def bimamba_inner_fn(x, W_in, conv_w, conv_b, W_x, W_dt, b_dt, A_log, Dp, W_out, bidirectional=True,
                     compute_dtype=None):
    """x (B, L, d_model) -> (B, L, d_model).  compute_dtype: activation dtype of the kernels and
    GEMMs (default: the autocast dtype when autocast is on, else x.dtype); scan state is fp32."""
    if compute_dtype is None:
        compute_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    return BiMambaInnerFn.apply(x, W_in, conv_w, conv_b, W_x, W_dt, b_dt, A_log, Dp, W_out, bool(bidirectional),
                                compute_dtype)
