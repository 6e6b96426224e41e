"""CPU-side checks of the drop-in boundary (no compute calls: there is no GPU here):
the C-ABI library loads, exports every function include/bimamba.h declares, the ctypes mirror of
`struct bimamba_scan_desc` has the C layout, argument errors come back as codes (nothing throws),
and the Python surface mirrors the reference's names and refuses CPU tensors (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

import bimamba_b200 as bm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bimamba.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bimamba_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = bm._lib.load()
    names = _declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/bimamba.h but not exported"
    assert sorted(bm._lib.EXPORTS) == names, "the ctypes binding must type exactly the declared entry points"
    assert lib.bimamba_abi_version() == bm._lib.ABI_VERSION


@pytest.mark.parametrize("cname,mirror", [("bimamba_scan_desc", "ScanDesc"), ("bimamba_block_desc", "BlockDesc"),
                                          ("bimamba_block_grads", "BlockGrads"), ("bimamba_layer_desc", "LayerDesc"),
                                          ("bimamba_layer_grads", "LayerGrads")])
def test_struct_layouts_match_header(tmp_path, cname, mirror):
    """sizeof / offsetof of the ctypes mirrors against the C compiler's view of the header."""
    cls = getattr(bm._lib, mirror)
    prog = tmp_path / "layout.c"
    fields = [f[0] for f in cls._fields_]
    body = "\n".join(f'  printf("{f} %zu\\n", offsetof({cname}, {f}));' for f in fields)
    prog.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HEADER}"\nint main(void) {{\n'
                    f'  printf("sizeof %zu\\n", sizeof({cname}));\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", str(prog), "-o", str(exe)])
    out = dict(ln.split() for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == C.sizeof(cls)
    for f in fields:
        assert int(out[f]) == getattr(cls, f).offset, f


def test_argument_errors_are_codes_not_exceptions():
    lib = bm._lib.load()
    assert lib.bimamba_selective_scan_fwd(None, None) == -1
    assert b"null descriptor" in lib.bimamba_last_error()
    d = bm._lib.ScanDesc()
    d.batch, d.ndir, d.dim, d.seqlen, d.dstate = 1, 1, 8, 4, 8          # d_state 8 is not supported
    assert lib.bimamba_selective_scan_fwd(C.byref(d), None) == -2
    d.dstate = 16
    assert lib.bimamba_selective_scan_bwd(C.byref(d), None) < 0          # null operands
    assert lib.bimamba_causal_conv1d_fwd(None, None, None, None, 1, 1, 8, 4, 4, 0, 0, 0, 0, 0, 0, 0, None) == -1
    assert lib.bimamba_layernorm_fwd(None, None, None, None, None, None, 4, 8, 1e-5, 0, 0, None) == -1
    # empty problems are a no-op success (the reference accepts zero-length batches)
    d.batch = 0
    assert lib.bimamba_selective_scan_fwd(C.byref(d), None) == 0
    assert lib.bimamba_reduce_partials(None, None, 0, 0, 0, 0, 0, 0, 0, 0, None) == 0


def test_scan_plan_geometry():
    for L, dim, rows, bwd in [(201, 288, 128, False), (201, 288, 128, True), (8192, 288, 64, True), (1, 17, 1, False)]:
        g, ng, nck = bm._lib.scan_plan(L, dim, rows, bwd)
        assert g % 32 == 0 and 32 <= g <= 128
        assert ng == -(-dim // g)
        assert nck == -(-L // 8)


def test_host_side_geometry_of_the_new_entry_points():
    """Launch-geometry helpers are pure host functions: workspace sizes the Python side allocates from."""
    lib = bm._lib.load()
    # backward scan: one warp per CTA -> 32-channel groups at every size
    for rows in (2, 128, 4096):
        g, ng, _ = bm._lib.scan_plan(201, 288, rows, True)
        assert (g, ng) == (32, 9)
    # weight-gradient GEMM: the split over the rows fills about one wave and never exceeds the 64-row blocks
    for M, n1, n2 in [(12864, 576, 144), (25728, 288, 48), (64, 64, 64), (1, 8, 8), (130, 128, 16)]:
        ns = lib.bimamba_gemm_tn_splits(M, n1, n2)
        assert 1 <= ns <= max(1, -(-M // 64)) and ns <= 148
    assert lib.bimamba_gemm_tn_splits(12864, 576, 144) > 8
    assert lib.bimamba_adamw_chunk() == 4096
    assert lib.bimamba_conv_bwd_slices(64, 201, 288) == 64 * 13
    # argument errors of the new entry points are codes too
    assert lib.bimamba_gemm_tn(None, 8, None, 8, None, None, 64, 8, 8, 1, None) == -1
    assert lib.bimamba_adamw_step(None, None, 4, None, None, None) == -1
    assert lib.bimamba_adamw_step(None, None, 0, None, None, None) == 0
    assert lib.bimamba_head_fwd(None, None, None, None, None, None, None, None, None, 2, 5, 144, 2, 1e-5, 0, None) == -1
    assert lib.bimamba_head_fwd(None, None, None, None, None, None, None, None, None, 0, 5, 144, 2, 1e-5, 0, None) == 0


def test_python_surface_mirrors_reference_and_has_no_cpu_fallback():
    m = bm.Mamba(144, 16)                                   # DualStreamSEMamba.py:455 positional call
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == {
        "A_log": (288, 16), "D": (288,), "in_proj.weight": (576, 144), "conv1d.weight": (288, 1, 4),
        "conv1d.bias": (288,), "x_proj.weight": (41, 288), "dt_proj.weight": (288, 9), "dt_proj.bias": (288,),
        "out_proj.weight": (144, 288)}
    enc = bm.PN_BiMambas_Encoder(144, 16)                   # DualStreamSEMamba.py:451-465 attribute names
    assert {n for n, _ in enc.named_children()} == {"mamba", "norm1", "norm2", "feed_forward"}
    x = torch.randn(2, 5, 144)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bm.selective_scan_fn(torch.zeros(1, 8, 4), torch.zeros(1, 8, 4), -torch.ones(8, 16), torch.zeros(1, 16, 4),
                             torch.zeros(1, 16, 4))
    pkg = bm.install_mamba_ssm_shim()
    from mamba_ssm.modules.mamba_simple import Mamba as ShimMamba   # the import at DualStreamSEMamba.py:43
    assert ShimMamba is bm.Mamba and pkg.Mamba is bm.Mamba


def test_workspace_queries_and_tuning_knobs():
    """bimamba_*_workspace_bytes reproduce what ops.py allocates (a non-Python host sizes its workspaces from
    these); the tuning knobs are plain process-wide integers (no getenv on launch paths)."""
    lib = bm._lib.load()
    B, ndir, L, D, N = 64, 2, 201, 288, 16
    nck = -(-L // 8)
    assert lib.bimamba_scan_fwd_workspace_bytes(B, ndir, L, D, bm._lib.BF16, 1) == B * ndir * nck * D * N * 4 + B * ndir * L * D * 2
    assert lib.bimamba_scan_fwd_workspace_bytes(B, ndir, L, D, bm._lib.F32, 1) == B * ndir * nck * D * N * 4 + B * ndir * L * D * 4
    assert lib.bimamba_scan_fwd_workspace_bytes(B, ndir, L, D, bm._lib.BF16, 0) == 0
    assert lib.bimamba_scan_fwd_workspace_bytes(1, 1, 8, 32, bm._lib.F32, 1) == 8 * 32 * 4       # one chunk: no checkpoints
    ng = bm._lib.scan_plan(L, D, B * ndir, True)[1]
    assert lib.bimamba_scan_bwd_workspace_bytes(B, ndir, L, D) == 4 * (B * ng * L * ndir * 2 * N + B * ndir * D * N + 2 * B * ndir * D)
    assert lib.bimamba_scan_bwd_workspace_bytes(0, 2, 201, 288) == 0
    assert lib.bimamba_get_tuning(bm._lib.TUNE_SCAN_FWD) == 0
    with bm._lib.tuning(bm._lib.TUNE_SCAN_FWD, 3):
        assert lib.bimamba_get_tuning(bm._lib.TUNE_SCAN_FWD) == 3
    assert lib.bimamba_get_tuning(bm._lib.TUNE_SCAN_FWD) == 0
    assert lib.bimamba_set_tuning(99, 1) == -1
    # new round-2 entry points: argument errors are codes
    assert lib.bimamba_gelu_fwd(None, None, 8, 0, None) == -1
    assert lib.bimamba_gelu_fwd(None, None, 0, 0, None) == 0
    assert lib.bimamba_reduce_rows32(None, None, 2, 9, 402, 48, 1, None) == -1
    assert lib.bimamba_head_pool_bwd(None, None, None, None, None, None, None, None, 2, 5, 144, 1e-5, 0, None) == -1
    assert lib.bimamba_finalize_param_grads(*([None] * 12), 144, 288, 16, 9, 2, 4, None) == -1


def test_native_block_entry_points_host_side():
    """bimamba_block_fwd / bimamba_block_bwd (the whole block in one call each way): workspace sizes follow the documented
    carve (256-byte aligned pieces), argument errors are codes, empty problems are a no-op success."""
    lib = bm._lib.load()
    B, L, dm, D, ndir, K, N = 64, 201, 144, 288, 2, 4, 16
    up = lambda v: (v + 255) // 256 * 256
    M, es, nck = B * L, 2, -(-L // 8)
    infer = up(M * 2 * D * es) + up(M * ndir * D * es) + up(M * ndir * 48 * es) + up(M * ndir * D * es)
    train = infer + up(B * ndir * nck * D * N * 4) + up(M * ndir * D * es)
    assert lib.bimamba_block_fwd_workspace_bytes(B, L, dm, D, ndir, bm._lib.BF16, 0) == infer
    assert lib.bimamba_block_fwd_workspace_bytes(B, L, dm, D, ndir, bm._lib.BF16, 1) == train
    assert lib.bimamba_block_fwd_workspace_bytes(2, 8, dm, D, ndir, bm._lib.BF16, 1) == (      # one chunk: no checkpoints
        up(16 * 2 * D * es) + 2 * up(16 * ndir * D * es) + up(16 * ndir * 48 * es) + up(16 * ndir * D * es))
    assert lib.bimamba_block_fwd_workspace_bytes(0, L, dm, D, ndir, bm._lib.BF16, 1) == 0
    bwd = lib.bimamba_block_bwd_workspace_bytes(B, L, dm, D, K, ndir, bm._lib.BF16)
    ng = bm._lib.scan_plan(L, D, B * ndir, True)[1]
    # at least: the activation-sized gradients + the scan's partial rows
    floor = M * D * es + 4 * M * ndir * D * es + M * ndir * 48 * es + M * 2 * D * es + B * ng * L * ndir * 2 * N * 4
    assert floor < bwd < 2 * floor
    assert bwd % 256 == 0
    # errors are codes
    assert lib.bimamba_block_fwd(None, None) == -1
    d = bm._lib.BlockDesc()
    d.batch, d.seqlen, d.d_model, d.d_inner, d.dt_rank, d.d_conv, d.ndir = 2, 5, 144, 288, 9, 4, 2
    d.io_dtype = bm._lib.F32
    assert lib.bimamba_block_fwd(C.byref(d), None) == -6                   # fp32 goes through the op-level entries
    assert b"bf16 or fp16" in lib.bimamba_last_error()
    d.io_dtype = bm._lib.BF16
    d.d_model = 20
    assert lib.bimamba_block_fwd(C.byref(d), None) == -7                   # d_model must be a multiple of 8
    d.d_model = 144
    assert lib.bimamba_block_fwd(C.byref(d), None) == -7                   # null operands
    d.ndir = 3
    assert lib.bimamba_block_fwd(C.byref(d), None) == -3
    d.ndir = 2
    assert lib.bimamba_block_bwd(C.byref(d), None, None) < 0


def test_time_split_plan_and_errors_host_side():
    """bimamba_scan_fwd_split_plan: only long walks on an under-filled GPU split; segments are whole 16-step chunks and the
    last one is never empty; the knob forces a segment count for the parity tests; argument errors are codes."""
    lib = bm._lib.load()
    assert bm._lib.scan_split_plan(64, 1, 8192, 288) == (3, 2736)       # config 5, last point: 576 warps -> 3 segments
    assert bm._lib.scan_split_plan(1, 1, 8192, 288) == (8, 1024)
    assert bm._lib.scan_split_plan(64, 1, 8192, 288, bm._lib.F32) == (1, 8192)   # fp32 I/O: the serial walk was measured faster
    assert bm._lib.scan_split_plan(64, 1, 8192, 288, bm._lib.F16) == (3, 2736)
    assert bm._lib.scan_split_plan(64, 2, 201, 288) == (1, 201)         # the Phase-6 training shape never splits
    assert bm._lib.scan_split_plan(128, 1, 4096, 288) == (1, 4096)      # 1152 warps fill the GPU
    assert bm._lib.scan_split_plan(8, 1, 1024, 288) == (1, 1024)        # short walk
    with bm._lib.tuning(bm._lib.TUNE_SCAN_SPLIT, 4):
        assert bm._lib.scan_split_plan(2, 2, 201, 288) == (4, 64)
        assert bm._lib.scan_split_plan(2, 2, 37, 40) == (3, 16)
        assert bm._lib.scan_split_plan(1, 1, 1, 8) == (1, 1)
    with bm._lib.tuning(bm._lib.TUNE_SCAN_SPLIT, 1):
        assert bm._lib.scan_split_plan(64, 1, 8192, 288) == (1, 8192)
    for L in (17, 201, 499, 8192, 10000):
        for force in (2, 3, 5, 8):
            with bm._lib.tuning(bm._lib.TUNE_SCAN_SPLIT, force):
                ns, sl = bm._lib.scan_split_plan(2, 1, L, 64)
            assert ns == 1 or (sl % 16 == 0 and (ns - 1) * sl < L <= ns * sl), (L, force, ns, sl)
    assert lib.bimamba_scan_fwd_split_workspace_bytes(64, 1, 288, 3) == 64 * 2 * 288 * 17 * 4
    assert lib.bimamba_scan_fwd_split_workspace_bytes(64, 1, 288, 1) == 0
    # validation happens before any launch: a descriptor with fake (non-null) pointers and a bad segmentation
    d = bm._lib.ScanDesc()
    d.batch, d.ndir, d.dim, d.seqlen, d.dstate, d.group_channels = 2, 1, 32, 100, 16, 32
    d.u = d.A = d.bc = d.delta = d.out = 4096
    assert lib.bimamba_selective_scan_fwd_split(C.byref(d), 3, 40, None, 0, None) == -5       # not whole chunks
    assert lib.bimamba_selective_scan_fwd_split(C.byref(d), 3, 64, None, 0, None) == -5       # last segment empty
    assert lib.bimamba_selective_scan_fwd_split(C.byref(d), 2, 64, None, 0, None) == -10      # no carry workspace
    assert b"carry workspace" in lib.bimamba_last_error()
    assert lib.bimamba_selective_scan_fwd_split(None, 2, 64, None, 0, None) == -1


def test_native_layer_entry_points_host_side():
    """bimamba_layer_fwd / bimamba_layer_bwd (the whole encoder layer in one call each way): the workspaces contain the
    block's, argument errors are codes."""
    lib = bm._lib.load()
    B, L, dm, D, dff, ndir, K = 64, 201, 144, 288, 576, 2, 4
    up = lambda v: (v + 255) // 256 * 256
    M, es = B * L, 2
    own = 3 * up(M * dm * es) + 2 * up(M * dff * es) + up(M * 16) + 4 * up(dff * dm * es)
    for save in (0, 1):
        assert lib.bimamba_layer_fwd_workspace_bytes(B, L, dm, D, dff, ndir, bm._lib.BF16, save) == (
            own + lib.bimamba_block_fwd_workspace_bytes(B, L, dm, D, ndir, bm._lib.BF16, save))
    bwd = lib.bimamba_layer_bwd_workspace_bytes(B, L, dm, D, dff, K, ndir, bm._lib.BF16)
    assert bwd > lib.bimamba_block_bwd_workspace_bytes(B, L, dm, D, K, ndir, bm._lib.BF16) + 4 * M * dm * es + 2 * M * dff * es
    assert bwd % 256 == 0 and lib.bimamba_layer_bwd_workspace_bytes(0, L, dm, D, dff, K, ndir, bm._lib.BF16) == 0
    assert lib.bimamba_layer_fwd(None, None) == -1
    d = bm._lib.LayerDesc()
    k = d.block
    k.batch, k.seqlen, k.d_model, k.d_inner, k.dt_rank, k.d_conv, k.ndir, k.io_dtype = 2, 5, 144, 288, 9, 4, 2, bm._lib.BF16
    d.d_ff, d.x_dtype = 576, bm._lib.F16
    assert lib.bimamba_layer_fwd(C.byref(d), None) == -6                   # x must be fp32 or the compute dtype
    d.x_dtype = bm._lib.F32
    assert lib.bimamba_layer_fwd(C.byref(d), None) == -7                   # null operands
    d.d_ff = 100
    assert lib.bimamba_layer_fwd(C.byref(d), None) == -7                   # d_ff must be a multiple of 8
    assert lib.bimamba_layer_bwd(C.byref(d), None, None) < 0


def test_header_is_plain_c_and_the_integration_example_compiles(tmp_path):
    """include/bimamba.h is C99 (no C++ leaks into the boundary) and the host-side example of INTEGRATION.md section 4 - a C
    program that fills the block / layer descriptors and calls the one-call entry points - compiles against it and links
    against the library (gcc only; nothing is executed on a device)."""
    prog = tmp_path / "host.c"
    prog.write_text(f"""#include <stdio.h>
#include "{HEADER}"
int run_block(void* stream, const void* x, void* out, void* ws, size_t ws_bytes) {{
  bimamba_block_desc d = {{0}};
  d.x = x; d.out = out;
  d.batch = 64; d.seqlen = 201; d.d_model = 144; d.d_inner = 288; d.dt_rank = 9; d.d_conv = 4; d.ndir = 2;
  d.io_dtype = BIMAMBA_BF16; d.save_for_backward = 1;
  d.workspace = ws; d.workspace_bytes = ws_bytes;
  if (ws_bytes < bimamba_block_fwd_workspace_bytes(64, 201, 144, 288, 2, BIMAMBA_BF16, 1)) return -10;
  int rc = bimamba_block_fwd(&d, stream);
  if (rc) fprintf(stderr, "%s\\n", bimamba_last_error());
  bimamba_block_grads g = {{0}};
  g.workspace_bytes = bimamba_block_bwd_workspace_bytes(64, 201, 144, 288, 4, 2, BIMAMBA_BF16);
  return rc ? rc : bimamba_block_bwd(&d, &g, stream);
}}
int run_layer(void* stream, const float* x, float* out) {{
  bimamba_layer_desc l = {{0}};
  l.x = x; l.out = out; l.eps1 = l.eps2 = 1e-5f; l.d_ff = 576; l.x_dtype = BIMAMBA_F32;
  l.block.batch = 8; l.block.seqlen = 201; l.block.d_model = 144; l.block.d_inner = 288; l.block.dt_rank = 9;
  l.block.d_conv = 4; l.block.ndir = 2; l.block.io_dtype = BIMAMBA_F16; l.block.save_for_backward = 1;
  l.workspace_bytes = bimamba_layer_fwd_workspace_bytes(8, 201, 144, 288, 576, 2, BIMAMBA_F16, 1);
  bimamba_layer_grads g = {{0}};
  g.workspace_bytes = bimamba_layer_bwd_workspace_bytes(8, 201, 144, 288, 576, 4, 2, BIMAMBA_F16);
  int rc = bimamba_layer_fwd(&l, stream);
  return rc ? rc : bimamba_layer_bwd(&l, &g, stream);
}}
int main(void) {{
  int nseg = 0, seg_len = 0;
  bimamba_scan_fwd_split_plan(64, 1, 8192, 288, BIMAMBA_BF16, &nseg, &seg_len);
  printf("%d %d %d %d\\n", bimamba_abi_version(), nseg, seg_len, run_layer(0, 0, 0));
  return 0;
}}
""")
    exe = tmp_path / "host"
    lib = bm._lib.LIB_PATH
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", str(prog), "-o", str(exe), lib,
                           "-Wl,-rpath," + os.path.dirname(lib)])
    # argument validation only (null operands -> error code before any CUDA call): safe without a GPU
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == bm._lib.ABI_VERSION and (int(out[1]), int(out[2])) == (3, 2736) and int(out[3]) < 0
