"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own classes by path:
  * MambaBlock            /root/reference/src/models/modules/mamba_block.py:6-122
  * PN_BiMambas_Encoder   /root/reference/src/models/DualStreamSEMamba.py:445-486
    (with `mamba_ssm.modules.mamba_simple.Mamba` bound to MambaBlock - the same
    module-injection trick as the reference's utils/check_model.py:6-23, but bound
    to the real math instead of an identity mock)
  * compute_eer           /root/reference/src/evaluation.py:154-160
runs them in fp64 on seeded inputs and stores inputs, weights, outputs and
gradients as .npz.  The fixtures pin oracle/bimamba_oracle.py
(tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, os.path.join(REF, "src", "models", "modules"))
    from mamba_block import MambaBlock  # noqa: E402

    pkg = types.ModuleType("mamba_ssm")
    mods = types.ModuleType("mamba_ssm.modules")
    simple = types.ModuleType("mamba_ssm.modules.mamba_simple")
    simple.Mamba = MambaBlock
    pkg.modules = mods
    mods.mamba_simple = simple
    sys.modules["mamba_ssm"] = pkg
    sys.modules["mamba_ssm.modules"] = mods
    sys.modules["mamba_ssm.modules.mamba_simple"] = simple
    sys.path.insert(0, os.path.join(REF, "src"))
    from models.DualStreamSEMamba import PN_BiMambas_Encoder  # noqa: E402
    from evaluation import compute_eer  # noqa: E402
    return MambaBlock, PN_BiMambas_Encoder, compute_eer


def _perturb(module, seed):
    """Move every parameter off its init value (trained-like), seeded."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g, dtype=torch.float64).to(p.dtype))
            p.copy_(p.float().double())  # fp32-representable, so fp32 storage below is lossless


def main():
    MambaBlock, PN_BiMambas_Encoder, compute_eer = _import_reference()
    torch.manual_seed(1234)  # reference default seed, src/main.py:1145

    # ---- 1. one MambaBlock, small width, outputs + every gradient ----
    for tag, d_model, Bsz, L in (("small", 32, 2, 19), ("phase6", 144, 2, 13)):
        blk = MambaBlock(d_model, 16).double()
        _perturb(blk, 11)
        x = torch.randn(Bsz, L, d_model).double().requires_grad_(True)
        cot = torch.randn(Bsz, L, d_model).double()
        out = blk(x)
        (out * cot).sum().backward()
        rec = {"x": x.detach().numpy().astype(np.float32), "cot": cot.numpy().astype(np.float32),
               "out": out.detach().numpy(), "grad.x": x.grad.numpy()}
        for name, p in blk.named_parameters():
            rec["param." + name] = p.detach().numpy().astype(np.float32)   # lossless (see _perturb)
            rec["grad." + name] = p.grad.numpy().astype(np.float32)        # 6e-8 relative rounding
        np.savez_compressed(os.path.join(HERE, f"mamba_block_{tag}.npz"), **rec)
        print(tag, "MambaBlock out", out.shape, float(out.detach().abs().max()))

    # ---- 2. PN_BiMambas_Encoder (bidirectional, shared weights), Phase-6 width ----
    enc = PN_BiMambas_Encoder(144, 16).double()
    _perturb(enc, 12)
    x = torch.randn(2, 17, 144).double().requires_grad_(True)
    cot = torch.randn(2, 17, 144).double()
    out = enc(x)
    (out * cot).sum().backward()
    rec = {"x": x.detach().numpy().astype(np.float32), "cot": cot.numpy().astype(np.float32),
           "out": out.detach().numpy(), "grad.x": x.grad.numpy()}
    for name, p in enc.named_parameters():
        rec["param." + name] = p.detach().numpy().astype(np.float32)
        rec["grad." + name] = p.grad.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "pn_bimamba_encoder_phase6.npz"), **rec)
    print("encoder out", out.shape, float(out.detach().abs().max()))

    # ---- 3. compute_eer known answers on seeded score sets ----
    rng = np.random.RandomState(1234)
    cases = {}
    for i, (nt, nn_, shift) in enumerate(((700, 6000, 1.5), (50, 80, 0.3), (1000, 1000, 3.0))):
        tgt = rng.randn(nt) + shift
        non = rng.randn(nn_)
        if i == 1:  # ties
            tgt = np.round(tgt, 1)
            non = np.round(non, 1)
        eer, thr = compute_eer(tgt, non)
        cases[f"tgt{i}"] = tgt
        cases[f"non{i}"] = non
        cases[f"eer{i}"] = np.float64(eer)
        cases[f"thr{i}"] = np.float64(thr)
        print("eer case", i, eer, thr)
    np.savez_compressed(os.path.join(HERE, "compute_eer_cases.npz"), **cases)

    # The two official baseline CM score files give:
    #   B01_LA_primary_eval.txt -> 9.572028207 %, thr 2.909863
    #   B02_LA_primary_eval.txt -> 8.089825328 %, thr 1.030046
    # They are 71 237 lines each and are not copied into this repo; re-derive them here
    # as a check on the imported function.
    for fname, want in (("B01_LA_primary_eval.txt", 9.572028207), ("B02_LA_primary_eval.txt", 8.089825328)):
        path = os.path.join(REF, "tDCF_python_v2", "scores", fname)
        if os.path.exists(path):
            data = np.genfromtxt(path, dtype=str)
            keys = data[:, 4]
            scores = data[:, 5].astype(np.float64)
            eer, thr = compute_eer(scores[keys == "bonafide"], scores[keys == "spoof"])
            print(fname, eer * 100, thr, "expected", want)


if __name__ == "__main__":
    main()
