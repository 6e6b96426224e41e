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
  * Model                 /root/reference/src/models/DualStreamSEMamba.py:643-769 with its WavLM frontend replaced
    by a stub that returns preset (B, T, 1024) features (the pretrained checkpoint is not shipped and there is no
    network): SincNetEncoder, DualStreamFusion, the 4 backbone layers, norm_f, attention pooling and the
    classifier are the reference's own objects, run through the reference's own Model.forward
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


def _model_tail_fixtures():
    """The reference Model (DualStreamSEMamba.py:643-769) from the fused features on: fusion (:537-637), 4 backbone
    layers (:697-700, :755-756), norm_f, attention pooling, classifier (:759-767), eval mode, fp64."""
    import contextlib
    import io

    import models.DualStreamSEMamba as ds

    class _StubFrontend(torch.nn.Module):
        """Stands in for WavLMFrontend (DualStreamSEMamba.py:276-437): returns preset hidden features."""
        def __init__(self, freeze_layers=18):
            super().__init__()
            self.out_dim = 1024
            self.feats = None

        def forward(self, x):
            return self.feats

    ds.WavLMFrontend = _StubFrontend
    for tag, T1, n_samples, seed in (("nearest", 201, 64600, 21), ("linear", 60, 32000, 22)):
        torch.manual_seed(seed)
        args = types.SimpleNamespace(emb_size=144, num_encoders=4 if tag == "nearest" else 1, d_state=16, sinc_channels=70)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ds.Model(args=args, device="cpu").double().eval()
        _perturb(model, seed)
        g = torch.Generator().manual_seed(seed)
        wav = 0.1 * torch.randn(2, n_samples, generator=g, dtype=torch.float64)
        f_wavlm = torch.randn(2, T1, 1024, generator=g, dtype=torch.float64).float().double().requires_grad_(True)
        model.wavlm_stream.feats = f_wavlm
        captured = {}
        # the SincNet stream builds its band-pass filters in fp32 on the fly (DualStreamSEMamba.py:119-141): it runs in
        # fp32 and its output enters the fp64 tail as the fixture's f_sinc
        model.sinc_stream.float()
        model.sinc_stream.register_forward_pre_hook(lambda m, a: (a[0].float(),) + tuple(a[1:]))

        def _sinc_out(m, i, o):
            o = o.detach().double().requires_grad_(True)
            captured["f_sinc"] = o
            return o
        h1 = model.sinc_stream.register_forward_hook(_sinc_out)
        h2 = model.fusion.register_forward_hook(lambda m, i, o: captured.__setitem__("f_fused", o))
        feats, logits = model(wav)
        h1.remove()
        h2.remove()
        f_sinc, f_fused = captured["f_sinc"], captured["f_fused"]
        f_fused.retain_grad()
        cot = torch.randn(logits.shape, generator=g, dtype=torch.float64)
        cot_f = torch.randn(feats.shape, generator=g, dtype=torch.float64)
        ((logits * cot).sum() + (feats * cot_f).sum()).backward()
        rec = {"f_wavlm": f_wavlm.detach().numpy().astype(np.float32), "f_sinc": f_sinc.detach().numpy(),
               "f_fused": f_fused.detach().numpy(), "features": feats.detach().numpy(), "logits": logits.detach().numpy(),
               "cot_logits": cot.numpy(), "cot_features": cot_f.numpy(),
               "grad.f_wavlm": f_wavlm.grad.numpy().astype(np.float32), "grad.f_sinc": f_sinc.grad.numpy(),
               "grad.f_fused": f_fused.grad.numpy()}
        # nearest (the Phase-6 shape): everything after the two streams; backbone gradients are stored as every 5th
        # element of the flattened tensor (the fixture would otherwise double in size).  linear: the fusion block only -
        # grad.f_fused is the cotangent that reproduces its gradients, so no backbone weights are needed.
        for name, p in model.named_parameters():
            if name.startswith(("fusion.", "norm_f.", "attention_pool.", "classifier.")):
                rec["param." + name] = p.detach().numpy().astype(np.float32)
                rec["grad." + name] = p.grad.numpy().astype(np.float32)
            elif name.startswith("backbone_layers.") and tag == "nearest":
                rec["param." + name] = p.detach().numpy().astype(np.float32)
                rec["grad5." + name] = p.grad.numpy().reshape(-1)[::5].astype(np.float32)
        if tag == "nearest":
            rec["grad5.f_wavlm"] = rec.pop("grad.f_wavlm").reshape(-1)[::5]
        else:
            for k in ("features", "logits", "cot_logits", "cot_features"):
                rec.pop(k)
            # the SincNet stream (DualStreamSEMamba.py:206-273, eval mode, fp32) of this case: waveform and weights, so
            # the harness's stand-in stream (tools/phase6_model.py) can be checked against f_sinc
            rec["wav"] = wav.numpy().astype(np.float32)
            for name, t in model.sinc_stream.state_dict().items():
                rec["sinc." + name] = t.detach().numpy()
        np.savez_compressed(os.path.join(HERE, f"model_tail_{tag}.npz"), **rec)
        print("model tail", tag, "f_sinc", tuple(f_sinc.shape), "f_fused", tuple(f_fused.shape), "logits", logits.detach().numpy())


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

    # ---- 2b. the reference Model from the fused features on (fusion + backbone + head) ----
    _model_tail_fixtures()

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
