"""Host-side logic of the Phase-6 training-loop pieces (robust-audio-deepfake-evolution_b200/training.py) and
of the config-3 harness scaffolding (tools/phase6_model.py) - CPU only, no CUDA kernels involved."""
import copy
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
tr = importlib.import_module("robust-audio-deepfake-evolution_b200.training")


class _Tiny(nn.Module):
    """(features, logits) model with a parameter named like the FGM target."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.feature_projection = nn.Linear(6, 8)
        self.body = nn.Linear(8, 5)
        self.classifier = nn.Linear(5, 2)

    def forward(self, x):
        f = torch.tanh(self.body(torch.tanh(self.feature_projection(x))))
        return f, self.classifier(f)


def _ce(out, feats, y):
    return nn.functional.cross_entropy(out.float(), y, weight=torch.tensor([0.1, 0.9]))


def test_sincnet_stand_in_matches_reference_stream(golden_dir):
    """tools/phase6_model.SincNetStream against the reference's SincNetEncoder run (fixture: waveform, weights, f_sinc)."""
    import phase6_model as pm
    g = dict(np.load(os.path.join(golden_dir, "model_tail_linear.npz")))
    st = pm.SincNetStream().eval()
    st.load_state_dict({k[len("sinc."):]: torch.tensor(v) for k, v in g.items() if k.startswith("sinc.")}, strict=True)
    with torch.no_grad():
        out = st(torch.tensor(g["wav"]))
    assert out.shape == g["f_sinc"].shape
    assert float((out.double() - torch.tensor(g["f_sinc"])).abs().max()) <= 1e-5 * float(np.abs(g["f_sinc"]).max())


def test_fgm_attack_and_restore():
    m = _Tiny()
    x = torch.randn(4, 6)
    _ce(m(x)[1], None, torch.tensor([0, 1, 1, 0])).backward()
    w0 = m.feature_projection.weight.detach().clone()
    b0 = m.body.weight.detach().clone()
    fgm = tr.FGM(m, "feature_projection", epsilon=0.5)
    fgm.attack()
    gw = m.feature_projection.weight.grad
    assert torch.allclose(m.feature_projection.weight, w0 + 0.5 * gw / gw.norm(), atol=1e-7)      # main.py:93-94
    assert torch.equal(m.body.weight, b0)                                                         # only the named params
    fgm.restore()
    assert torch.equal(m.feature_projection.weight, w0) and fgm.backup == {}
    m.feature_projection.weight.grad.zero_()                                                      # zero gradient: no move, no NaN
    fgm.attack()
    assert torch.equal(m.feature_projection.weight, w0)
    fgm.restore()


def test_mixup_batch_and_loss():
    g = torch.Generator().manual_seed(0)
    x = torch.arange(12.0).view(4, 3)
    y = torch.tensor([0, 1, 0, 1])
    mx, ya, yb, lam = tr.mixup_batch(x, y, alpha=1.0, generator=g, lam=0.3)
    perm = torch.randperm(4, generator=torch.Generator().manual_seed(0))
    assert torch.allclose(mx, 0.3 * x + 0.7 * x[perm]) and torch.equal(yb, y[perm]) and torch.equal(ya, y) and lam == 0.3
    one = tr.mixup_batch(x[:1], y[:1], alpha=1.0)
    assert one[3] == 1.0 and torch.equal(one[0], x[:1])                                           # main.py:1038 batch_size > 1
    out = torch.randn(4, 2)
    l = tr.mixup_loss(_ce, out, None, ya, yb, 0.3)
    assert torch.allclose(l, 0.3 * _ce(out, None, ya) + 0.7 * _ce(out, None, yb))


def test_minimal_lora():
    m = nn.ModuleDict({"attn": nn.ModuleDict({"q_proj": nn.Linear(8, 8), "k_proj": nn.Linear(8, 8), "v_proj": nn.Linear(8, 8)})})
    x = torch.randn(3, 8)
    before = m["attn"]["q_proj"](x)
    n = tr.apply_lora(m, ("q_proj", "v_proj"), r=2, alpha=8, dropout=0.0)
    assert n == 2 and isinstance(m["attn"]["q_proj"], tr.LoRALinear) and isinstance(m["attn"]["k_proj"], nn.Linear)
    assert torch.allclose(m["attn"]["q_proj"](x), before)                                         # B = 0: identity update
    trainable = [k for k, p in m.named_parameters() if p.requires_grad]
    assert sorted(trainable) == ["attn.q_proj.lora_A.weight", "attn.q_proj.lora_B.weight",
                                 "attn.v_proj.lora_A.weight", "attn.v_proj.lora_B.weight"]
    m["attn"]["q_proj"](x).sum().backward()
    assert m["attn"]["q_proj"].lora_B.weight.grad.abs().sum() > 0 and m["attn"]["q_proj"].base.weight.grad is None


def test_phase6_train_step_equals_reference_loop():
    """Phase6TrainStep against a straight restatement of train_epoch's body (src/main.py:1031-1116): two micro-batches of
    gradient accumulation, Mixup, FGM's second backward, clip at 3.0, AdamW, EMA - same parameters afterwards."""
    gen = torch.Generator().manual_seed(5)
    batches = [(torch.randn(4, 6, generator=gen), torch.randint(0, 2, (4,), generator=gen)) for _ in range(2)]
    lam = 0.4

    a = _Tiny()
    opt_a = torch.optim.AdamW(a.parameters(), lr=1e-2, weight_decay=1e-4)
    ema_a = torch.optim.swa_utils.AveragedModel(a, multi_avg_fn=torch.optim.swa_utils.get_ema_multi_avg_fn(0.999))
    step = tr.Phase6TrainStep(a, opt_a, _ce, autocast_dtype=None, fgm=tr.FGM(a, "feature_projection", 0.5), mixup_alpha=1.0,
                              accumulation_steps=2, ema_model=ema_a)
    loss = step(batches, generator=torch.Generator().manual_seed(9), lam=lam)
    assert loss.ndim == 0 and torch.isfinite(loss)

    b = _Tiny()
    opt_b = torch.optim.AdamW(b.parameters(), lr=1e-2, weight_decay=1e-4)
    ema_b = torch.optim.swa_utils.AveragedModel(b, multi_avg_fn=torch.optim.swa_utils.get_ema_multi_avg_fn(0.999))
    g2 = torch.Generator().manual_seed(9)
    opt_b.zero_grad()
    for x, y in batches:
        idx = torch.randperm(4, generator=g2)
        mx, ya, yb = lam * x + (1 - lam) * x[idx], y, y[idx]
        out = b(mx)[1]
        ((lam * _ce(out, None, ya) + (1 - lam) * _ce(out, None, yb)) / 2).backward()
        w = b.feature_projection.weight
        bias = b.feature_projection.bias
        keep = (w.data.clone(), bias.data.clone())
        w.data.add_(0.5 * w.grad / w.grad.norm())
        bias.data.add_(0.5 * bias.grad / bias.grad.norm())
        out = b(mx)[1]
        ((lam * _ce(out, None, ya) + (1 - lam) * _ce(out, None, yb)) / 2).backward()
        w.data, bias.data = keep
    torch.nn.utils.clip_grad_norm_(b.parameters(), 3.0)
    opt_b.step()
    ema_b.update_parameters(b)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, atol=1e-7)
    for p, q in zip(ema_a.parameters(), ema_b.parameters()):
        assert torch.allclose(p, q, atol=1e-7)
    assert all(p.grad is None for p in a.parameters())                                            # zeroed after the step
