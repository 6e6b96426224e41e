"""`PN_BiMambas_Encoder` and the 4-layer backend with the reference's attribute names
(src/models/DualStreamSEMamba.py:445-486, :697-710, :755-767), on top of the fused Bi-Mamba block."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .mamba_simple import Mamba
from .ops import encoder_layer_native, feed_forward_fn, head_fwd, head_pool_fn, layer_norm_fn, native_layer_ok


class PN_BiMambas_Encoder(nn.Module):
    """Pre-norm Bi-Mamba layer.  Same constructor, attributes (`mamba`, `norm1`, `norm2`,
    `feed_forward`) and state_dict keys as DualStreamSEMamba.py:451-465."""

    def __init__(self, d_model, n_state):
        super().__init__()
        self.d_model = d_model
        self.mamba = Mamba(d_model, n_state)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.feed_forward = nn.Sequential(
            nn.Linear(d_model, d_model * 4),
            nn.GELU(),
            nn.Linear(d_model * 4, d_model),
        )

    def forward(self, x):
        if native_layer_ok(x, self):
            # eager mode: the whole layer in one native call each way (same kernels, same order, same bits; a captured
            # step keeps the sequenced Functions below, which overlap the weight-gradient products on a second stream)
            return encoder_layer_native(x, self)
        # :471-472; the residual branch leaves the LayerNorm Function as its second output, so its gradient is added to dx
        # inside the LayerNorm backward kernel (no separate autograd add)
        x_norm, residual = layer_norm_fn(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, with_residual=True)
        mamba_out = self.mamba.forward_bidirectional(x_norm)                                         # :473-481 in one fused pass
        mamba_out = layer_norm_fn(mamba_out, self.norm2.weight, self.norm2.bias, self.norm2.eps)    # :482
        if isinstance(self.feed_forward[1], nn.GELU) and self.feed_forward[1].approximate == "none":
            l1, l2 = self.feed_forward[0], self.feed_forward[2]
            return feed_forward_fn(mamba_out, l1.weight, l1.bias, l2.weight, l2.bias, residual=residual)   # :483-485
        return self.feed_forward(mamba_out) + residual


class BiMambaBackend(nn.Module):
    """The part of the reference `Model` after fusion (DualStreamSEMamba.py:697-710, :755-767):
    `backbone_layers`, `norm_f`, `attention_pool`, `dropout`, `classifier` with the same names."""

    def __init__(self, emb_size=144, num_encoders=4, d_state=16):
        super().__init__()
        self.backbone_layers = nn.ModuleList(
            [PN_BiMambas_Encoder(d_model=emb_size, n_state=d_state) for _ in range(num_encoders)])
        self.norm_f = nn.LayerNorm(emb_size)
        self.attention_pool = nn.Linear(emb_size, 1)
        self.dropout = nn.Dropout(0.1)
        self.classifier = nn.Linear(emb_size, 2)

    def forward_features(self, f_fused):
        for layer in self.backbone_layers:
            f_fused = layer(f_fused)
        return f_fused

    def forward(self, f_fused):
        return backend_head(self, self.forward_features(f_fused))


def backend_head(m, f_fused):
    """norm_f -> attention pooling -> dropout -> classifier on a module that carries the reference's attribute names
    (`norm_f`, `attention_pool`, `dropout`, `classifier`; DualStreamSEMamba.py:759-767).  Returns (features, logits)."""
    if not torch.is_grad_enabled() and not m.training and f_fused.shape[-1] <= 256:
        # scoring (src/main.py:958-995): norm_f, attention pooling and the classifier in one launch
        feats, logits = head_fwd(f_fused, m.norm_f.weight, m.norm_f.bias, m.attention_pool.weight,
                                 m.attention_pool.bias, m.classifier.weight, m.classifier.bias, m.norm_f.eps)
        return feats.to(f_fused.dtype), logits.to(f_fused.dtype)
    if f_fused.shape[-1] <= 256:
        # training: norm_f + attention pooling as one kernel forward and one backward (:759-763)
        features = head_pool_fn(f_fused, m.norm_f.weight, m.norm_f.bias, m.attention_pool.weight,
                                m.attention_pool.bias, m.norm_f.eps).to(f_fused.dtype)
    else:
        f_fused = layer_norm_fn(f_fused, m.norm_f.weight, m.norm_f.bias, m.norm_f.eps, out_dtype=f_fused.dtype)   # :759
        attn = F.softmax(m.attention_pool(f_fused), dim=1)                          # :762
        features = torch.matmul(attn.transpose(1, 2), f_fused).squeeze(1)           # :763
    features = m.dropout(features)                                                  # :764
    return features, m.classifier(features)                                         # :767
