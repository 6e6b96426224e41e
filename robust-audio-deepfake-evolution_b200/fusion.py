"""`DualStreamFusion` and `SELayer` with the reference's constructor, attributes and state_dict keys
(src/models/DualStreamSEMamba.py:492-531, :537-637): the block that turns the WavLM stream (B, T1, 1024) and the SincNet
stream (B, T2, 64) into the (B, T1, 144) features the Bi-Mamba backend reads.

B200 path: both input LayerNorms and the final one are this repository's kernels; the three Linear layers run on the
tcgen05 GEMM with bias / addend epilogues (forward, data gradient, MN-major weight gradient).  Two identities remove
the (B, T1, 288) concatenation and move the interpolation to the short side:
  fusion_proj(cat[f_w, f_s]) = f_w Wa^T + f_s Wb^T + b          (Wa, Wb = the two column halves of fusion_proj.weight)
  interp_T(f_s) Wb^T = interp_T(f_s Wb^T)                        (interpolation acts on time, the projection on channels)
so the SincNet branch is projected at its own T2 = 29 frames and enters the WavLM-side GEMM as its addend.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import layer_norm_fn, linear_fn


class SELayer(nn.Module):
    """Squeeze-and-excitation over (B, T, C) sequences, DualStreamSEMamba.py:509-531."""

    def __init__(self, channel: int, reduction: int = 16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool1d(1)
        self.fc = nn.Sequential(
            nn.Linear(channel, channel // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channel // reduction, channel, bias=False),
            nn.Sigmoid(),
        )

    def gate(self, x):
        """(B, T, C) -> (B, 1, C) channel weights (fp32 statistics)."""
        y = x.float().mean(dim=1)                                                   # :524-526
        return self.fc(y).unsqueeze(1)                                              # :527

    def forward(self, x):
        return x * self.gate(x).to(x.dtype)                                         # :528


class DualStreamFusion(nn.Module):
    def __init__(self, wavlm_dim: int, sinc_dim: int, out_dim: int, reduction: int = 16):
        super().__init__()
        self.ln_wavlm = nn.LayerNorm(wavlm_dim)
        self.ln_sinc = nn.LayerNorm(sinc_dim)
        self.wavlm_proj = nn.Linear(wavlm_dim, out_dim)
        self.sinc_proj = nn.Linear(sinc_dim, out_dim)
        self.fusion_proj = nn.Linear(out_dim * 2, out_dim)
        self.se_layer = SELayer(out_dim, reduction=reduction)
        self.norm = nn.LayerNorm(out_dim)
        self.dropout = nn.Dropout(0.1)

    def forward(self, f_wavlm, f_sinc):
        D = self.wavlm_proj.out_features
        T1, T2 = f_wavlm.shape[1], f_sinc.shape[1]
        res_dtype = f_wavlm.dtype if f_wavlm.dtype == torch.float32 or not torch.is_autocast_enabled("cuda") else torch.float32
        fw = layer_norm_fn(f_wavlm, self.ln_wavlm.weight, self.ln_wavlm.bias, self.ln_wavlm.eps)        # :591
        fs = layer_norm_fn(f_sinc, self.ln_sinc.weight, self.ln_sinc.bias, self.ln_sinc.eps)            # :592
        f_w = linear_fn(fw, self.wavlm_proj.weight, self.wavlm_proj.bias)                               # :595
        f_s = linear_fn(fs, self.sinc_proj.weight, self.sinc_proj.bias)                                 # :596
        Wa, Wb = self.fusion_proj.weight[:, :D], self.fusion_proj.weight[:, D:]
        s2 = linear_fn(f_s, Wb)                                                     # SincNet half of :630, at T2 frames
        if T2 != T1:                                                                # :601-626
            mode = "nearest" if T1 / T2 > 4.0 else "linear"
            kw = {} if mode == "nearest" else {"align_corners": False}
            s2 = F.interpolate(s2.transpose(1, 2), size=T1, mode=mode, **kw).transpose(1, 2).contiguous()
        fused = linear_fn(f_w, Wa, self.fusion_proj.bias, addend=s2)                # :629-630
        fused = fused * self.se_layer.gate(fused).to(fused.dtype)                   # :633
        fused = layer_norm_fn(fused, self.norm.weight, self.norm.bias, self.norm.eps, out_dtype=res_dtype)   # :636
        return self.dropout(fused)                                                  # :637
