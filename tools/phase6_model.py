"""Harness scaffolding for BASELINE.json config 3 (the full Phase-6 dual-stream detector step): stand-ins for the two
front-end streams, which are OUT of the hot-path scope but are needed around it to time a whole
training step.  NOT product code: the product is `bimamba_b200` (fusion + Bi-Mamba backend + head), used unchanged here.

  * RandomInitWavLMFrontend - the role of WavLMFrontend (DualStreamSEMamba.py:276-437): a WavLM-Large-SHAPED encoder built
    from `transformers.WavLMConfig` with random weights (the pretrained checkpoint is not shipped and there is no network),
    all 25 hidden states, learnable softmax layer weights (:420-435).
  * SincNetStream - the role of SincNetEncoder (DualStreamSEMamba.py:206-273): fixed mel-spaced sinc band-pass bank,
    |.| + 3x3 max-pool, BatchNorm + SELU, six residual 2-D conv blocks, max over the frequency axis.  Same attribute
    names (state_dict keys `first_bn.*`, `encoder.{i}.0.*`), so reference checkpoints load; checked against the
    reference's own stream through tests/golden/model_tail_linear.npz (tests/test_training_cpu.py).
  * Phase6Model - the reference Model's wiring (DualStreamSEMamba.py:728-769) with the reference's attribute names
    (`wavlm_stream`, `sinc_stream`, `fusion`, `backbone_layers`, `norm_f`, `attention_pool`, `dropout`, `classifier`).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def mel_sinc_bank(n_filters: int = 70, kernel: int = 129, sample_rate: int = 16000) -> torch.Tensor:
    """(n_filters, kernel) Hamming-windowed ideal band-pass filters whose edges are equally spaced on the mel scale
    between 0 and sample_rate / 2 (the fixed, non-learned bank of DualStreamSEMamba.py:94-117)."""
    to_mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    to_hz = lambda mel: 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    grid = to_mel(int(sample_rate / 2) * np.linspace(0, 1, 257))
    edges = to_hz(np.linspace(grid.min(), grid.max(), n_filters + 1))
    n = torch.arange(-(kernel - 1) / 2, (kernel - 1) / 2 + 1)                 # float32 support, as the reference's hsupp
    bank = torch.zeros(n_filters, kernel)
    win = torch.tensor(np.hamming(kernel), dtype=torch.float32)
    for i in range(n_filters):
        lo, hi = edges[i], edges[i + 1]
        ideal = (2 * hi / sample_rate) * np.sinc(2 * hi * n / sample_rate) - (2 * lo / sample_rate) * np.sinc(2 * lo * n / sample_rate)
        bank[i] = win * torch.as_tensor(ideal, dtype=torch.float32)
    return bank


class _ResBlock(nn.Module):
    """Residual 2-D block of the SincNet stream (DualStreamSEMamba.py:150-203).  As in the reference, conv1 reads the
    block INPUT: the bn1 + SELU result of non-first blocks is computed and discarded there (:185-190), so bn1 only exists
    for its parameters / running statistics."""

    def __init__(self, c_in: int, c_out: int, first: bool = False):
        super().__init__()
        self.first = first
        if not first:
            self.bn1 = nn.BatchNorm2d(c_in)
        self.conv1 = nn.Conv2d(c_in, c_out, kernel_size=(2, 3), padding=(1, 1))
        self.selu = nn.SELU()
        self.bn2 = nn.BatchNorm2d(c_out)
        self.conv2 = nn.Conv2d(c_out, c_out, kernel_size=(2, 3), padding=(0, 1))
        self.downsample = c_in != c_out
        if self.downsample:
            self.conv_downsample = nn.Conv2d(c_in, c_out, kernel_size=(1, 3), padding=(0, 1))
        self.mp = nn.MaxPool2d((1, 3))

    def forward(self, x):
        if not self.first and self.training and self.bn1.training:
            self.bn1(x)                                   # running statistics only (the reference discards the output)
        out = self.conv2(self.selu(self.bn2(self.conv1(x))))
        skip = self.conv_downsample(x) if self.downsample else x
        return self.mp(out + skip)


class SincNetStream(nn.Module):
    def __init__(self, sinc_channels: int = 70, sinc_kernel: int = 128):
        super().__init__()
        kernel = sinc_kernel + 1 if sinc_kernel % 2 == 0 else sinc_kernel
        self.band_pass = mel_sinc_bank(sinc_channels, kernel)      # plain attribute (not a buffer): absent from state_dict
        self.first_bn = nn.BatchNorm2d(1)
        self.selu = nn.SELU()
        widths = [(1, 32), (32, 32), (32, 64), (64, 64), (64, 64), (64, 64)]
        self.encoder = nn.Sequential(*[nn.Sequential(_ResBlock(a, b, first=(i == 0))) for i, (a, b) in enumerate(widths)])
        self.out_dim = 64

    def forward(self, x, freq_aug: bool = False):
        bank = self.band_pass.to(x.device, torch.float32)
        if freq_aug:                                     # random band of up to 20 filters silenced (:121-125)
            width = int(torch.randint(0, 20, (1,)))
            start = int(torch.randint(0, bank.shape[0] - width + 1, (1,)))
            bank = bank.clone()
            bank[start:start + width] = 0
        h = F.conv1d(x.float().unsqueeze(1), bank.unsqueeze(1))                       # (B, 70, T)
        h = F.max_pool2d(h.abs().unsqueeze(1), (3, 3))                                 # (B, 1, 23, T / 3)
        h = self.encoder(self.selu(self.first_bn(h)))                                  # (B, 64, F, T')
        return h.abs().amax(dim=2).transpose(1, 2)                                     # (B, T', 64)


class RandomInitWavLMFrontend(nn.Module):
    def __init__(self, hidden: int = 1024, layers: int = 24, heads: int = 16, ffn: int = 4096, large_shaped: bool = True):
        super().__init__()
        from transformers import WavLMConfig, WavLMModel
        kw = dict(feat_extract_norm="layer", do_stable_layer_norm=True) if large_shaped else {}
        cfg = WavLMConfig(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=ffn,
                          output_hidden_states=True, mask_time_prob=0.0, mask_feature_prob=0.0, layerdrop=0.0, **kw)
        self.model = WavLMModel(cfg)
        self.out_dim = hidden
        self.layer_weights = nn.Parameter(torch.zeros(layers + 1))

    def forward(self, x):
        hs = self.model(x.float(), output_hidden_states=True).hidden_states
        w = F.softmax(self.layer_weights, dim=0)
        return (w.view(-1, 1, 1, 1) * torch.stack(hs)).sum(dim=0)                       # :420-435


class Phase6Model(nn.Module):
    """waveform (B, samples) -> (features (B, emb), logits (B, 2)); DualStreamSEMamba.py:728-769."""

    def __init__(self, wavlm_stream: nn.Module, sinc_stream: nn.Module, emb_size: int = 144, num_encoders: int = 4,
                 d_state: int = 16):
        super().__init__()
        import bimamba_b200 as bm
        self.wavlm_stream = wavlm_stream
        self.sinc_stream = sinc_stream
        self.fusion = bm.DualStreamFusion(wavlm_stream.out_dim, sinc_stream.out_dim, emb_size, reduction=16)
        tail = bm.BiMambaBackend(emb_size, num_encoders, d_state)
        self.backbone_layers, self.norm_f = tail.backbone_layers, tail.norm_f
        self.attention_pool, self.dropout, self.classifier = tail.attention_pool, tail.dropout, tail.classifier

    def forward(self, x, Freq_aug: bool = False):
        if x.ndim == 3:
            x = x.squeeze(-1)
        f_fused = self.fusion(self.wavlm_stream(x), self.sinc_stream(x, freq_aug=Freq_aug))
        for layer in self.backbone_layers:
            f_fused = layer(f_fused)
        import bimamba_b200 as bm
        return bm.backend_head(self, f_fused)
