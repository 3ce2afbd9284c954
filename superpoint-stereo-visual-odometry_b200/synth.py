"""Synthetic SuperPoint head outputs (there is no dataset / checkpoint access: SURVEY.md section 8d).

A stereo stream is a fixed random "world" of detector logits and unit-norm descriptor cells that
scrolls by `temporal_shift` cells per frame; the right eye sees the left eye's field displaced by
`disparity` cells.  Small independent noise is added per image and descriptors are re-normalised,
as the network's ReduceL2 -> Div tail does (models/*.onnx).  This keeps matching non-degenerate
(purely random unit vectors give 0 ratio-test survivors).
  semi [F, 2, 65, H/8, W/8]   raw logits ~ N(0, sigma^2)            (eye 0 = left, 1 = right)
  desc [F, 2, 256, H/8, W/8]  unit L2 norm over the channel dim
"""
from __future__ import annotations

import torch


def make_stream(num_pairs: int, H: int, W: int, seed: int = 0, sigma: float = 1.0, disparity: int = 2,
                temporal_shift: int = 1, semi_noise: float = 0.05, desc_noise: float = 0.05,
                device: str | torch.device = "cpu", first_frame: int = 0):
    assert H % 8 == 0 and W % 8 == 0
    Hc, Wc = H // 8, W // 8
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(0xC0FFEE + seed)
    world_semi = torch.randn(65, Hc, Wc, generator=g, device=dev) * sigma
    world_desc = torch.randn(256, Hc, Wc, generator=g, device=dev)
    semi = torch.empty(num_pairs, 2, 65, Hc, Wc, device=dev)
    desc = torch.empty(num_pairs, 2, 256, Hc, Wc, device=dev)
    for i in range(num_pairs):
        f = first_frame + i
        gf = torch.Generator(device=dev)
        gf.manual_seed((0xC0FFEE + seed) * 1000003 + f)  # per-frame stream: shards generate identical frames
        for eye in range(2):
            sh = (f * temporal_shift + eye * disparity) % Wc
            s = torch.roll(world_semi, shifts=sh, dims=2)
            d = torch.roll(world_desc, shifts=sh, dims=2)
            s = s + semi_noise * sigma * torch.randn(s.shape, generator=gf, device=dev)
            d = d + desc_noise * torch.randn(d.shape, generator=gf, device=dev)
            d = d / d.norm(dim=0, keepdim=True)
            semi[i, eye] = s
            desc[i, eye] = d
    return semi, desc


def random_descriptors(n: int, seed: int = 0, device="cpu"):
    g = torch.Generator(device=torch.device(device))
    g.manual_seed(seed)
    d = torch.randn(n, 256, generator=g, device=device)
    return d / d.norm(dim=1, keepdim=True)
