"""Synthetic inputs of BASELINE.json's configs (there are no datasets or checkpoints offline):
seeded U-Net parameters, CLIP-like voxel grids and MPM particle scenes.
Pure data generation — no arithmetic of the hot path lives here."""
from __future__ import annotations

import math
import re
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

_NORM_RE = re.compile(r"(in_layers\.0|out_layers\.0|\.norm|unet\.out\.0|projector\.net\.(1|4|7))\.(weight|bias)$")


def seeded_state_dict(keys: Dict[str, Tuple[int, ...]], seed: int = 0, conv_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Well-conditioned parameters for every state-dict entry (`keys`: name -> shape, e.g.
    pixie_b200.unet._expected_keys). A freshly constructed reference network outputs exactly 0 (its last
    convolutions are zero_module'd, nn.py:67-73), so every tensor is overwritten: conv weights
    ~ N(0, gain / fan_in), conv biases 0.1 N(0,1), norm scales 1 + 0.1 N(0,1), norm shifts 0.1 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in keys.items():
        is_norm = bool(_NORM_RE.search(name))
        if name.endswith("weight") and not is_norm:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = torch.randn(shape, generator=g) * math.sqrt(conv_gain / fan_in)
        elif name.endswith("weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[name] = 0.1 * torch.randn(shape, generator=g)
    return sd


def synthetic_features_ndhwc(n: int, channels: int, grid: int, seed: int = 0, scale: float = 0.05) -> torch.Tensor:
    """fp16 (N, D, H, W, C): the on-disk layout of clip_features_features.npy (voxelize.py:86,111)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, grid, grid, grid, channels, generator=g) * scale).to(torch.float16)


def synthetic_scene(n: int, n_grid: int, grid_lim: float = 2.0, seed: int = 0, materials: Sequence[int] = (0,),
                    lo: float = 0.6, hi: float = 1.4):
    """Config 3 scene: uniform particles in [lo,hi]^3 of a grid_lim box, vol = dx^3 / count_in_cell
    (PhysGaussian particle_filling/filling.py:247-288), per-particle E / nu / density in the U-Net
    field's post-unscale ranges (normalization_stats/normalization_ranges.yaml)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    dx = grid_lim / n_grid
    cell = np.floor(x / dx).astype(np.int64)
    key = (cell[:, 0] * n_grid + cell[:, 1]) * n_grid + cell[:, 2]
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    vol = (dx ** 3 / cnt[inv]).astype(np.float32)
    density = rng.uniform(200.0, 2000.0, size=n).astype(np.float32)
    E = (10.0 ** rng.uniform(4.0, 6.5, size=n)).astype(np.float32)
    nu = rng.uniform(0.21, 0.45, size=n).astype(np.float32)
    material = np.asarray(materials, dtype=np.int32)[rng.integers(0, len(materials), size=n)]
    v = (0.1 * rng.standard_normal((n, 3))).astype(np.float32)
    return dict(x=x, v=v, vol=vol, density=density, E=E, nu=nu, material=material)
