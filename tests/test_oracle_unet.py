"""Pins oracle/unet_ref.py: (1) against the golden vectors produced by the reference modules
(tests/golden/make_unet_golden.py), (2) bit-for-bit against the reference modules themselves when
/root/reference is present (build container only)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import unet_ref as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "unet_small.npz")
REF = "/root/reference/third_party/Wavelet-Generation"


def _small(out, seed):
    net = (O.RegressionUNet if out == 3 else O.SegmentationUNet)(64, 32, 64, 3, (1, 1, 2, 4), (), 8, out).eval()
    net.load_state_dict(O.seeded_state_dict(net, seed))
    return net


@pytest.mark.parametrize("name,out", [("reg", 3), ("seg", 8)])
def test_oracle_matches_golden(name, out):
    g = np.load(GOLD)
    net = _small(out, int(g[f"{name}_seed"]))
    torch.set_num_threads(1)
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]))
    # same torch ops in the same order: agreement to fp32 round-off across torch builds / ISAs
    assert np.abs(y.numpy() - g[f"{name}_y"]).max() < 2e-5
    assert np.abs(g[f"{name}_y"]).max() > 0.5        # non-vacuous: zero_module'd tensors were re-seeded


def test_fresh_reference_like_network_is_zero():
    """default initialisation gives exactly 0, hence the seeded parameters."""
    net = O.RegressionUNet(64, 32, 64, 3, (1, 1, 2, 4), (), 8, 3).eval()
    with torch.no_grad():
        y = net(O.synthetic_features(1, 64, 8, seed=0))
    assert float(y.abs().max()) == 0.0


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present on this box")
def test_oracle_bit_identical_to_reference_modules():
    sys.path.insert(0, REF)
    from models.module.diffusion_network import FeatureProjector, MyUNetModel
    import torch.nn as nn

    class RefNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.projector = FeatureProjector(96, out_channels=32, hidden_channels=128)
            self.unet = MyUNetModel(in_channels=32, model_channels=64, out_channels=3, num_res_blocks=3,
                                    channel_mult=(1, 1, 2, 4), attention_resolutions=(), spatial_size=16, dims=3,
                                    activation=nn.LeakyReLU(0.02))

        def forward(self, x):
            return self.unet(self.projector(x))

    ref = RefNet().eval()
    mine = O.RegressionUNet(96, 32, 64, 3, (1, 1, 2, 4), (), 16, 3).eval()
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    sd = O.seeded_state_dict(mine, 0)
    ref.load_state_dict(sd)
    mine.load_state_dict(sd)
    x = O.synthetic_features(1, 96, 16, seed=3)
    with torch.no_grad():
        assert torch.equal(ref(x), mine(x))


def test_light_projector_and_no_projector_variants():
    """feature types rgb/occupancy use the single-layer projector; feature_channels == cond_dim uses none
    (training_discrete.py:62-68)."""
    for C in (3, 32):
        net = O.SegmentationUNet(C, 32, 64, 1, (1, 2), (), 8, 8).eval()
        net.load_state_dict(O.seeded_state_dict(net, 1))
        with torch.no_grad():
            y = net(O.synthetic_features(1, C, 8, seed=2, scale=1.0))
        assert y.shape == (1, 8, 8, 8, 8) and torch.isfinite(y).all()
        assert (net.projector is None) == (C == 32)
