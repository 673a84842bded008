"""particle volume and the per-frame export transform. CPU: oracle vs hand-computed cases; GPU: device vs oracle."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frame_export_ref as R  # noqa: E402


def _rot(deg, axis):
    c, s = np.cos(np.deg2rad(deg)), np.sin(np.deg2rad(deg))
    m = {0: [[1, 0, 0], [0, c, -s], [0, s, c]], 1: [[c, 0, s], [0, 1, 0], [-s, 0, c]], 2: [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[axis]
    return np.array(m, np.float32)


def test_oracle_particle_volume_hand_case():
    pos = np.array([[0.05, 0.05, 0.05], [0.06, 0.07, 0.01], [0.15, 0.05, 0.05], [0.95, 0.95, 0.95]], np.float32)
    vol = R.get_particle_volume(pos, 10, 0.1)
    assert np.allclose(vol, [0.0005, 0.0005, 0.001, 0.001], rtol=1e-6)
    assert np.allclose(R.get_particle_volume(pos, 10, 0.1, unifrom=True), np.mean(vol))


def test_oracle_frame_transform_inverts_the_forward_transform():
    rng = np.random.default_rng(0)
    x = rng.uniform(-0.3, 0.4, size=(50, 3))
    Rs = [_rot(30, 0), _rot(-20, 2)]
    mean, scale, zs = np.array([0.1, -0.2, 0.05]), 0.8, 0.1
    fwd = x.copy()
    for Rm in Rs:                                   # apply_rotations: position @ R.T
        fwd = fwd @ Rm.astype(np.float64).T
    fwd = (fwd - mean) * scale + 1.0 + np.array([0, 0, zs])          # transform2origin + shift2center111
    back, _ = R.render_frame_transform(fwd, None, zs, scale, mean, Rs)
    assert np.abs(back - x).max() < 1e-6
    cov = rng.uniform(0.1, 1, size=(50, 6))
    _, c = R.render_frame_transform(fwd, cov, zs, scale, mean, [])
    assert np.allclose(c, cov / scale ** 2)


@pytest.mark.gpu
def test_particle_volume_matches_oracle():
    from pixie_b200.frame_export import get_particle_volume
    rng = np.random.default_rng(1)
    pos = rng.uniform(0.0, 2.0, size=(20000, 3)).astype(np.float32)
    for uni in (False, True):
        got = get_particle_volume(torch.from_numpy(pos).cuda(), 64, 2.0 / 64, uni).cpu().numpy()
        want = R.get_particle_volume(pos, 64, 2.0 / 64, uni)
        assert np.allclose(got, want, rtol=2e-6)


@pytest.mark.gpu
def test_frame_transform_matches_oracle():
    from pixie_b200.frame_export import render_frame_transform
    rng = np.random.default_rng(2)
    pos = rng.uniform(0.6, 1.4, size=(5000, 3)).astype(np.float32)
    cov = rng.uniform(1e-4, 1e-2, size=(5000, 6)).astype(np.float32)
    Rs = [_rot(15, 1), _rot(40, 0), _rot(-75, 2)]
    mean, scale, zs = np.array([0.3, 0.1, -0.4], np.float32), 0.37, 0.2
    p, c = render_frame_transform(torch.from_numpy(pos).cuda(), torch.from_numpy(cov).cuda(), zs, torch.tensor(scale).cuda(),
                                  torch.from_numpy(mean).cuda(), [torch.from_numpy(r).cuda() for r in Rs])
    wp, wc = R.render_frame_transform(pos, cov, zs, np.float32(scale), mean, Rs)
    assert np.abs(p.cpu().numpy() - wp).max() < 2e-6 * 10
    assert np.abs(c.cpu().numpy() - wc).max() < 1e-6 * np.abs(wc).max() * 10
    p2, c2 = render_frame_transform(torch.from_numpy(pos).cuda(), None, zs, scale, mean.tolist(), [])
    assert c2 is None and np.abs(p2.cpu().numpy() - R.render_frame_transform(pos, None, zs, scale, mean, [])[0]).max() < 1e-5
