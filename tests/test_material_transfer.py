"""material field -> particles. CPU: the oracle against hand-computed cases (the reference modules cannot be
imported here, see oracle/material_transfer_ref.py). GPU: the device path against the oracle on seeded inputs."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import material_transfer_ref as R  # noqa: E402

RANGES = dict(density_min=1.703, density_max=3.871, E_min=3.018, E_max=10.882, nu_min=0.2103, nu_max=0.4493)


def _scene(D=12, K=8, seed=0, fill=0.4):
    rng = np.random.default_rng(seed)
    pred = np.zeros((3 + K, D, D, D), np.float32)
    pred[:3] = rng.uniform(-1.3, 1.3, size=(3, D, D, D))                     # exercises the clip
    ids = rng.integers(0, K, size=(D, D, D))
    ids[: D // 2] = np.minimum(ids[: D // 2], 2)                              # spatially coherent-ish classes
    pred[3:] = np.eye(K, dtype=np.float32)[ids].transpose(3, 0, 1, 2)         # one-hot like save_predictions
    mask = (rng.uniform(size=(D, D, D)) < fill).astype(np.float32)
    return pred, mask, np.array([-0.5, -0.4, -0.3]), np.array([0.5, 0.6, 0.7])


# ------------------------------------------------------------------------------------------------ oracle, CPU
def test_oracle_unscale_hand_values():
    pred = np.zeros((11, 2, 2, 2), np.float32)
    pred[0] = -1.0; pred[1] = 1.0; pred[2] = 0.0
    pred[0, 0, 0, 0] = 5.0                                                    # clipped to +1
    out = R.unscale_prediction(pred, RANGES)
    assert np.allclose(out[0, 1, 1, 1], 10 ** 1.703, rtol=1e-6)
    assert np.allclose(out[0, 0, 0, 0], 10 ** 3.871, rtol=1e-6)
    assert np.allclose(out[1], 10 ** 10.882, rtol=1e-6)
    assert np.allclose(out[2], (0.2103 + 0.4493) / 2, rtol=1e-6)
    assert out.dtype == np.float32 and np.array_equal(out[3:], pred[3:])


def test_oracle_vertex_table_order_and_coords():
    pred, mask, lo, hi = _scene(D=4)
    t = R.vertex_table(pred, mask, lo, hi, RANGES)
    valid = np.argwhere(mask > 0)                                             # C order
    assert len(t["pos"]) == len(valid)
    lin = [np.linspace(lo[d], hi[d], 4) for d in range(3)]
    want = np.stack([lin[0][valid[:, 0]], lin[1][valid[:, 1]], lin[2][valid[:, 2]]], axis=1).astype(np.float32)
    assert np.array_equal(t["pos"], want)
    assert np.array_equal(t["material_id"], np.argmax(pred[3:], axis=0)[mask > 0])
    assert np.all(t["conf"] == 1.0)


def test_oracle_knn_mean_mode_and_defaults():
    pos = np.array([[0, 0, 0], [0.01, 0, 0], [0.02, 0, 0], [0.03, 0, 0], [1, 1, 1]], np.float32)
    params = dict(pos=pos, density=np.array([1, 2, 3, 4, 100], np.float32), E=np.array([10, 20, 30, 40, 1000], np.float32),
                  nu=np.full(5, 0.3, np.float32), material_id=np.array([0, 2, 2, 0, 5], np.int32), part_labels=np.array([0, 2, 2, 0, 5], np.int32),
                  conf=np.ones(5, np.float32))
    q = np.array([[0.004, 0, 0], [5, 5, 5]] + [[0.005 * i, 0.001, 0] for i in range(20)], np.float32)      # 1 of 22 too far (< 10 %)
    part, dens, E, nu, mat, conf = R.perform_knn_smoothing(q, params, k_smoothing_neighbors=3, nn_distance_threshold=0.1)
    assert dens[0] == np.float32((1 + 2 + 3) / 3) and E[0] == np.float32(20.0)
    assert mat[0] == 2                                                        # neighbours (0, 2, 2) -> mode 2
    assert mat[1] == R.STATIONARY_ID and part[1] == 0 and dens[1] == np.float32(np.mean(params["density"]))
    # tie in the mode -> the value met first in neighbour (distance) order: Counter.most_common
    q2 = np.array([[0.0149, 0, 0]] + [[0.01, 0.0005 * i, 0] for i in range(1, 12)], np.float32)
    _, _, _, _, mat2, _ = R.perform_knn_smoothing(q2, params, k_smoothing_neighbors=2, nn_distance_threshold=0.1)
    assert mat2[0] == 2                                                       # neighbours: idx 1 (id 2) then idx 2 (id 2)


def test_oracle_additional_params_last_box_wins():
    x = np.array([[0.5, 0.5, 0.5], [0.5005, 0.5, 0.5], [0.7, 0.5, 0.5]], np.float32)
    E, nu, d, m = R.apply_additional_params(x, [1, 2, 3], [10, 20, 30], [0.1, 0.2, 0.3], [0, 1, 2])
    assert list(m) == [1, 1, 2] and list(E) == [20, 20, 30]                   # particle 0 sits inside particle 1's box, which is applied later


# ------------------------------------------------------------------------------------------------ device
@pytest.mark.gpu
def test_extract_matches_oracle():
    from pixie_b200.material_transfer import extract_material_points
    pred, mask, lo, hi = _scene(D=16, seed=3)
    want = R.vertex_table(pred, mask, lo, hi, RANGES)
    got = extract_material_points(torch.from_numpy(pred).cuda(), torch.from_numpy(mask).cuda(), lo, hi, RANGES)
    assert got["pos"].shape[0] == len(want["pos"])
    assert np.array_equal(got["pos"].cpu().numpy(), want["pos"])
    assert np.array_equal(got["material_id"].cpu().numpy(), want["material_id"])
    assert np.array_equal(got["conf"].cpu().numpy(), want["conf"])
    for k in ("density", "E", "nu"):
        g, w = got[k].cpu().numpy(), want[k]
        assert np.max(np.abs(g - w) / np.abs(w)) < 2e-6, k                    # powf vs numpy float32 power: <= 2 ulp


@pytest.mark.gpu
@pytest.mark.parametrize("weighted", [False, True])
def test_knn_smoothing_matches_oracle(weighted):
    from pixie_b200.material_transfer import extract_material_points, perform_knn_smoothing
    pred, mask, lo, hi = _scene(D=16, seed=5, fill=0.5)
    field_np = R.vertex_table(pred, mask, lo, hi, RANGES)
    rng = np.random.default_rng(1)
    q = rng.uniform(lo + 0.05, hi - 0.05, size=(3000, 3)).astype(np.float32)
    q[:50] += 3.0                                                             # too far -> defaults
    want = R.perform_knn_smoothing(q, field_np, 10, 0.1, weighted)
    field = extract_material_points(torch.from_numpy(pred).cuda(), torch.from_numpy(mask).cuda(), lo, hi, RANGES)
    got = perform_knn_smoothing(torch.from_numpy(q).cuda(), field, 10, 0.1, weighted)
    names = ("part_labels", "density", "E", "nu", "material_id", "conf")
    for name, g, w in zip(names, got, want):
        g = g.cpu().numpy()
        if name in ("part_labels", "material_id"):
            assert np.mean(g == w) > 0.999, name                              # exact ties in vote / distance order may differ
        else:
            rel = np.abs(g - w) / np.maximum(np.abs(w), 1e-30)
            assert np.quantile(rel, 0.999) < 5e-6, (name, rel.max())


@pytest.mark.gpu
def test_apply_to_solver_matches_box_semantics():
    from pixie_b200.material_transfer import apply_material_properties_to_solver
    from pixie_b200.mpm_solver_warp import MPM_Simulator_WARP
    rng = np.random.default_rng(2)
    n = 400
    x = rng.uniform(0.3, 0.7, size=(n, 3)).astype(np.float32)
    x[1] = x[0] + np.float32(0.0004)                                          # inside each other's boxes
    s = MPM_Simulator_WARP(10, device="cuda:0")
    s.load_initial_data_from_torch(torch.from_numpy(x).cuda(), torch.full((n,), 1e-6).cuda(), None, n_grid=32, grid_lim=1.0)
    d = rng.uniform(500, 3000, n).astype(np.float32); E = rng.uniform(1e4, 1e6, n).astype(np.float32)
    nu = rng.uniform(0.2, 0.45, n).astype(np.float32); m = rng.integers(0, 7, n).astype(np.int32)
    apply_material_properties_to_solver(s, torch.from_numpy(d), torch.from_numpy(E), torch.from_numpy(nu), torch.from_numpy(m))
    wE, wnu, wd, wm = R.apply_additional_params(x, d, E, nu, m)
    assert np.array_equal(s.mpm_model.E.numpy(), wE) and np.array_equal(s.mpm_model.nu.numpy(), wnu)
    assert np.array_equal(s.mpm_state.particle_density.numpy(), wd) and np.array_equal(s.mpm_state.particle_material.numpy(), wm)
    assert wm[0] == m[1]                                                      # the quirk is exercised
    mass = s.mpm_state.particle_mass.numpy()
    assert np.allclose(mass, wd * np.float32(1e-6), rtol=1e-6)
    mu = s.mpm_model.mu.numpy()
    assert np.allclose(mu, wE / (2 * (1 + wnu)), rtol=1e-5)
