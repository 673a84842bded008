"""§8 f-1 / f-2 parity against vectors produced by the REFERENCE'S OWN FUNCTIONS (tests/golden/make_transfer_golden.py pulls
`unscale_prediction`, `map_pred_to_ply`, `MaterialProperties`, `perform_knn_smoothing`, `_apply_material_properties_to_solver`,
`get_particle_volume` and the transformation_utils helpers out of the reference files with `ast` and executes them).

CPU: the numpy oracles (oracle/material_transfer_ref.py, oracle/frame_export_ref.py) against the fixture — bit-exact where they
use the same numpy expression.  GPU (`-m gpu`): the device kernels through the C ABI against the same fixture: integer fields
exact, float fields to 2e-6 relative (powf / summation order)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import frame_export_ref as FR  # noqa: E402
from oracle import material_transfer_ref as R  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "transfer_golden.npz"))
RANGES = dict(density_min=1.703, density_max=3.871, E_min=3.018, E_max=10.882, nu_min=0.2103, nu_max=0.4493)
KEYS = ("part_labels", "density", "E", "nu", "material_id", "conf")


def _dense():
    idx, vals = G["field/idx"], G["field/vals"]
    pred = np.zeros((vals.shape[1], 64, 64, 64), np.float32)
    mask = np.zeros((64, 64, 64), np.float32)
    pred[:, idx[:, 0], idx[:, 1], idx[:, 2]] = vals.T
    mask[idx[:, 0], idx[:, 1], idx[:, 2]] = 1.0
    return pred, mask


def _table():
    t = {k: G[f"field/table/{k}"] for k in ("x", "y", "z", "part_label", "density", "E", "nu", "material_id", "conf")}
    return {"pos": np.stack([t["x"], t["y"], t["z"]], axis=1).astype(np.float32), "part_labels": t["part_label"], "density": t["density"],
            "E": t["E"], "nu": t["nu"], "material_id": t["material_id"], "conf": t["conf"]}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ------------------------------------------------------------------------------------------------------- CPU: oracles
def test_oracle_unscale_and_vertex_table_bit_exact():
    pred, mask = _dense()
    idx = G["field/idx"]
    un = R.unscale_prediction(pred, RANGES)
    assert np.array_equal(un[:, idx[:, 0], idx[:, 1], idx[:, 2]].T, G["field/unscaled_at_idx"])
    t, want = R.vertex_table(pred, mask, G["field/min_bounds"], G["field/max_bounds"], RANGES), _table()
    for k in want:
        assert np.array_equal(t[k], want[k]), k
    # the fixture holds exact ties between two class scores: argmax must have taken the first maximum
    vals = G["field/vals"][:, 3:]
    tied = np.flatnonzero((vals == vals.max(1, keepdims=True)).sum(1) > 1)
    assert len(tied) >= 10


def test_oracle_default_values_follow_reference():
    assert G["knn/DEFAULT_E"] == 5000.0 and R.DEFAULT_VALUES["E"] == 5000.0          # material_field.py:19
    empty = {k: np.zeros(0, np.float32) for k in KEYS}
    d = R._fallback_values(empty, 3)
    for k in KEYS:
        assert np.array_equal(np.asarray(d[k], np.float64), np.asarray(G[f"knn/empty_defaults/{k}"], np.float64)), k


@pytest.mark.parametrize("weighted", [False, True], ids=["plain", "weighted"])
def test_oracle_knn_smoothing_bit_exact(weighted):
    params = _table()
    out = R.perform_knn_smoothing(G["knn/q_field"], params, 10, 0.1, weighted)
    tag = "weighted" if weighted else "plain"
    for k, v in zip(KEYS, out):
        assert np.array_equal(np.asarray(v), G[f"knn/{tag}/{k}"]), k


def test_oracle_upload_box_semantics():
    E, nu, d, m = R.apply_additional_params(G["upload/x"], G["upload/in_density"], G["upload/in_E"], G["upload/in_nu"], G["upload/in_ids"])
    assert np.array_equal(E, G["upload/E"]) and np.array_equal(nu, G["upload/nu"]) and np.array_equal(d, G["upload/density"])
    assert np.array_equal(m, G["upload/material"])
    assert (G["upload/E"] != G["upload/in_E"]).sum() == 2                               # "last box containing the particle wins"


def test_oracle_particle_volume_and_frame_transform():
    n, dx = int(G["volume/grid_n"]), float(G["volume/grid_dx"])
    assert np.array_equal(FR.get_particle_volume(G["volume/pos"], n, dx), G["volume/vol"])
    assert _rel(FR.get_particle_volume(G["volume/pos"], n, dx, unifrom=True), G["volume/vol_uniform"]) < 1e-6
    p, c = FR.render_frame_transform(G["frame/pos"], G["frame/cov"], float(G["frame/z_shift"]), float(G["knn/scale"]), G["knn/mean"], G["knn/rots"])
    assert _rel(p, G["frame/pos_render"]) < 2e-6 and _rel(c, G["frame/cov_render"]) < 2e-6
    # and the query transform of perform_knn_smoothing (material_field.py:245-248) is the same chain with z_shift = 0
    q, _ = FR.render_frame_transform(G["knn/q_sim"], None, 0.0, float(G["knn/scale"]), G["knn/mean"], G["knn/rots"])
    assert _rel(q, G["knn/q_field"]) < 2e-6


# ------------------------------------------------------------------------------------------------------- GPU: product
@pytest.mark.gpu
def test_cuda_field_extract_matches_reference(built_lib, cuda_dev):
    import torch
    from pixie_b200 import material_transfer as MT
    pred, mask = _dense()
    t = MT.extract_material_points(torch.from_numpy(pred).to(cuda_dev), torch.from_numpy(mask).to(cuda_dev), G["field/min_bounds"],
                                   G["field/max_bounds"], RANGES)
    want = _table()
    assert np.array_equal(t["material_id"].cpu().numpy(), want["material_id"])           # incl. the exact-tie voxels
    assert np.array_equal(t["part_labels"].cpu().numpy(), want["part_labels"])
    assert np.array_equal(t["pos"].cpu().numpy(), want["pos"])
    assert np.array_equal(t["conf"].cpu().numpy(), want["conf"])
    for k in ("density", "E", "nu"):
        assert _rel(t[k].cpu().numpy(), want[k]) < 2e-6 and np.abs(t[k].cpu().numpy() / want[k] - 1).max() < 5e-6, k


@pytest.mark.gpu
@pytest.mark.parametrize("weighted", [False, True], ids=["plain", "weighted"])
def test_cuda_knn_smoothing_matches_reference(built_lib, cuda_dev, weighted):
    import torch
    from pixie_b200 import frame_export as FE
    from pixie_b200 import material_transfer as MT
    params = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_dev) for k, v in _table().items()}
    # the reference's own query transform (material_field.py:245-248), on the device
    q, _ = FE.render_frame_transform(torch.from_numpy(G["knn/q_sim"]).to(cuda_dev), None, 0.0, float(G["knn/scale"]),
                                     torch.from_numpy(G["knn/mean"]), [torch.from_numpy(r) for r in G["knn/rots"]])
    assert _rel(q.cpu().numpy(), G["knn/q_field"]) < 2e-6
    out = MT.perform_knn_smoothing(torch.from_numpy(G["knn/q_field"]).to(cuda_dev), params, 10, 0.1, weighted)
    tag = "weighted" if weighted else "plain"
    for k, v in zip(KEYS, out):
        ref, got = G[f"knn/{tag}/{k}"], v.cpu().numpy()
        if k in ("part_labels", "material_id"):
            assert np.array_equal(got, ref), k
        else:
            assert np.abs(got / ref - 1).max() < 2e-6, k


@pytest.mark.gpu
def test_cuda_upload_volume_and_frame_match_reference(built_lib, cuda_dev):
    import torch
    from pixie_b200 import frame_export as FE
    from pixie_b200 import material_transfer as MT
    from pixie_b200.mpm_solver_warp import MPM_Simulator_WARP
    x, vol = G["upload/x"], G["upload/vol"]
    s = MPM_Simulator_WARP(len(x), n_grid=16, grid_lim=2.0, device=cuda_dev)
    s.load_initial_data_from_torch(torch.from_numpy(x).to(cuda_dev), torch.from_numpy(vol).to(cuda_dev), None, n_grid=16, grid_lim=2.0, device=cuda_dev)
    s.set_parameters_dict({"material": "jelly", "E": 1e5, "nu": 0.3, "density": 1000.0}, device=cuda_dev)
    t = lambda k: torch.from_numpy(G[k]).to(cuda_dev)
    MT.apply_material_properties_to_solver(s, t("upload/in_density"), t("upload/in_E"), t("upload/in_nu"), t("upload/in_ids"), device=cuda_dev)
    assert np.array_equal(s.mpm_model.E.numpy(), G["upload/E"]) and np.array_equal(s.mpm_model.nu.numpy(), G["upload/nu"])
    assert np.array_equal(s.mpm_state.particle_density.numpy(), G["upload/density"])
    assert np.array_equal(s.mpm_state.particle_material.numpy(), G["upload/material"])
    for got, k in ((s.mpm_state.particle_mass.numpy(), "mass"), (s.mpm_model.mu.numpy(), "mu"), (s.mpm_model.lam.numpy(), "lam")):
        assert np.abs(got / G[f"upload/{k}"] - 1).max() < 1e-6, k
    n, dx = int(G["volume/grid_n"]), float(G["volume/grid_dx"])
    pos = torch.from_numpy(G["volume/pos"]).to(cuda_dev)
    assert np.array_equal(FE.get_particle_volume(pos, n, dx).cpu().numpy(), G["volume/vol"])
    assert _rel(FE.get_particle_volume(pos, n, dx, unifrom=True).cpu().numpy(), G["volume/vol_uniform"]) < 1e-6
    p, c = FE.render_frame_transform(torch.from_numpy(G["frame/pos"]).to(cuda_dev), torch.from_numpy(G["frame/cov"]).to(cuda_dev),
                                     float(G["frame/z_shift"]), float(G["knn/scale"]), torch.from_numpy(G["knn/mean"]),
                                     [torch.from_numpy(r) for r in G["knn/rots"]])
    assert _rel(p.cpu().numpy(), G["frame/pos_render"]) < 2e-6 and _rel(c.cpu().numpy(), G["frame/cov_render"]) < 2e-6
