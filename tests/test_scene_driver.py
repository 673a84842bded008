"""the batched multi-scene driver. CPU: host-side helpers against the reference's own functions (fixture made
from transformation_utils.py, tests/golden/transfer_golden.npz). GPU: three scenes through one warm driver equal the same
scenes processed one at a time through the individual (oracle-checked) entry points."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def test_host_helpers_follow_reference():
    from pixie_b200 import scene_driver as SD
    g = torch.Generator().manual_seed(0)
    p = torch.randn(50, 3, generator=g)
    t, scale, mean = SD.transform2origin(p)
    assert torch.allclose(t.max(0)[0] - t.min(0)[0], (p.max(0)[0] - p.min(0)[0]) * scale)
    assert float((t.max(0)[0] - t.min(0)[0]).max()) == pytest.approx(1.0, abs=1e-6)            # unit box (transformation_utils.py:6-16)
    assert torch.allclose(mean, (p.min(0)[0] + p.max(0)[0]) / 2)
    R = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    assert torch.allclose(SD.apply_rotations(p, [R]), p @ R.T)

    class Rec:
        def __init__(self):
            self.calls = []

        def __getattr__(self, name):
            return lambda **kw: self.calls.append((name, kw))

    r = Rec()
    SD.set_boundary_conditions(r, [{"type": "bounding_box"}, {"type": "cuboid", "point": [1, 1, 1], "size": [1, 1, 0.1], "velocity": [0, 0, 0]},
                                   {"type": "particle_impulse", "force": [1, 0, 0], "num_dt": 3}], {"substep_dt": 1e-4})
    assert [c[0] for c in r.calls] == ["add_bounding_box", "set_velocity_on_cuboid", "add_impulse_on_particles"]
    assert r.calls[2][1]["dt"] == 1e-4
    with pytest.raises(TypeError):
        SD.set_boundary_conditions(r, [{"type": "nope"}], {"substep_dt": 1e-4})


def _scenes(tmp, n_scenes, C, G, n_particles):
    from oracle import unet_ref as O
    from pixie_b200 import scene_driver as SD
    from pixie_b200 import voxel_io as V
    scenes = []
    for i in range(n_scenes):
        a = O.synthetic_features(1, C, G, seed=70 + i)[0].permute(1, 2, 3, 0).contiguous().to(torch.float16).numpy()
        d = os.path.join(tmp, f"obj{i}")
        os.makedirs(d)
        np.save(os.path.join(d, V.FEATURE_FILE), a)
        rng = np.random.default_rng(i)
        mask = (rng.uniform(size=(G, G, G)) < 0.6).astype(np.float32)
        np.save(os.path.join(d, V.MASK_FILE), mask)
        pts = torch.from_numpy(rng.uniform(-0.45, 0.45, size=(n_particles, 3)).astype(np.float32))
        A = rng.standard_normal((n_particles, 3, 3)).astype(np.float32) * 0.01
        cm = A @ A.transpose(0, 2, 1)
        cov = torch.from_numpy(np.stack([cm[:, 0, 0], cm[:, 0, 1], cm[:, 0, 2], cm[:, 1, 1], cm[:, 1, 2], cm[:, 2, 2]], axis=1).copy())
        ang = 0.2 * (i + 1)
        R = torch.tensor([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
        scenes.append(SD.Scene(name=f"obj{i}", grid=os.path.join(d, V.FEATURE_FILE), mask=os.path.join(d, V.MASK_FILE),
                               min_bounds=[-0.5, -0.5, -0.5], max_bounds=[0.5, 0.5, 0.5], particles=pts, cov=cov,
                               material_params={"n_grid": 32, "grid_lim": 2.0, "material": "jelly", "g": [0.0, 0.0, -9.8], "density": 1000.0,
                                                "E": 1e5, "nu": 0.3, "grid_v_damping_scale": 0.9999},
                               bc_params=[{"type": "bounding_box"}],
                               time_params={"substep_dt": 1e-4, "frame_dt": 2e-3, "frame_num": 3},
                               rotation_matrices=[R], z_shift_value=0.05, nn_distance_threshold=0.2))
    return scenes


@pytest.mark.gpu
def test_three_scenes_through_one_warm_driver(built_lib, cuda_dev, tmp_path):
    from oracle import unet_ref as O
    from pixie_b200 import frame_export as FE
    from pixie_b200 import material_transfer as MT
    from pixie_b200 import scene_driver as SD
    from pixie_b200 import voxel_io as V
    from pixie_b200.mpm_solver_warp import MPM_Simulator_WARP
    C, G, NP = 64, 16, 1500
    seg, reg = O.build_pair(C, G, seed=3)
    scenes = _scenes(str(tmp_path), 3, C, G, NP)
    # random-weight networks leave the trained [-1, 1] output range; narrow un-scaling ranges keep the decoded E / nu / density
    # of every particle physical (nu < 0.5) and the rollout CFL-stable, so that both runs stay comparable
    ranges = dict(density_min=2.95, density_max=3.05, E_min=4.4, E_max=4.6, nu_min=0.29, nu_max=0.31)
    drv = SD.SceneBatchDriver(feature_channels=C, grid_size=G, device=cuda_dev, seg_state_dict=seg.state_dict(),
                              cont_state_dict=reg.state_dict(), ranges=ranges, **O.DEFAULT_CFG)
    out = drv.run(scenes, out_dir=str(tmp_path / "preds"))
    assert [r["name"] for r in out] == ["obj0", "obj1", "obj2"]
    for sc, rec in zip(scenes, out):
        # (1) the field written like save_predictions: (3 + 8, D, D, D) float32, one-hot classes
        pred = np.load(os.path.join(str(tmp_path / "preds"), sc.name, "sample_0_pred.npy"))
        assert pred.shape == (11, G, G, G) and pred.dtype == np.float32 and np.all(pred[3:].sum(0) == 1.0)
        single = drv.predictor.predict_packed_host(V.load_feature_grid(sc.grid))[0].numpy()
        # two runs of the same scene differ by the order of the fp32 atomics (split-K, norm moments) re-rounded through the
        # fp16 / E5M2 operand split: a few 1e-4, inside the 1e-3 parity budget both hold against the oracle
        assert np.abs(pred[:3] - single[:3]).max() < 5e-4 and (pred[3:] == single[3:]).mean() > 0.999
        # (2) the rollout: same scene by hand through the individual entry points
        dev = cuda_dev
        R = sc.rotation_matrices[0].to(dev)
        rotated = sc.particles.to(dev) @ R.T
        t, scale, mean = SD.transform2origin(rotated)
        pos0 = t + torch.tensor([1.0, 1.0, 1.05], device=dev)
        cloud = MT.extract_material_points(torch.from_numpy(pred).to(dev), V.load_mask(sc.mask).to(dev), sc.min_bounds, sc.max_bounds, ranges)
        s = MPM_Simulator_WARP(10, device=dev)
        vol = FE.get_particle_volume(pos0, 32, 2.0 / 32)
        cm = sc.cov.to(dev)
        m = torch.stack([cm[:, 0], cm[:, 1], cm[:, 2], cm[:, 1], cm[:, 3], cm[:, 4], cm[:, 2], cm[:, 4], cm[:, 5]], dim=1).view(-1, 3, 3)
        m = R @ m @ R.T
        cov0 = torch.stack([m[:, 0, 0], m[:, 0, 1], m[:, 0, 2], m[:, 1, 1], m[:, 1, 2], m[:, 2, 2]], dim=1) * scale ** 2
        s.load_initial_data_from_torch(pos0, vol, cov0, n_grid=32, grid_lim=2.0, device=dev)
        s.set_parameters_dict(dict(sc.material_params), device=dev)
        s.add_bounding_box()
        q, _ = FE.render_frame_transform(s.export_particle_x_to_torch(), None, 0.0, scale, mean, [R])
        props = MT.perform_knn_smoothing(q, cloud, 10, 0.2)
        MT.apply_material_properties_to_solver(s, props[1], props[2], props[3], props[4], device=dev, exact_box_semantics=False)
        assert torch.equal(props[4].cpu(), rec["material_ids"].cpu())
        assert len(rec["frames_pos"]) == 3 and rec["substeps"] == 60 and abs(rec["time"] - 60e-4) < 1e-12
        for f in range(3):
            pr, cr = FE.render_frame_transform(s.export_particle_x_to_torch(), s.export_particle_cov_to_torch().view(-1, 6), 0.05, scale, mean, [R])
            assert (pr - rec["frames_pos"][f]).abs().max() < 2e-5, (sc.name, f)
            assert (cr - rec["frames_cov"][f]).abs().max() < 1e-6
            assert torch.isfinite(pr).all()
            s.p2g2p_n(20, 1e-4)
        # frame 0 is the untouched input cloud back in its own frame
        assert (rec["frames_pos"][0].cpu() - sc.particles).abs().max() < 1e-5
    assert (out[0]["frames_pos"][2] - out[1]["frames_pos"][2]).abs().max() > 1e-3          # scenes are really different
