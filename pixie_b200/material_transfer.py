"""Material field -> MPM particles, on the device.

The reference goes through files and the CPU: `map_pred_to_ply` (pixie/voxel/map_pred_to_coords.py:128-283) unscales the
packed (3+8, D, D, D) prediction, keeps the occupied voxels and writes a PLY; `apply_material_field_to_simulation`
(PG/material_field.py:295-341) reads it back, runs a scikit-learn kNN (k = 10) with a Python loop over every particle
(`perform_knn_smoothing`, :228-293) and uploads the result with one kernel launch per particle
(`_apply_material_properties_to_solver`, :343-363). Here the same three steps stay on the GPU:

    field = extract_material_points(pred, mask, min_bounds, max_bounds, ranges)      # = the PLY's vertex table
    props = perform_knn_smoothing(query_positions, field)                            # same tuple as the reference returns
    apply_material_properties_to_solver(mpm_solver, *props[1:5])

Names, argument meaning, defaults and return order follow the reference functions. No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from .mpm_solver_warp import get_material_name

#: normalization_stats/normalization_ranges.yaml p1/p99 (config keys training.{density,E,nu}_{min,max})
DEFAULT_RANGES = dict(density_min=1.703, density_max=3.871, E_min=3.018, E_max=10.882, nu_min=0.2103, nu_max=0.4493)
#: material_field.py:16-23
DEFAULT_VALUES = {"density": 1000.0, "E": 5000.0, "nu": 0.3, "part_label": 0, "material_id": "stationary"}


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def extract_material_points(pred: torch.Tensor, mask: torch.Tensor, min_bounds: Sequence[float], max_bounds: Sequence[float],
                            ranges: Optional[Dict[str, float]] = None) -> Dict[str, torch.Tensor]:
    """`unscale_prediction` + the vertex table of `map_pred_to_ply` (map_pred_to_coords.py:41-75, 198-245).

    pred: (3 + n_classes, D, D, D) float32 cuda — 3 continuous channels in ~[-1, 1] followed by class scores / one-hot;
    mask: (D, D, D), occupied where > 0. Returns the `params` dict PG/material_field.py works on: pos (M,3), density, E, nu,
    material_id, part_labels (= material_id, map_pred_to_coords.py:232), conf — device tensors, voxels in C order."""
    lib = _lib.require_device()
    r = dict(DEFAULT_RANGES if ranges is None else ranges)
    if pred.dim() != 4 or pred.shape[1] != pred.shape[2] or pred.shape[2] != pred.shape[3] or pred.shape[0] < 4:
        raise ValueError(f"pred must be (3 + n_classes, D, D, D), got {tuple(pred.shape)}")
    D, K = int(pred.shape[1]), int(pred.shape[0]) - 3
    if tuple(mask.shape) != (D, D, D):
        raise ValueError(f"Mask shape {tuple(mask.shape)} does not match grid shape {(D, D, D)}")      # map_pred_to_coords.py:190-191
    if not pred.is_cuda:
        raise _lib.PixieError("extract_material_points requires CUDA tensors; there is no CPU fallback")
    dev = pred.device
    pred = pred.detach().to(torch.float32).contiguous()
    maskf = mask.detach().to(dev, torch.float32).contiguous()
    n = D ** 3
    with torch.cuda.device(dev):
        pos = torch.empty((n, 3), dtype=torch.float32, device=dev)
        dens, E, nu, conf = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(4))
        mat = torch.empty(n, dtype=torch.int32, device=dev)
        rng = (C.c_double * 6)(r["density_min"], r["density_max"], r["E_min"], r["E_max"], r["nu_min"], r["nu_max"])
        bmin = (C.c_double * 3)(*[float(v) for v in min_bounds])
        bmax = (C.c_double * 3)(*[float(v) for v in max_bounds])
        cnt = C.c_int(0)
        _lib.check(lib.pixie_field_extract(_ptr(pred), K, _ptr(maskf), D, rng, bmin, bmax, _ptr(pos), _ptr(dens), _ptr(E), _ptr(nu), _ptr(mat),
                                           _ptr(conf), C.byref(cnt), _stream(dev)))
    m = cnt.value
    return {"pos": pos[:m], "density": dens[:m], "E": E[:m], "nu": nu[:m], "material_id": mat[:m], "part_labels": mat[:m].clone(),
            "conf": conf[:m]}


def perform_knn_smoothing(query_positions: torch.Tensor, params: Dict[str, torch.Tensor], k_smoothing_neighbors: int = 10,
                          nn_distance_threshold: float = 0.1, weighted_assignment: bool = False
                          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """material_field.py:228-293. `query_positions`: (Np, 3) MPM particle positions ALREADY in the material field's frame
    (the reference applies undoshift2center111 / undotransform2origin / inverse rotations first, :245-248 — elementwise torch).
    Returns (part_labels, densities, E_values, nu_values, material_ids, conf_values) like the reference."""
    lib = _lib.require_device()
    n_particles = int(query_positions.shape[0])
    keys = ("part_labels", "density", "E", "nu", "material_id", "conf")
    if len(params["part_labels"]) == n_particles:                                   # :236-238 no smoothing needed
        return tuple(params[k] for k in keys)
    if not query_positions.is_cuda:
        raise _lib.PixieError("perform_knn_smoothing requires CUDA tensors; there is no CPU fallback")
    dev = query_positions.device
    f = lambda t: t.detach().to(dev, torch.float32).contiguous()
    i = lambda t: t.detach().to(dev, torch.int32).contiguous()
    q, pos = f(query_positions), f(params["pos"])
    dens, E, nu, conf = f(params["density"]), f(params["E"]), f(params["nu"]), f(params["conf"])
    mat, part = i(params["material_id"]), i(params["part_labels"])
    m = int(pos.shape[0])
    # get_defaults (:38-50): mean of each continuous property, "stationary" material, part label 0
    mean = lambda t, key: float(t.mean().item()) if m > 0 else float(DEFAULT_VALUES.get(key, 0.0))
    defaults = (C.c_float * 4)(mean(dens, "density"), mean(E, "E"), mean(nu, "nu"), mean(conf, "conf"))
    with torch.cuda.device(dev):
        o_d, o_E, o_nu, o_c = (torch.empty(n_particles, dtype=torch.float32, device=dev) for _ in range(4))
        o_m, o_p = (torch.empty(n_particles, dtype=torch.int32, device=dev) for _ in range(2))
        too_far = C.c_int(0)
        _lib.check(lib.pixie_knn_assign(_ptr(q), n_particles, _ptr(pos), _ptr(dens), _ptr(E), _ptr(nu), _ptr(mat), _ptr(part), _ptr(conf), m,
                                        int(k_smoothing_neighbors), float(nn_distance_threshold), int(bool(weighted_assignment)), defaults,
                                        int(get_material_name("stationary")), int(DEFAULT_VALUES["part_label"]),
                                        _ptr(o_d), _ptr(o_E), _ptr(o_nu), _ptr(o_m), _ptr(o_p), _ptr(o_c), C.byref(too_far), _stream(dev)))
    n_too_far = too_far.value
    print(f"Particles too far from nearest neighbor: {n_too_far}, Assigned: {n_particles - n_too_far}")
    assert n_too_far <= 0.1 * n_particles, (f"[CRITICAL] More than 10% of particles are too far from nearest neighbor. "
                                            f"Distance threshold: {nn_distance_threshold}.")           # :271
    return o_p, o_d, o_E, o_nu, o_m, o_c


def apply_material_properties_to_solver(mpm_solver, densities: torch.Tensor, E_values: torch.Tensor, nu_values: torch.Tensor,
                                        material_ids: torch.Tensor, device="cuda:0", exact_box_semantics: bool = True):
    """material_field.py:343-363. The reference passes one tiny box (+-0.001) per particle to `apply_additional_params`
    (mpm_utils.py:591-610), so a particle takes the properties of the LAST particle whose box contains it — itself unless a
    later particle sits within 1e-3 of it on every axis. `exact_box_semantics=True` reproduces that (one launch over all
    boxes, O(Np^2) box tests on the device); False writes each particle's own properties (what the loop intends)."""
    n = mpm_solver.n_particles
    dev = mpm_solver._device
    f = lambda t: t.detach().to(dev, torch.float32).contiguous()
    d, E, nu = f(densities), f(E_values), f(nu_values)
    mat = material_ids.detach().to(dev, torch.int32).contiguous()
    assert d.numel() == n and E.numel() == n and nu.numel() == n and mat.numel() == n
    if exact_box_semantics:
        x = mpm_solver.mpm_state.particle_x.tensor.reshape(n, 3)
        size = torch.full((n, 3), 0.001, dtype=torch.float32, device=dev)
        boxes = torch.cat([x, size, E.view(n, 1), nu.view(n, 1), d.view(n, 1), mat.to(torch.float32).view(n, 1)], dim=1).contiguous()
        mpm_solver._apply_additional_params_boxes(boxes)
    else:
        mpm_solver.mpm_model.E = E
        mpm_solver.mpm_model.nu = nu
        mpm_solver.mpm_state.particle_density = d
        mpm_solver.mpm_state.particle_material = mat
        _lib.check(_lib.load().pixie_mpm_compute_mass(mpm_solver._handle, mpm_solver._stream()))
    mpm_solver.finalize_mu_lam(device=device)
