"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's material-field -> particle
transfer, following the reference line by line with numpy / scikit-learn as the reference does.

  unscale_prediction            pixie/voxel/map_pred_to_coords.py:41-75
  vertex_table                  pixie/voxel/map_pred_to_coords.py:198-245   (what map_pred_to_ply writes into the PLY)
  MaterialProperties, perform_knn_smoothing
                                third_party/PhysGaussian/material_field.py:26-86, 228-293
  apply_additional_params       third_party/PhysGaussian/mpm_solver_warp/mpm_utils.py:591-610 via material_field.py:343-363

PARITY UNPINNED: the reference modules import hydra / plyfile / warp / taichi (absent here), so they cannot be imported to
generate golden vectors, and the reference has no tests for this path. The functions below are pinned by hand-computed
cases in tests/test_material_transfer.py only. Only tests/ may import this file.
"""
from __future__ import annotations

from collections import Counter

import numpy as np

DEFAULT_VALUES = {"density": 1000.0, "E": 1e6, "nu": 0.3, "part_label": 0, "material_id": "stationary"}   # material_field.py:16-23
STATIONARY_ID = 6                                                                                       # mpm_solver_warp.py:10-26


def unscale_prediction(pred_tensor: np.ndarray, r: dict) -> np.ndarray:
    cont = pred_tensor[:3]
    cont = np.clip(cont, -1.0, 1.0)
    out = pred_tensor.copy().astype(np.float32)
    dens_log = (cont[0] + 1.0) * (r["density_max"] - r["density_min"]) / 2.0 + r["density_min"]
    out[0] = 10 ** dens_log
    E_log = (cont[1] + 1.0) * (r["E_max"] - r["E_min"]) / 2.0 + r["E_min"]
    out[1] = 10 ** E_log
    out[2] = (cont[2] + 1.0) * (r["nu_max"] - r["nu_min"]) / 2.0 + r["nu_min"]
    return out


def vertex_table(scaled_pred: np.ndarray, mask: np.ndarray, min_bounds, max_bounds, r: dict) -> dict:
    pred = unscale_prediction(scaled_pred, r)
    grid_shape = mask.shape
    cont, seg = pred[:3, :], pred[3:, :]
    material_id = seg[0] if seg.shape[0] == 1 else np.argmax(seg, axis=0)          # get_mat_id :122-126
    x = np.linspace(min_bounds[0], max_bounds[0], grid_shape[0])
    y = np.linspace(min_bounds[1], max_bounds[1], grid_shape[1])
    z = np.linspace(min_bounds[2], max_bounds[2], grid_shape[2])
    gx, gy, gz = np.meshgrid(x, y, z, indexing="ij")
    coords = np.stack([gx, gy, gz], axis=-1)
    valid = mask > 0
    conf = np.max(seg, axis=0)[valid] if seg.shape[0] > 1 else np.ones(int(valid.sum()), dtype=np.float32)
    return {"pos": coords[valid].astype(np.float32),                               # PLY fields are 'f4' / 'i4' (:222-231)
            "density": cont[0][valid].astype(np.float32), "E": cont[1][valid].astype(np.float32), "nu": cont[2][valid].astype(np.float32),
            "material_id": material_id[valid].astype(np.int32), "part_labels": material_id[valid].astype(np.int32),
            "conf": conf.astype(np.float32)}


class MaterialProperties:
    def __init__(self, part_labels, densities, E_values, nu_values, material_ids, conf_values):
        self.properties = {"part_labels": part_labels, "density": densities, "E": E_values, "nu": nu_values,
                           "material_id": material_ids, "conf": conf_values}

    def get_defaults(self, n_particles):
        defaults = {}
        for key, values in self.properties.items():
            if key == "material_id":
                default_val = STATIONARY_ID
            elif key in ["part_labels"]:
                default_val = DEFAULT_VALUES["part_label"]
            else:
                default_val = np.mean(values) if len(values) > 0 else DEFAULT_VALUES.get(key, 0.0)
            defaults[key] = np.full(n_particles, default_val, dtype=values.dtype if hasattr(values, "dtype") else np.float32)
        return defaults

    def assign_from_neighbors(self, particle_idx, neighbor_indices, distances, weighted=False):
        results = {}
        weights = 1.0 / (distances + 1e-8)
        weights = weights / np.sum(weights)
        for prop_name, prop_values in self.properties.items():
            neighbor_values = prop_values[neighbor_indices]
            if prop_name in ["material_id", "part_labels"]:
                if weighted:
                    unique_vals, inv_indices = np.unique(neighbor_values, return_inverse=True)
                    votes = np.bincount(inv_indices, weights=weights)
                    results[prop_name] = unique_vals[np.argmax(votes)]
                else:
                    results[prop_name] = Counter(neighbor_values).most_common(1)[0][0]
            else:
                results[prop_name] = np.dot(weights, neighbor_values) if weighted else np.mean(neighbor_values)
        return results


def perform_knn_smoothing(query_positions: np.ndarray, params: dict, k_smoothing_neighbors=10, nn_distance_threshold=0.1,
                          weighted_assignment=False):
    from sklearn.neighbors import NearestNeighbors
    n_particles = len(query_positions)
    props = MaterialProperties(params["part_labels"], params["density"], params["E"], params["nu"], params["material_id"], params["conf"])
    if len(props.properties["part_labels"]) == n_particles:
        return tuple(props.properties.values())
    nn_model = NearestNeighbors(n_neighbors=k_smoothing_neighbors, algorithm="auto").fit(params["pos"])
    distances_all_k, k_indices = nn_model.kneighbors(query_positions)
    too_far_mask = distances_all_k[:, 0] > nn_distance_threshold
    n_too_far = int(np.sum(too_far_mask))
    assert n_too_far <= 0.1 * n_particles
    mapped = props.get_defaults(n_particles)
    for i in np.where(~too_far_mask)[0]:
        a = props.assign_from_neighbors(i, k_indices[i], distances_all_k[i], weighted_assignment)
        for prop_name, value in a.items():
            mapped[prop_name][i] = value
    return tuple(mapped.values())


def apply_additional_params(x: np.ndarray, densities, E_values, nu_values, material_ids, size=0.001):
    """One box per particle, applied in order (material_field.py:347-358 + mpm_utils.py:591-610): returns the per-particle
    (E, nu, density, material) after all launches. float32 comparisons like the Warp kernel."""
    x = x.astype(np.float32)
    n = len(x)
    E = np.zeros(n, np.float32); nu = np.zeros(n, np.float32); d = np.zeros(n, np.float32); m = np.zeros(n, np.int32)
    s = np.float32(size)
    for i in range(n):
        inside = np.all((x > x[i] - s) & (x < x[i] + s), axis=1)
        E[inside] = np.float32(E_values[i]); nu[inside] = np.float32(nu_values[i]); d[inside] = np.float32(densities[i]); m[inside] = int(material_ids[i])
    return E, nu, d, m
