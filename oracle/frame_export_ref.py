"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's per-frame helpers.

  get_particle_volume        third_party/PhysGaussian/particle_filling/filling.py:247-288 (Taichi kernels, f32 / i32 fields)
  render_frame_transform     third_party/PhysGaussian/gs_simulation.py:591-600 with utils/transformation_utils.py:19-20
                             (undotransform2origin), :57-87 (cov helpers), :101-126 (undoshift2center111, inverse rotations)

PINNED BY THE REFERENCE'S OWN FUNCTIONS: tests/golden/make_transfer_golden.py executes the reference's Taichi kernels on a
minimal float32 `ti` stand-in and its transformation_utils functions on CPU torch (the device="cuda" literals dropped);
tests/test_transfer_golden.py compares against that fixture (volume bit-exact, transforms to 2e-6). Only tests/ may import this file.
"""
import numpy as np


def get_particle_volume(pos: np.ndarray, grid_n: int, grid_dx: float, unifrom: bool = False) -> np.ndarray:
    pos = pos.reshape(-1, 3).astype(np.float32)
    dx = np.float32(grid_dx)
    idx = np.floor(pos / dx).astype(np.int64)
    idx = np.clip(idx, 0, grid_n - 1)                      # the Taichi kernel has no bounds check; the product clamps
    flat = (idx[:, 0] * grid_n + idx[:, 1]) * grid_n + idx[:, 2]
    counts = np.bincount(flat, minlength=grid_n ** 3)
    vol = (dx * dx * dx) / counts[flat].astype(np.float32)
    vol = vol.astype(np.float32)
    if unifrom:
        return np.full(len(pos), np.mean(vol), dtype=np.float32)
    return vol


_FULL_FROM_UPPER = [0, 1, 2, 1, 3, 4, 2, 4, 5]      # symmetric 3x3 (row-major) from (xx, xy, xz, yy, yz, zz)   transformation_utils.py:63-77
_UPPER_FROM_FULL = [0, 1, 2, 4, 5, 8]               # and back                                                   :80-87


def _mat_from_upper(u):
    return u.reshape(-1, 6)[:, _FULL_FROM_UPPER].reshape(-1, 3, 3)


def _upper_from_mat(m):
    return m.reshape(-1, 9)[:, _UPPER_FROM_FULL]


def render_frame_transform(pos, cov, z_shift_value, scale_origin, original_mean_pos, rotation_matrices):
    pos = pos.reshape(-1, 3).astype(np.float64)
    p = pos - np.array([1.0, 1.0, 1.0]) - np.array([0.0, 0.0, z_shift_value])          # undoshift2center111
    p = np.asarray(original_mean_pos, np.float64) + p / float(scale_origin)             # undotransform2origin
    for i in range(len(rotation_matrices)):                                             # apply_inverse_rotations
        R = np.asarray(rotation_matrices[len(rotation_matrices) - 1 - i], np.float64)
        p = p @ R
    c_out = None
    if cov is not None:
        m = _mat_from_upper(cov.astype(np.float64) / float(scale_origin) ** 2)
        for i in range(len(rotation_matrices)):                                         # apply_inverse_cov_rotations
            R = np.asarray(rotation_matrices[len(rotation_matrices) - 1 - i], np.float64)
            Rt = R.T
            m = Rt @ (m @ Rt.T)                                                         # apply_cov_rotation(cov, R.T)
        c_out = _upper_from_mat(m)
    return p, c_out
