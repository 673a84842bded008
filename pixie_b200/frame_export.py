"""Either side of the MPM substep loop, on the device.

    get_particle_volume(pos, grid_n, grid_dx, unifrom=False)      PG/particle_filling/filling.py:273-288 (Taichi in the reference)
    render_frame_transform(pos, cov, z_shift_value, scale_origin, original_mean_pos, rotation_matrices)
                                                                   PG/gs_simulation.py:591-600 + utils/transformation_utils.py
Same argument meaning as the reference (including its `unifrom` spelling). No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def get_particle_volume(pos: torch.Tensor, grid_n: int, grid_dx: float, unifrom: bool = False) -> torch.Tensor:
    """vol[p] = grid_dx^3 / (number of particles in p's cell); with `unifrom` the mean volume for every particle (:282-285)."""
    lib = _lib.require_device()
    if not pos.is_cuda:
        raise _lib.PixieError("get_particle_volume requires a CUDA tensor; there is no CPU fallback")
    p = pos.detach().reshape(-1, 3).to(torch.float32).contiguous()
    n = p.shape[0]
    with torch.cuda.device(p.device):
        vol = torch.empty(n, dtype=torch.float32, device=p.device)
        _lib.check(lib.pixie_particle_volume(C.c_void_p(p.data_ptr()), n, int(grid_n), float(grid_dx), C.c_void_p(vol.data_ptr()), _stream(p.device)))
    if unifrom:
        return torch.mean(vol).repeat(n)
    return vol


def render_frame_transform(pos: torch.Tensor, cov: Optional[torch.Tensor], z_shift_value: float, scale_origin, original_mean_pos,
                           rotation_matrices: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(pos_render, cov3D_render) of gs_simulation.py:594-598: simulation frame -> the Gaussians' original frame."""
    lib = _lib.require_device()
    if not pos.is_cuda:
        raise _lib.PixieError("render_frame_transform requires CUDA tensors; there is no CPU fallback")
    dev = pos.device
    p = pos.detach().reshape(-1, 3).to(torch.float32).contiguous()
    n = p.shape[0]
    c = None if cov is None else cov.detach().reshape(-1, 6).to(dev, torch.float32).contiguous()
    if c is not None and c.shape[0] != n:
        raise ValueError("pos and cov disagree on the particle count")
    scale = float(scale_origin.item() if torch.is_tensor(scale_origin) else scale_origin)
    mean = [float(v) for v in (original_mean_pos.detach().cpu().tolist() if torch.is_tensor(original_mean_pos) else original_mean_pos)]
    rots = [r.detach().to("cpu", torch.float32).reshape(9).tolist() for r in rotation_matrices]
    if len(rots) > 8:
        raise ValueError("at most 8 rotation matrices")
    flat = (C.c_float * max(1, 9 * len(rots)))(*[v for r in rots for v in r])
    with torch.cuda.device(dev):
        po = torch.empty_like(p)
        co = None if c is None else torch.empty_like(c)
        _lib.check(lib.pixie_frame_transform(C.c_void_p(p.data_ptr()), C.c_void_p(c.data_ptr()) if c is not None else None, n, float(z_shift_value),
                                             scale, (C.c_float * 3)(*mean), flat, len(rots), C.c_void_p(po.data_ptr()),
                                             C.c_void_p(co.data_ptr()) if co is not None else None, _stream(dev)))
    return po, co
