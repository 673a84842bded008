"""Test double for pixie_b200.mpm_slab: the backend protocol of CudaSlabBackend implemented on the CPU oracle, so the
slab orchestration (overlap exchange, migration, id bookkeeping) is exercised without a GPU — in one process
(LocalSlabCluster) and over gloo (DistSlabDriver)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mpm_ref as O   # noqa: E402

FIELDS = [(name, O.WIDTH[idx]) for name, idx in sorted(O.FIELDS.items(), key=lambda kv: kv[1])]


class OracleSlabBackend:
    def __init__(self, n_grid: int, grid_lim: float, capacity: int, precision: str = "f64"):
        self.sim = O.MpmRef(capacity, n_grid, grid_lim, precision)
        self.n_grid, self.capacity = n_grid, capacity
        self.inv_dx = float(n_grid / grid_lim)
        self._active = capacity

    def scatter(self, dt):
        self.sim.scatter(dt)

    def finish(self, dt, lo, hi):
        self.sim.finish(dt, lo, hi)

    def planes(self, a, b):
        return torch.from_numpy(self.sim.planes_get(a, b))

    def planes_add(self, a, b, t):
        self.sim.planes_add(a, b, t.numpy())

    @property
    def active(self):
        return self._active

    def set_active(self, n):
        self.sim.set_active(n)
        self._active = n

    def get(self, name):
        w = O.WIDTH[O.FIELDS[name]]
        return torch.from_numpy(np.asarray(self.sim.get(name), dtype=np.float64).reshape(self._active, w))

    def records(self):
        return torch.cat([self.get(name) for name, _ in FIELDS], dim=1)

    def set_records(self, rec):
        n = rec.shape[0]
        self.set_active(n)
        c = 0
        for name, w in FIELDS:
            self.sim.set(name, rec[:, c:c + w].numpy())
            c += w


def make_scene(n=600, n_grid=16, grid_lim=1.0, seed=0):
    """Jelly block drifting in +x across the middle of the domain with gravity and a bounding box."""
    rng = np.random.default_rng(seed)
    x = rng.uniform([0.30, 0.30, 0.30], [0.62, 0.70, 0.70], size=(n, 3))
    v = np.tile(np.array([1.5, 0.0, 0.2]), (n, 1)) + rng.normal(0, 0.05, size=(n, 3))
    dx = grid_lim / n_grid
    fields = dict(X=x, V=v, VOL=np.full(n, dx ** 3 / 6.0), DENSITY=np.full(n, 1000.0), E=rng.uniform(2e4, 5e4, n), NU=np.full(n, 0.3),
                  F=np.tile(np.eye(3).reshape(-1), (n, 1)), F_TRIAL=np.tile(np.eye(3).reshape(-1), (n, 1)),
                  MATERIAL=np.zeros(n), SELECTION=np.zeros(n))
    return fields


def load_scene(sim, fields, idx=None):
    """Fills the oracle `sim` with the particles `idx` of the scene (all if None) and finishes the setup."""
    n = len(fields["X"]) if idx is None else len(idx)
    sim.set_active(n)
    for k, v in fields.items():
        a = np.asarray(v, dtype=np.float64)
        sim.set(k, a if idx is None else a[idx])
    sim.compute_mass()
    sim.compute_mu_lam()
    sim.set_params(g=(0.0, 0.0, -9.8), grid_v_damping_scale=0.9999)
    sim.add_bc(O.BC_BBOX)
