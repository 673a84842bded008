"""Slab-decomposed MLS-MPM rollout: one big scene over several GPUs (BASELINE.json configs[4]).

The reference runs one `MPM_Simulator_WARP` on "cuda:0" (mpm_solver_warp.py:47); it has no multi-GPU path. This module
shards ONE simulation along x: rank r owns grid planes [x0, x1) and the particles whose stencil base plane
(`int(x/dx - 0.5)`, mpm_utils.py:344) lies there. One substep is

    scatter   every rank scatters its particles into its own full-size {mv, m} grid         (pixie_mpm_substep_scatter)
    exchange  neighbours swap the partial sums of the planes both of them touch and add      (NCCL send/recv, 4 planes/face)
    finish    every rank normalises / applies the BCs on its owned + overlap planes and
              gathers back to its particles                                                  (pixie_mpm_substep_finish)

and every `migrate_every` substeps particles whose base plane left [x0, x1) move to the neighbour (packed records over
send/recv, live prefix of the bound arrays shrinks / grows). `slack` is how many planes a particle may drift outside
its slab between two migrations; the overlap with the right neighbour is [x1 - slack, x1 + 2 + slack) because a particle
touches planes base .. base+2. After the exchange both neighbours hold the COMPLETE sums on the overlap, so the grid
update there is computed redundantly and no second exchange is needed.

Restrictions (checked): slabs must be at least 2 + 2*slack planes wide; boundary conditions that carry per-particle masks
(impulses, velocity translation / rotation) are not migrated and are rejected.

The orchestration is backend-agnostic: `CudaSlabBackend` drives the C ABI; the CPU tests drive the same orchestration
with a CPU test double (tests/slab_backends.py), in one process and over gloo.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib


# ------------------------------------------------------------------------------------------- backends
class CudaSlabBackend:
    """One rank's solver: a `MPM_Simulator_WARP` shim created with `capacity` particles of which a prefix is live."""

    #: per-particle fields that migrate with a particle: (C-ABI field name, width)
    FIELDS = [("X", 3), ("V", 3), ("F", 9), ("F_TRIAL", 9), ("C", 9), ("STRESS", 9), ("R", 9), ("COV", 6), ("INIT_COV", 6),
              ("VOL", 1), ("MASS", 1), ("DENSITY", 1), ("MATERIAL", 1), ("SELECTION", 1), ("E", 1), ("NU", 1), ("MU", 1),
              ("LAM", 1), ("BULK", 1), ("YIELD", 1)]

    def __init__(self, solver, n_active: int):
        self.solver = solver
        self.lib = _lib.require_device()
        self.n_grid = int(solver.mpm_model.n_grid)
        self.capacity = int(solver.n_particles)
        self.device = solver._device
        self.inv_dx = float(solver.mpm_model.inv_dx)
        if solver._masks:
            raise ValueError("slab-decomposed runs do not migrate per-particle BC masks (impulses, velocity modifiers)")
        n = self.n_grid
        with torch.cuda.device(self.device):
            self.grid = torch.zeros((n, n * n * 4), dtype=torch.float32, device=self.device)
        self._check(self.lib.pixie_mpm_bind_grid(solver._handle, C.c_void_p(self.grid.data_ptr())))
        self._active = -1
        self._slab = None
        self.set_active(n_active)

    def _check(self, rc: int):
        if rc != 0:
            raise _lib.PixieError(self.lib.pixie_last_error().decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- substep halves
    def scatter(self, dt: float):
        with torch.cuda.device(self.device):
            self._check(self.lib.pixie_mpm_substep_scatter(self.solver._handle, float(dt), self._stream()))

    def finish(self, dt: float, lo: int, hi: int):
        with torch.cuda.device(self.device):
            if (lo, hi) != self._slab:
                self._check(self.lib.pixie_mpm_set_slab(self.solver._handle, int(lo), int(hi)))
                self._slab = (lo, hi)
            self._check(self.lib.pixie_mpm_substep_finish(self.solver._handle, float(dt), self._stream()))

    # -- grid planes {mv.xyz, m}
    def planes(self, a: int, b: int) -> torch.Tensor:
        return self.grid[a:b].clone()

    def planes_view(self, a: int, b: int) -> torch.Tensor:
        """No copy: x is the slowest grid dimension, so a plane range is one contiguous block (valid until the next add)."""
        return self.grid[a:b]

    def planes_add(self, a: int, b: int, t: torch.Tensor):
        self.grid[a:b] += t.to(self.grid.dtype)

    # -- particles
    @property
    def active(self) -> int:
        return self._active

    def set_active(self, n: int):
        if n > self.capacity:
            raise RuntimeError(f"slab holds {n} particles but was created with capacity {self.capacity}")
        if n != self._active:
            self._check(self.lib.pixie_mpm_set_active_count(self.solver._handle, int(n)))
            self._active = int(n)

    def get(self, name: str) -> torch.Tensor:
        t = self.solver._t[name]
        return t.view(self.capacity, t.numel() // self.capacity)[: self._active]

    def records(self) -> torch.Tensor:
        """[n_active, W] float32; integer fields are bit-cast, not converted."""
        cols = []
        for name, w in self.FIELDS:
            t = self.get(name)
            cols.append(t.view(torch.float32) if t.dtype == torch.int32 else t)
        return torch.cat(cols, dim=1)

    def records_at(self, index: torch.Tensor) -> torch.Tensor:
        """Records of the particles `index` only (migration touches a few thousand of them, not the whole slab)."""
        cols = []
        for name, w in self.FIELDS:
            t = self.get(name)[index]
            cols.append(t.view(torch.float32) if t.dtype == torch.int32 else t)
        return torch.cat(cols, dim=1)

    def compact_and_append(self, keep_index: torch.Tensor, arrivals: Optional[torch.Tensor]):
        """Live prefix <- particles `keep_index` (in order) followed by the `arrivals` records."""
        n_keep = int(keep_index.shape[0])
        n_new = 0 if arrivals is None else int(arrivals.shape[0])
        old = self._active
        if n_keep + n_new > self.capacity:
            raise RuntimeError(f"slab would hold {n_keep + n_new} particles but was created with capacity {self.capacity}")
        c = 0
        for name, w in self.FIELDS:
            full = self.solver._t[name].view(self.capacity, w)
            if n_keep != old:
                full[:n_keep] = full[:old][keep_index]
            if n_new:
                src = arrivals[:, c:c + w].contiguous()
                full[n_keep:n_keep + n_new] = src.view(torch.int32) if full.dtype == torch.int32 else src
            c += w
        self.set_active(n_keep + n_new)

    def set_records(self, rec: torch.Tensor):
        n = rec.shape[0]
        self.set_active(n)
        c = 0
        for name, w in self.FIELDS:
            dst = self.solver._t[name].view(self.capacity, w)
            src = rec[:, c:c + w].contiguous()
            dst[:n] = src.view(torch.int32) if dst.dtype == torch.int32 else src
            c += w


# ------------------------------------------------------------------------------------------- orchestration
def slab_bounds(n_grid: int, world: int, rank: int) -> Tuple[int, int]:
    return rank * n_grid // world, (rank + 1) * n_grid // world


class SlabRank:
    """Phase methods of one rank; a driver (`DistSlabDriver` or `LocalSlabCluster`) sequences them."""

    def __init__(self, backend, rank: int, world: int, slack: int = 1, migrate_every: int = 8, ids: Optional[torch.Tensor] = None):
        self.b, self.rank, self.world, self.slack, self.migrate_every = backend, rank, world, slack, migrate_every
        n = backend.n_grid
        self.x0, self.x1 = slab_bounds(n, world, rank)
        if world > 1 and (self.x1 - self.x0) < 2 + 2 * slack:
            raise ValueError(f"slab [{self.x0}, {self.x1}) is narrower than 2 + 2*slack = {2 + 2 * slack} planes")
        self.has_left, self.has_right = rank > 0, rank < world - 1
        # planes shared with the neighbours
        self.left_ov = (max(0, self.x0 - slack), min(n, self.x0 + 2 + slack)) if self.has_left else None
        self.right_ov = (max(0, self.x1 - slack), min(n, self.x1 + 2 + slack)) if self.has_right else None
        # planes this rank updates: owned + what its (drifted) particles can touch
        self.lo = max(0, self.x0 - slack) if self.has_left else 0
        self.hi = min(n, self.x1 + 2 + slack) if self.has_right else n
        self.steps = 0
        #: global particle ids travel with the records so results can be compared with a single-domain run
        self.ids = ids if ids is not None else torch.arange(backend.active, dtype=torch.int64)

    # -- substep phases
    def scatter(self, dt: float):
        self.b.scatter(dt)

    def snapshot(self):
        """Partial sums on the overlaps, taken BEFORE anything is added."""
        left = self.b.planes(*self.left_ov) if self.has_left else None
        right = self.b.planes(*self.right_ov) if self.has_right else None
        return left, right

    def snapshot_views(self):
        """Same without copies, for drivers that finish sending before they accumulate (backends without views copy)."""
        pv = getattr(self.b, "planes_view", self.b.planes)
        left = pv(*self.left_ov) if self.has_left else None
        right = pv(*self.right_ov) if self.has_right else None
        return left, right

    def accumulate(self, from_left, from_right):
        if self.has_left:
            self.b.planes_add(self.left_ov[0], self.left_ov[1], from_left)
        if self.has_right:
            self.b.planes_add(self.right_ov[0], self.right_ov[1], from_right)

    def finish(self, dt: float):
        self.b.finish(dt, self.lo, self.hi)
        self.steps += 1

    def due_for_migration(self) -> bool:
        return self.world > 1 and self.steps % self.migrate_every == 0

    # -- migration phases
    def migrate_collect(self):
        """Splits the live particles into stay / to_left / to_right. Returns (stay_index, (rec, ids) left, (rec, ids) right)."""
        x = self.b.get("X")[:, 0]
        # base plane exactly as the kernels compute it: float32 product, truncation toward zero (mpm_utils.py:344-346)
        base = (x.to(torch.float32) * torch.tensor(self.b.inv_dx, dtype=torch.float32, device=x.device) - 0.5).to(torch.int32)
        go_left = (base < self.x0) if self.has_left else torch.zeros_like(base, dtype=torch.bool)
        go_right = (base >= self.x1) if self.has_right else torch.zeros_like(base, dtype=torch.bool)
        stay_idx = torch.nonzero(~(go_left | go_right)).flatten()
        il, ir = torch.nonzero(go_left).flatten(), torch.nonzero(go_right).flatten()
        ids = self.ids.to(x.device)
        if hasattr(self.b, "records_at"):
            pack = lambda i: (self.b.records_at(i), ids[i])
        else:
            rec = self.b.records()
            pack = lambda i: (rec[i], ids[i])
        return stay_idx, pack(il), pack(ir)

    def migrate_apply(self, stay_idx, from_left, from_right):
        ids = self.ids.to(stay_idx.device)
        arrivals = [p for p in (from_left, from_right) if p is not None and p[0].shape[0] > 0]
        new_rec = torch.cat([p[0] for p in arrivals], dim=0) if arrivals else None
        if hasattr(self.b, "compact_and_append"):
            self.b.compact_and_append(stay_idx, new_rec)
        else:
            rec = self.b.records()[stay_idx]
            self.b.set_records(rec if new_rec is None else torch.cat([rec, new_rec.to(rec.dtype)], dim=0))
        self.ids = torch.cat([ids[stay_idx]] + [p[1].to(ids.device) for p in arrivals], dim=0)


class LocalSlabCluster:
    """All slabs in ONE process (tests, single-GPU emulation): the exchanges are plain hand-overs."""

    def __init__(self, ranks: Sequence[SlabRank]):
        self.ranks = list(ranks)

    def substep(self, dt: float):
        R = self.ranks
        for r in R:
            r.scatter(dt)
        snaps = [r.snapshot() for r in R]
        for i, r in enumerate(R):
            from_left = snaps[i - 1][1] if r.has_left else None       # left neighbour's right overlap = my left overlap
            from_right = snaps[i + 1][0] if r.has_right else None
            r.accumulate(from_left, from_right)
        for r in R:
            r.finish(dt)
        if R[0].due_for_migration():
            parts = [r.migrate_collect() for r in R]
            for i, r in enumerate(R):
                from_left = parts[i - 1][2] if r.has_left else None
                from_right = parts[i + 1][1] if r.has_right else None
                r.migrate_apply(parts[i][0], from_left, from_right)

    def gather(self, name: str) -> torch.Tensor:
        """Field `name` of every particle, ordered by global id."""
        vals = torch.cat([r.b.get(name).detach().cpu().to(torch.float64) for r in self.ranks], dim=0)
        ids = torch.cat([r.ids.cpu() for r in self.ranks], dim=0)
        out = torch.empty_like(vals)
        out[ids] = vals
        return out


class DistSlabDriver:
    """One rank of a `torch.distributed` job (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, rank_obj: SlabRank, group=None):
        import torch.distributed as dist
        self.r, self.dist, self.group = rank_obj, dist, group
        self._recv = None

    def _swap(self, to_left: Optional[torch.Tensor], to_right: Optional[torch.Tensor], like_left=None, like_right=None, reuse=False):
        """Symmetric neighbour exchange. `like_*` give the shape of what is received (default: what is sent); with
        `reuse` they ARE the receive buffers."""
        dist, r = self.dist, self.r
        ops, from_left, from_right = [], None, None
        if r.has_left:
            from_left = like_left if reuse else torch.empty_like(to_left if like_left is None else like_left)
            ops += [dist.P2POp(dist.isend, to_left.contiguous(), r.rank - 1, self.group), dist.P2POp(dist.irecv, from_left, r.rank - 1, self.group)]
        if r.has_right:
            from_right = like_right if reuse else torch.empty_like(to_right if like_right is None else like_right)
            ops += [dist.P2POp(dist.isend, to_right.contiguous(), r.rank + 1, self.group), dist.P2POp(dist.irecv, from_right, r.rank + 1, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return from_left, from_right

    def _swap_var(self, to_left, to_right):
        """Variable-length record exchange: counts first, then the records, then the global ids."""
        r = self.r
        ref = to_left if to_left is not None else to_right
        dev, width, dtype = ref[0].device, ref[0].shape[1], ref[0].dtype
        cnt = lambda p: torch.tensor([p[0].shape[0]], dtype=torch.int64, device=dev)
        nl, nr = self._swap(cnt(to_left) if r.has_left else None, cnt(to_right) if r.has_right else None)
        nl = int(nl.item()) if r.has_left else 0
        nr = int(nr.item()) if r.has_right else 0
        rec_l, rec_r = self._swap(to_left[0] if r.has_left else None, to_right[0] if r.has_right else None,
                                  torch.empty((nl, width), dtype=dtype, device=dev), torch.empty((nr, width), dtype=dtype, device=dev))
        ids_l, ids_r = self._swap(to_left[1].to(dev) if r.has_left else None, to_right[1].to(dev) if r.has_right else None,
                                  torch.empty((nl,), dtype=torch.int64, device=dev), torch.empty((nr,), dtype=torch.int64, device=dev))
        return ((rec_l, ids_l) if r.has_left else None), ((rec_r, ids_r) if r.has_right else None)

    def substep(self, dt: float):
        r = self.r
        r.scatter(dt)
        left, right = r.snapshot_views()
        if self._recv is None:          # overlap-plane receive buffers, allocated once
            self._recv = (torch.empty_like(left) if r.has_left else None, torch.empty_like(right) if r.has_right else None)
        from_left, from_right = self._swap(left, right, self._recv[0], self._recv[1], reuse=True)
        r.accumulate(from_left, from_right)
        r.finish(dt)
        if r.due_for_migration():
            stay, go_left, go_right = r.migrate_collect()
            from_left, from_right = self._swap_var(go_left if r.has_left else None, go_right if r.has_right else None)
            r.migrate_apply(stay, from_left, from_right)

    def gather(self, name: str, dst: int = 0) -> Optional[torch.Tensor]:
        """Field `name` of every particle ordered by global id, on rank `dst` (None elsewhere)."""
        dist, r = self.dist, self.r
        vals = r.b.get(name).detach().cpu().to(torch.float64).contiguous()
        ids = r.ids.cpu()
        objs = [None] * r.world if r.rank == dst else None
        dist.gather_object((vals, ids), objs, dst=dst, group=self.group)
        if r.rank != dst:
            return None
        v = torch.cat([o[0] for o in objs], dim=0)
        i = torch.cat([o[1] for o in objs], dim=0)
        out = torch.empty_like(v)
        out[i] = v
        return out
