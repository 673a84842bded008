"""Slab-decomposed MPM (BASELINE configs[4]): the decomposition must reproduce the single-domain run.
CPU: orchestration on the oracle test double, in one process and over gloo (world_size 2).
GPU: the same orchestration on two CUDA solvers sharing one device, against the single-domain CUDA run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pixie_b200.mpm_slab import DistSlabDriver, LocalSlabCluster, SlabRank, slab_bounds  # noqa: E402

N, G, LIM, DT = 600, 16, 1.0, 2e-3


def _owned(fields, world, rank):
    base = (fields["X"][:, 0].astype(np.float32) * np.float32(G / LIM) - np.float32(0.5)).astype(np.int32)
    x0, x1 = slab_bounds(G, world, rank)
    lo = -10 ** 9 if rank == 0 else x0
    hi = 10 ** 9 if rank == world - 1 else x1
    return np.where((base >= lo) & (base < hi))[0]


def _oracle_rank(fields, world, rank, migrate_every, slack=1, lazy_trigger=None):
    from slab_backends import OracleSlabBackend, load_scene
    idx = _owned(fields, world, rank)
    b = OracleSlabBackend(G, LIM, N, "f64")
    load_scene(b.sim, fields, idx)
    b._active = len(idx)
    return SlabRank(b, rank, world, slack=slack, migrate_every=migrate_every, ids=torch.from_numpy(idx.astype(np.int64)),
                    lazy_trigger=lazy_trigger)


def _reference(fields, steps):
    from slab_backends import O, load_scene
    ref = O.MpmRef(N, G, LIM, "f64")
    load_scene(ref, fields)
    ref.step(steps, DT)
    return ref


@pytest.mark.parametrize("world,migrate_every", [(2, 4), (3, 2), (2, 1)])
def test_local_cluster_matches_single_domain(world, migrate_every):
    from slab_backends import make_scene
    fields = make_scene(N, G, LIM)
    ranks = [_oracle_rank(fields, world, r, migrate_every) for r in range(world)]
    before = [r.b.active for r in ranks]
    cl = LocalSlabCluster(ranks)
    steps = 60
    for _ in range(steps):
        cl.substep(DT)
    ref = _reference(fields, steps)
    assert sum(r.b.active for r in ranks) == N
    assert [r.b.active for r in ranks] != before, "the scene must push particles across a slab face"
    for name in ("X", "V", "F_TRIAL", "C"):
        got = cl.gather(name).numpy().reshape(N, -1)
        want = np.asarray(ref.get(name)).reshape(N, -1)
        assert np.abs(got - want).max() < 1e-11, name    # f64: only the summation order at shared nodes differs


def test_lazy_migration_matches_single_domain_and_migrates_less():
    """Check points every 2 substeps, but particles only move once one of them is a whole plane past the first slack plane."""
    from slab_backends import make_scene
    fields = make_scene(N, G, LIM)
    world, steps = 2, 80
    ranks = [_oracle_rank(fields, world, r, 2, slack=2, lazy_trigger=2) for r in range(world)]
    before = [r.b.active for r in ranks]
    cl = LocalSlabCluster(ranks)
    for _ in range(steps):
        cl.substep(DT)
    ref = _reference(fields, steps)
    assert sum(r.b.active for r in ranks) == N
    assert ranks[0].checks == steps // 2 and 1 <= ranks[0].migrations < ranks[0].checks // 2
    assert [r.b.active for r in ranks] != before
    for name in ("X", "V", "F_TRIAL"):
        got = cl.gather(name).numpy().reshape(N, -1)
        assert np.abs(got - np.asarray(ref.get(name)).reshape(N, -1)).max() < 1e-11, name
    with pytest.raises(ValueError):
        _oracle_rank(fields, world, 0, 2, slack=2, lazy_trigger=3)


def test_balanced_slab_bounds_properties():
    """Cuts at the particle quantiles: contiguous cover of [0, n_grid), every slab at least min_width planes, particle counts
    within a plane's worth of each other; the block of BASELINE configs[4] (particles on 40 % of the x range) must not leave
    ranks empty the way equal-width slabs do."""
    from pixie_b200.mpm_slab import balanced_slab_bounds
    rng = np.random.default_rng(0)
    n_grid, n = 256, 200_000
    base = rng.integers(77, 179, size=n)                       # stencil base planes of a 0.4-wide block
    for world in (2, 3, 4, 8):
        b = balanced_slab_bounds(base, n_grid, world, 6)
        assert b[0][0] == 0 and b[-1][1] == n_grid and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(x1 - x0 >= 6 for x0, x1 in b)
        counts = [int(((base >= (x0 if i else -1)) & (base < (x1 if i < world - 1 else n_grid + 1))).sum()) for i, (x0, x1) in enumerate(b)]
        assert sum(counts) == n and min(counts) > 0
        per_plane = n / 102
        assert max(counts) - min(counts) <= 2 * per_plane + 1, (world, counts)
        equal = [int(((base >= r * n_grid // world) & (base < (r + 1) * n_grid // world)).sum()) for r in range(world)]
        assert max(counts) <= max(equal) + per_plane + 1                      # never worse than equal-width slabs (up to one plane)
    # a domain that cannot hold that many slabs of the minimum width is an error, not a silent squeeze
    with pytest.raises(ValueError):
        balanced_slab_bounds(base, 16, 4, 6)
    # degenerate input: all particles on one plane -> widths still legal, cover still complete
    b = balanced_slab_bounds(np.full(1000, 128), n_grid, 4, 6)
    assert b[0][0] == 0 and b[-1][1] == n_grid and all(x1 - x0 >= 6 for x0, x1 in b)


def test_slab_too_narrow_is_rejected():
    from slab_backends import OracleSlabBackend
    b = OracleSlabBackend(8, 1.0, 4, "f64")
    with pytest.raises(ValueError):
        SlabRank(b, 0, 4, slack=1)          # 2 planes per slab < 2 + 2*slack


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, steps, q, lazy=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    import torch.distributed as dist
    from slab_backends import make_scene
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fields = make_scene(N, G, LIM)
    drv = DistSlabDriver(_oracle_rank(fields, world, rank, 4, slack=2, lazy_trigger=2) if lazy else _oracle_rank(fields, world, rank, 4))
    for _ in range(steps):
        drv.substep(DT)
    x = drv.gather("X")
    f = drv.gather("F_TRIAL")
    q.put((rank, drv.r.b.active, None if x is None else x.numpy(), None if f is None else f.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lazy", [False, True])
def test_gloo_world2_matches_single_domain(lazy):
    from slab_backends import make_scene
    world, steps = 2, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, steps, q, lazy)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=180)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _reference(make_scene(N, G, LIM), steps)
    assert res[0][1] + res[1][1] == N
    assert res[1][2] is None
    assert np.abs(res[0][2].reshape(N, 3) - np.asarray(ref.get("X"))).max() < 1e-11
    assert np.abs(res[0][3].reshape(N, 3, 3) - np.asarray(ref.get("F_TRIAL"))).max() < 1e-11


# ------------------------------------------------------------------------------------------------ GPU
def _cuda_solver(fields, idx, capacity):
    from pixie_b200.mpm_solver_warp import MPM_Simulator_WARP
    dev = "cuda:0"
    s = MPM_Simulator_WARP(capacity, n_grid=G, grid_lim=LIM, device=dev)
    n = len(idx)

    def put(fid, arr, dtype=torch.float32):
        t = s._t[fid]
        if n:      # a slab may start empty (particles arrive by migration)
            t.view(capacity, -1)[:n] = torch.as_tensor(np.asarray(arr)[idx].reshape(n, -1), dtype=dtype, device=dev)

    for fid in ("X", "V", "F", "F_TRIAL", "VOL", "DENSITY", "E", "NU"):
        put(fid, fields[fid])
    put("MATERIAL", fields["MATERIAL"], torch.int32)
    put("SELECTION", fields["SELECTION"], torch.int32)
    s.mpm_model.gravitational_accelaration = (0.0, 0.0, -9.8)
    s.mpm_model.grid_v_damping_scale = 0.9999
    s._push_params()
    return s


def _cuda_cluster(fields, world, migrate_every, bounds=None, slack=1, lazy_trigger=None):
    import ctypes as C
    from pixie_b200 import _lib
    from pixie_b200.mpm_slab import FusedSlabBackend
    lib = _lib.require_device()

    def finish_setup(s):
        st = s._stream()
        _lib.check(lib.pixie_mpm_compute_mass(s._handle, st))
        _lib.check(lib.pixie_mpm_compute_mu_lam(s._handle, st))
        s.add_bounding_box()

    ranks = []
    for r in range(world):
        if bounds is None:
            idx = _owned(fields, world, r)
        else:
            base = (fields["X"][:, 0].astype(np.float32) * np.float32(G / LIM) - np.float32(0.5)).astype(np.int32)
            lo = -10 ** 9 if r == 0 else bounds[r][0]
            hi = 10 ** 9 if r == world - 1 else bounds[r][1]
            idx = np.where((base >= lo) & (base < hi))[0]
        s = _cuda_solver(fields, idx, N)
        finish_setup(s)
        ranks.append(SlabRank(FusedSlabBackend(s, len(idx)), r, world, slack=slack, migrate_every=migrate_every,
                              ids=torch.from_numpy(idx.astype(np.int64)), bounds=None if bounds is None else bounds[r],
                              lazy_trigger=lazy_trigger))
    return ranks, finish_setup


@pytest.mark.gpu
@pytest.mark.parametrize("world,migrate_every", [(2, 4), (3, 2), (4, 3)])
def test_cuda_slabs_match_single_domain_cuda(world, migrate_every):
    """`world` CUDA solvers (one per slab) on one device, overlap totals read by the halo kernel from the neighbours' exchange
    buffers (plain pointers inside one process): same trajectory as the undivided CUDA run and as the f64 oracle."""
    from slab_backends import make_scene
    fields = make_scene(N, G, LIM)
    steps = 60
    ranks, finish_setup = _cuda_cluster(fields, world, migrate_every,
                                        bounds=None if world != 3 else [(0, 6), (6, 10), (10, 16)])       # unequal slabs too
    whole = _cuda_solver(fields, np.arange(N), N)
    finish_setup(whole)
    whole.p2g2p_n(steps, DT)
    x_whole = whole._t["X"].view(N, 3).cpu().numpy().astype(np.float64)

    before = [r.b.active for r in ranks]
    cl = LocalSlabCluster(ranks)
    for _ in range(steps):
        cl.substep(DT)
    torch.cuda.synchronize()
    for r in ranks:
        r.check_device_error()
    assert sum(r.b.active for r in ranks) == N and [r.b.active for r in ranks] != before
    x_slab = cl.gather("X").numpy().reshape(N, 3)
    ref = _reference(fields, steps)
    x_ref = np.asarray(ref.get("X"))
    # fp32 atomics: both CUDA runs sit within the same distance of the f64 oracle, and of each other
    assert np.abs(x_slab - x_whole).max() < 2e-5
    assert np.abs(x_slab - x_ref).max() < 5e-5
    ft = cl.gather("F_TRIAL").numpy().reshape(N, 9)
    assert np.abs(ft - np.asarray(ref.get("F_TRIAL")).reshape(N, 9)).max() < 5e-4


@pytest.mark.gpu
def test_cuda_lazy_migration_matches_single_domain():
    """Migration checks by the excursion kernel (sorted positions, no write-back): same trajectory, fewer migrations."""
    from slab_backends import make_scene
    fields = make_scene(N, G, LIM)
    steps = 80
    ranks, finish_setup = _cuda_cluster(fields, 2, 2, slack=2, lazy_trigger=2)
    before = [r.b.active for r in ranks]
    cl = LocalSlabCluster(ranks)
    for _ in range(steps):
        cl.substep(DT)
    torch.cuda.synchronize()
    for r in ranks:
        r.check_device_error()
    assert ranks[0].checks == steps // 2 and 1 <= ranks[0].migrations < ranks[0].checks // 2
    assert sum(r.b.active for r in ranks) == N and [r.b.active for r in ranks] != before
    # the device excursion equals the one computed from the written-back positions
    for r in ranks:
        dev_e = int(r.b.excursion().item())
        x = r.b.get("X")[:, 0]
        base = (x * torch.tensor(r.b.inv_dx, dtype=torch.float32, device=x.device) - 0.5).to(torch.int32)
        want = max(0, (r.x0 - int(base.min())) if r.has_left else 0, (int(base.max()) - (r.x1 - 1)) if r.has_right else 0)
        assert dev_e == want
    ref = _reference(fields, steps)
    assert np.abs(cl.gather("X").numpy().reshape(N, 3) - np.asarray(ref.get("X"))).max() < 5e-5


@pytest.mark.gpu
def test_cuda_slab_drift_beyond_slack_is_reported():
    """A particle that out-runs the slack planes between two migrations must raise instead of silently corrupting the grid."""
    from slab_backends import make_scene
    fields = make_scene(N, G, LIM)
    fields["V"][:, 0] = 12.0                                  # 12 * 2e-3 * 16 = 0.38 cells per substep, 20 substeps without migration
    ranks, _ = _cuda_cluster(fields, 2, migrate_every=1000)
    cl = LocalSlabCluster(ranks)
    with pytest.raises(RuntimeError, match="drifted more than slack"):
        for _ in range(20):
            cl.substep(DT)
        torch.cuda.synchronize()
        for r in ranks:
            r.check_device_error()
