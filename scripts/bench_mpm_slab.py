"""BASELINE.json configs[4]: ONE 1M-particle / 256^3 MPM scene, slab-decomposed over the ranks of a torchrun job
(NCCL ghost-plane exchange + particle migration, pixie_b200/mpm_slab.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_mpm_slab.py --particles 1000000 --grid 256 --steps 200 --warmup 20

Prints one JSON line on rank 0: aggregate particle-steps/s (all particles x substeps / max-over-ranks device time).
With N = 1 the same script times the undivided scene through the same split-substep entry points (no exchange).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--dt", type=float, default=1e-4)
    ap.add_argument("--migrate-every", type=int, default=16)
    args = ap.parse_args()

    import torch.distributed as dist
    from pixie_b200.mpm_slab import CudaSlabBackend, DistSlabDriver, SlabRank, slab_bounds
    from pixie_b200.mpm_solver_warp import MPM_Simulator_WARP
    from pixie_b200.synthetic import synthetic_scene

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    n, G, lim = args.particles, args.grid, 2.0
    sc = synthetic_scene(n, G, seed=0, materials=(0,))          # identical on every rank (seeded)
    base = (sc["x"][:, 0].astype(np.float32) * np.float32(G / lim) - np.float32(0.5)).astype(np.int32)
    x0, x1 = slab_bounds(G, world, rank)
    lo = -10 ** 9 if rank == 0 else x0
    hi = 10 ** 9 if rank == world - 1 else x1
    idx = np.where((base >= lo) & (base < hi))[0]
    cap = n if world == 1 else min(n, int(1.5 * n / world) + 1024)
    if len(idx) > cap:
        cap = len(idx) + 1024

    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        s = MPM_Simulator_WARP(cap, n_grid=G, grid_lim=lim, device=dev)
        m = len(idx)

        def put(fid, arr, dtype=torch.float32):
            t = s._t[fid]
            t.view(cap, t.numel() // cap)[:m] = torch.as_tensor(np.asarray(arr)[idx].reshape(m, -1), dtype=dtype, device=dev)

        for fid, key in (("X", "x"), ("V", "v"), ("VOL", "vol"), ("DENSITY", "density"), ("E", "E"), ("NU", "nu")):
            put(fid, sc[key])
        put("MATERIAL", sc["material"], torch.int32)
        s.mpm_model.gravitational_accelaration = (0.0, 0.0, -9.8)
        s.mpm_model.grid_v_damping_scale = 0.9999
        s._push_params()
        from pixie_b200 import _lib
        lib = _lib.require_device()
        _lib.check(lib.pixie_mpm_compute_mass(s._handle, s._stream()))
        _lib.check(lib.pixie_mpm_compute_mu_lam(s._handle, s._stream()))
        s.add_bounding_box()
        s.set_velocity_on_cuboid(point=[1.0, 1.0, 0.62], size=[0.51, 0.51, 0.04], velocity=[0, 0, 0])
    r = SlabRank(CudaSlabBackend(s, m), rank, world, slack=1, migrate_every=args.migrate_every, ids=torch.from_numpy(idx.astype(np.int64)))
    drv = DistSlabDriver(r) if world > 1 else None

    def substep():
        if drv is not None:
            drv.substep(args.dt)
        else:
            r.scatter(args.dt)
            r.finish(args.dt)

    for _ in range(args.warmup):
        substep()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        substep()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    cnt = torch.tensor([float(r.b.active)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t.item())
        algo_bytes = 212.0 * n + 56.0 * G ** 3
        print(json.dumps({
            "metric": "mpm_particle_steps_per_s", "value": n * args.steps / (ms * 1e-3), "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "scaling": "strong",
            "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4]: one {n}-particle scene on a {G}^3 grid, x-slab decomposition over {world} GPU(s), "
                                   f"overlap planes exchanged every substep, migration every {args.migrate_every} substeps",
                       "parallelism": f"slab{world}", "particles_after": int(cnt.item())},
            "roofline": {"bound": "hbm", "achieved": algo_bytes * args.steps / (ms * 1e-3) / 1e9, "unit": "GB/s",
                         "note": "algorithmic bytes 212*Np + 56*Ng per substep over the whole job"},
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
