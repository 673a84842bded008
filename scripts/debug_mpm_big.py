"""compute-sanitizer target: the configs[4] scene (1M particles, 256^3) for a few substeps, and a 2-slab cluster on one device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch

mode = sys.argv[1] if len(sys.argv) > 1 else "big"
if mode == "big":
    from scripts.gpu_mpm_perf import make
    n, ng = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000, int(sys.argv[3]) if len(sys.argv) > 3 else 256
    s = make(n=n, ng=ng)
    s.p2g2p_n(int(sys.argv[4]) if len(sys.argv) > 4 else 6, 1e-4)
    torch.cuda.synchronize()
    x = s.mpm_state.particle_x.numpy()
    print("big ok", np.isfinite(x).all(), x.min(), x.max())
else:
    import test_slab_mpm as T
    from slab_backends import make_scene
    from pixie_b200.mpm_slab import LocalSlabCluster
    fields = make_scene(T.N, T.G, T.LIM)
    ranks, _ = T._cuda_cluster(fields, 2, 4)
    cl = LocalSlabCluster(ranks)
    for _ in range(10):
        cl.substep(T.DT)
    torch.cuda.synchronize()
    for r in ranks:
        r.check_device_error()
    print("slab ok", [r.b.active for r in ranks])
