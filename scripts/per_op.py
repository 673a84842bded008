"""Per-launch device times of one seg-network forward (kind, ms, GFLOP) — which launches the time goes to."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pixie_b200.inference import MaterialFieldPredictor

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
G, Cf = 64, 512
sd_seg, sd_reg = bench.make_state_dicts(Cf, G)
pred = MaterialFieldPredictor(feature_channels=Cf, grid_size=G, device="cuda:0", max_batch=1, precision=prec, **bench.UNET_CFG)
pred.load_state_dicts(sd_seg, sd_reg)
feat = bench.make_features(G, Cf, 1).cuda()
for _ in range(3):
    pred.predict(feat)
ops = pred.seg_network.profile(feat)
ops = pred.seg_network.profile(feat)
tot = {}
for i, (k, ms, fl) in enumerate(ops):
    tot[k] = tot.get(k, 0.0) + ms
    print(f"{i:4d} {k:10s} {ms*1e3:9.1f} us  {fl/1e9:8.2f} GF  {fl/ms/1e9 if ms > 0 and fl > 0 else 0:8.1f} TF/s")
print({k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
