import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_ref as O
from pixie_b200.unet import RegressionUNet
C, G = 64, 16
_, reg = O.build_pair(C, G, seed=2)
x = O.synthetic_features(2, C, G, seed=4)
with torch.no_grad(): y_ref = reg(x)
for prec in ("fp16", "fp16x3"):
    net = RegressionUNet(feature_channels=C, grid_size=G, out_channels=3, max_batch=2, precision=prec, **O.DEFAULT_CFG).to("cuda:0")
    net.load_state_dict(reg.state_dict())
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(torch.float16)
    xd, xcd = x.cuda(), x_cl.cuda()
    outs = {}
    for name, fn in (("ncdhw1", lambda: net(xd)), ("ncdhw2", lambda: net(xd)), ("cl1", lambda: net.forward_channels_last_f16(xcd)),
                     ("cl2", lambda: net.forward_channels_last_f16(xcd)), ("ncdhw3", lambda: net(xd)), ("b0", lambda: net(xd[:1])), ("b1", lambda: net(xd[1:]))):
        outs[name] = fn().cpu()
    net.check()
    print(prec, "graph" if not os.environ.get("PIXIE_NO_GRAPH") else "nograph")
    for k, v in outs.items():
        ref = y_ref if v.shape[0] == 2 else (y_ref[:1] if k == "b0" else y_ref[1:])
        print(f"   {k:8s} err vs oracle {(v-ref).abs().max():.3e}   vs ncdhw1 {(v-outs['ncdhw1'][:v.shape[0]] if k!='b1' else v-outs['ncdhw1'][1:]).abs().max():.3e}")
