"""GPU bring-up of the U-Net executor: layer-wise parity against the oracle, then timing.
Run under gpurun; writes gpurun_out/unet_check.log."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_ref as O
from pixie_b200.unet import RegressionUNet, SegmentationUNet

def log(*a):
    print(*a, flush=True)

def layerwise(C, G, precision, names_channels):
    seg, reg = O.build_pair(C, G, seed=0)
    x = O.synthetic_features(1, C, G, seed=1)
    acts = {}
    def hook(name):
        def f(m, i, o): acts[name] = o.detach()
        return f
    for n, m in reg.named_modules():
        if n in names_channels: m.register_forward_hook(hook(n))
    with torch.no_grad():
        y_ref = reg(x)
    mine = RegressionUNet(feature_channels=C, grid_size=G, out_channels=3, max_batch=1, precision=precision, **O.DEFAULT_CFG).to("cuda:0")
    mine.load_state_dict(reg.state_dict())
    y = mine(x.cuda()).cpu()
    mine.check()
    log(f"--- C={C} G={G} precision={precision}: out max|ref|={y_ref.abs().max():.3f} max_abs_err={(y-y_ref).abs().max():.3e} rms={(y-y_ref).pow(2).mean().sqrt():.3e}")
    for n, (ch, sp) in names_channels.items():
        try:
            a = mine.debug_fetch(n, ch, sp)
        except Exception as e:
            log(f"   {n:32s} fetch failed: {e}"); continue
        r = acts[n]
        log(f"   {n:32s} max|ref|={r.abs().max():8.3f} err={(a-r).abs().max():.3e}")
    return (y - y_ref).abs().max().item()

def main():
    torch.backends.cudnn.allow_tf32 = False
    names16 = {
        "projector.net.0": (128, 16), "projector.net.3": (128, 16), "projector.net.6": (32, 16),
        "unet.input_blocks.0": (64, 16), "unet.input_blocks.1.0": (64, 16), "unet.input_blocks.4.0": (64, 8),
        "unet.input_blocks.9.0": (128, 4), "unet.input_blocks.12.0": (128, 2), "unet.input_blocks.13.0": (256, 2),
        "unet.middle_block.0": (256, 2), "unet.middle_block.1": (256, 2), "unet.middle_block.2": (256, 2),
        "unet.output_blocks.0.0": (256, 2), "unet.output_blocks.3.1": (256, 4), "unet.output_blocks.7.1": (128, 8),
        "unet.output_blocks.11.1": (64, 16), "unet.output_blocks.15.0": (64, 16),
    }
    layerwise(128, 16, "fp16", names16)
    names32 = {k: (c, s * 2) for k, (c, s) in names16.items()}
    layerwise(512, 32, "fp16", names32)
    layerwise(512, 32, "fp16x3", names32)

    # timing at the benchmark size
    C, G = 512, 64
    seg, reg = O.build_pair(C, G, seed=0)
    x16 = (torch.randn(1, G, G, G, C, generator=torch.Generator().manual_seed(1)) * 0.05).to(torch.float16)
    for prec in ("fp16", "fp16x3"):
        mine = RegressionUNet(feature_channels=C, grid_size=G, out_channels=3, max_batch=1, precision=prec, **O.DEFAULT_CFG).to("cuda:0")
        mine.load_state_dict(reg.state_dict())
        xd = x16.cuda()
        y = mine.forward_channels_last_f16(xd); torch.cuda.synchronize(); mine.check()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): mine.forward_channels_last_f16(xd)
        e0.record()
        for _ in range(10): mine.forward_channels_last_f16(xd)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = mine.flops()
        log(f"TIME reg 64^3x512 {prec}: {ms:.3f} ms/forward  {fl/ms*1e-9:.1f} TFLOP/s algorithmic  launches={mine.launch_count()}")
        if prec == "fp16":
            t = time.time()
            with torch.no_grad():
                y_ref = reg(x16.float().permute(0, 4, 1, 2, 3).contiguous())
            log(f"CPU oracle 64^3x512: {time.time()-t:.1f}s on {torch.get_num_threads()} threads")
            log(f"64^3 parity {prec}: max_abs_err={(y.cpu()-y_ref).abs().max():.3e} rms={(y.cpu()-y_ref).pow(2).mean().sqrt():.3e}")
        else:
            log(f"64^3 parity {prec}: max_abs_err={(y.cpu()-y_ref).abs().max():.3e} rms={(y.cpu()-y_ref).pow(2).mean().sqrt():.3e}")
        del mine
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
