"""Throughput of the Inception-v3 feature extractor (jck_generation_b200/inception.py) on one B200, next to torchvision
eager (cuDNN) on the same GPU as the library bar.  Usage: python tests/notes/incep_bench.py [batch] [iters]"""
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.incep_fixture import calibrated_inception
from jck_generation_b200.inception import InceptionV3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    model = calibrated_inception(seed=1, calib_batch=2)
    flops = [0]

    def hook(m, i, o):
        if isinstance(m, torch.nn.Conv2d):
            flops[0] += 2 * o.shape[1] * o.shape[2] * o.shape[3] * m.in_channels * m.kernel_size[0] * m.kernel_size[1]
        elif isinstance(m, torch.nn.Linear):
            flops[0] += 2 * m.in_features * m.out_features
    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
    with torch.no_grad():
        model(torch.zeros(1, 3, 299, 299))
    for h in hs:
        h.remove()
    gf = flops[0] / 1e9
    net = InceptionV3(model.state_dict(), device="cuda")
    fake = torch.tanh(torch.randn(B, 3, 64, 64, device="cuda"))
    for _ in range(2):
        net.forward_generated(fake)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        net.forward_generated(fake)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"jck inception: B={B} {ms:.2f} ms/batch {B / ms * 1e3:.0f} img/s {gf * B / ms:.1f} TFLOP/s ({gf:.2f} GFLOP/img)", flush=True)
    # per-layer timing of one forward (eager, events around each launch group)
    if os.environ.get("JCK_INCEP_LAYERS"):
        orig = net._conv
        net.use_graph = False
        times = []

        def timed(name, src, dst, c_off=0):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = orig(name, src, dst, c_off)
            b.record()
            times.append((name, a, b, src, net.convs[name]))
            return r
        net._conv = timed
        others = []
        for meth in ("_stem", "_pool", "_head"):
            fn = getattr(net, meth)

            def wrap(*a, __fn=fn, __m=meth, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = __fn(*a, **k)
                e1.record()
                others.append((__m, e0, e1))
                return r
            setattr(net, meth, wrap)
        net.forward_generated(fake)
        torch.cuda.synchronize()
        agg = {}
        for m, e0, e1 in others:
            agg[m] = agg.get(m, 0.0) + e0.elapsed_time(e1)
        print("  " + "  ".join(f"{m}: {t * 1e3:.0f} us" for m, t in agg.items()))
        for name, a, b, src, cv in times:
            t = a.elapsed_time(b)
            Ho = (src.H + 2 * cv.pad[0] - cv.kh) // cv.stride + 1
            Wo = (src.W + 2 * cv.pad[1] - cv.kw) // cv.stride + 1
            fl = 2.0 * B * Ho * Wo * cv.N * cv.C * cv.kh * cv.kw
            print(f"  {name:28s} {src.H:3d}x{src.W:<3d} C={cv.C:4d} N={cv.N:4d} k={cv.kh}x{cv.kw} {t * 1e3:8.1f} us {fl / t / 1e9:7.1f} TFLOP/s")
    # library bar: torchvision eager on the same GPU
    m = model.cuda().eval()
    x = torch.randn(B, 3, 299, 299, device="cuda")
    for name, ctx, mm, xx in (("fp32 (TF32 off)", torch.autocast("cuda", enabled=False), m, x),
                              ("bf16 autocast channels_last", torch.autocast("cuda", dtype=torch.bfloat16),
                               m.to(memory_format=torch.channels_last), x.to(memory_format=torch.channels_last))):
        with torch.no_grad(), ctx:
            for _ in range(2):
                mm(xx)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                mm(xx)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"torchvision eager {name}: {ms:.2f} ms/batch {B / ms * 1e3:.0f} img/s {gf * B / ms:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
