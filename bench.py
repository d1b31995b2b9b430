#!/usr/bin/env python
"""Benchmark of the DCGAN G+D train step (BASELINE.json metric: train-step images/s at 1/2/4/8 B200).

    python bench.py --gpus 1 --steps 20 --warmup 5            # this framework (sm_100a kernels)
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # data parallel

Workload (config.workload): DCGAN 3x64x64, nz=100, ngf=ndf=64, 512 images per GPU per step, random-init
weights, synthetic U[-1,1) images -- BASELINE.json configs[2] at N=1 (batch 512) growing to configs[3] at
N=8 (global batch 4096); BatchNorm statistics and gradients are reduced over the global batch (weak
scaling).  A "step" is everything in the reference's train/dcgan_trainer.py:155-189: four D forwards, one
G forward, all backward sweeps incl. the gradient-penalty pass, both Adam updates, all random draws.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with the batch already in HBM (CUDA
events, max over ranks); `e2e` = the same step driven through the public trainer API with the batch in
pinned host memory (H2D copy + D2H read of the step's losses inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 512
METRIC = "dcgan_train_step_images_per_sec"
UNIT = "images/s"
# algorithmic FLOPs per image per step (observable convolutions only, nc = 3)
FLOP_PER_IMAGE = 2.690e9
WORKLOAD = ("DCGAN 3x64x64 nz=100 ngf=ndf=64, 512 images/GPU/step (BASELINE configs[2] at N=1 .. configs[3] "
            "at N=8), full G+D step incl. gradient penalty + Adam, SyncBN over the global batch")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU per step")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the FID-eval / input-pipeline side measurements")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-op device time table")
    ap.add_argument("--kernel-table", action="store_true", help="print per-kernel device time (CUPTI via torch.profiler)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tf_burst": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_sustained": 1400.0, "tf_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled every 2 ms (the timed
    region of a default run is well under a second, shorter than one `nvidia-smi` process start); when the NVML
    binding is missing the `nvidia-smi --query-gpu` line of B200_PROFILING.md is used instead."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.seen, self.stop_flag, self.how = index, [], [], set(), False, "none"
        self.nv = self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx.append(int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            self.nv, self.how = pynvml, "nvml, 2 ms poll"
        except Exception:
            self.nv = None

    def _poll_nvml(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for n, b in bits.items():
            if r & b:
                self.seen.add(n)

    def _poll_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                              str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
        c = [t.strip() for t in out.split(",")]
        if len(c) >= 6 and c[0].isdigit():
            self.sm.append(int(c[0]))
            if c[1].isdigit():
                self.mx.append(int(c[1]))
            for i, n in enumerate(self.NAMES):
                if c[2 + i].lower().startswith("active"):
                    self.seen.add(n)
            self.how = "nvidia-smi"

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self._poll_nvml()
                else:
                    self._poll_smi()
            except Exception:
                pass
            time.sleep(0.002 if self.nv is not None else 0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": [n for n in self.NAMES if n in self.seen], "samples": len(sm), "how": self.how}


# -------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's step, timed on the host cores
# -------------------------------------------------------------------------------------------------------
def time_cpu_port(batch, steps, warmup):
    import torch
    from oracle import models, steps as osteps
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g, d = models.build("DCGAN", seed=12345)
    og, od = osteps.make_optimizers(g, d, 2e-4)
    real = osteps.make_real(batch, n_steps=1)[0]
    rng = osteps.make_rng(batch, n_steps=1, seed=1)[0]
    for _ in range(warmup):
        osteps.dcgan_step(g, d, og, od, real, rng)
    t0 = time.perf_counter()
    for _ in range(steps):
        osteps.dcgan_step(g, d, og, od, real, rng)
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 128
    v, ms, cores = time_cpu_port(sample_batch, args.steps, max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{sample_batch}-image slice of the per-GPU batch per step"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle port of train/dcgan_trainer.py:155-189 (torch CPU, {cores} threads), "
                                       f"{args.steps} steps of {sample_batch} images, anomaly detection off"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference is pure Python on torch; it has no installable package (no setup.py / pyproject) and "
                    "does not exist on the GPU box, so this arm times the oracle port (bit-exact vs the reference, "
                    "tests/test_oracle_golden.py)"}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------------
# per-op device timing (CUDA events on the launching stream) for the roofline of the dominant kernel
# -------------------------------------------------------------------------------------------------------
class OpTimer:
    CONV = ("conv_down", "conv_up", "conv_wgrad")

    def __init__(self, ops, torch):
        self.ops, self.torch, self.rec, self.saved = ops, torch, [], {}

    def __enter__(self):
        names = [n for n in dir(self.ops) if callable(getattr(self.ops, n)) and not n.startswith("_") and
                 n not in ("dt", "L", "check", "wgrad_workspace_bytes", "edge_wgrad_workspace_bytes", "img_alloc")]
        for n in names:
            fn = getattr(self.ops, n)
            if getattr(fn, "__module__", "") != self.ops.__name__:
                continue
            self.saved[n] = fn

            def wrap(*a, __fn=fn, __n=n, **k):
                e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
                e0.record()
                r = __fn(*a, **k)
                e1.record()
                self.rec.append((__n, self._flops(__n, a, k), e0, e1))
                return r
            setattr(self.ops, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, fn in self.saved.items():
            setattr(self.ops, n, fn)

    @staticmethod
    def _flops(name, a, k):
        if name == "conv_down":       # (x_large, w_down, out_small, stats, Ca, Cb)
            B, Hs, Ws = a[2].shape[:3]
            return name + f"[{a[4]}x{a[5]}@{Hs}]", 2.0 * B * Hs * Ws * 16 * a[4] * a[5]
        if name == "conv_up":         # (x_small, w_up, out_large, stats, Ca, Cb)
            B, Hs, Ws = a[0].shape[:3]
            return name + f"[{a[4]}x{a[5]}@{Hs}]", 2.0 * B * Hs * Ws * 16 * a[4] * a[5]
        if name == "conv_wgrad":      # (small, large, dw4, workspace, Ca, Cb, accumulate)
            B, Hs, Ws = a[0].shape[:3]
            return name + f"[{a[4]}x{a[5]}@{Hs}]", 2.0 * B * Hs * Ws * 16 * a[4] * a[5]
        return name, 0.0

    def table(self):
        self.torch.cuda.synchronize()
        agg = {}
        for name, (key, fl), e0, e1 in self.rec:
            t = agg.setdefault(key, {"op": name, "ms": 0.0, "calls": 0, "flop": 0.0})
            t["ms"] += e0.elapsed_time(e1)
            t["calls"] += 1
            t["flop"] += fl
        return agg


def secondary_paths(torch, trainer, dev):
    """SURVEY.md 8f rows measured beside the headline (device time, CUDA events, after a warm-up pass):
    * fid_eval: BASELINE configs[4] on one GPU, scaled to 4096 samples: generator -> fused resize/normalise -> Inception-v3
      pool3 features on our tcgen05 conv kernels -> 2048 x 2048 covariance (ops.feature_moments);
    * input_pipeline: the reference's Resize(64)/ToTensor/Normalize loader on a CIFAR-shaped uint8 set resident in HBM."""
    import numpy as np
    from jck_generation_b200 import ops
    from jck_generation_b200.inception import InceptionV3
    from jck_generation_b200.preprocess.device_pipeline import DeviceImageLoader
    from torchvision import models
    out = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    torch.manual_seed(12345)
    net = models.inception_v3(weights=None, aux_logits=True, init_weights=False)
    ext = InceptionV3(net.state_dict(), feature="pool3", device=dev)
    n, bsz = 4096, 128
    z = torch.randn(n, 100, 1, 1, device=dev)

    def fid_pass():
        feats = []
        with torch.no_grad():
            for i in range(0, n, bsz):
                feats.append(ext.forward_generated(trainer.model_g(z[i:i + bsz]).float()))
        return ops.feature_moments(torch.cat(feats).contiguous())
    fid_pass()
    torch.cuda.synchronize()
    ev[0].record()
    fid_pass()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    out["fid_eval"] = {"images_per_s": n / (ms * 1e-3), "ms": ms, "samples": n, "feature": "pool3 (2048-d)",
                       "inception_tflops": 11.42e9 * n / (ms * 1e-3) / 1e12,
                       "what": "G forward + Inception-v3 (94 tcgen05 implicit-GEMM convs, one CUDA graph per 128 images) + "
                               "2048x2048 covariance, random-init weights"}
    del ext
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, (50000, 32, 32, 3), dtype=np.uint8)
    loader = DeviceImageLoader(data, None, 512, 64, [0.5] * 3, [0.5] * 3, shuffle=True)
    for _ in loader:
        pass
    torch.cuda.synchronize()
    ev[0].record()
    cnt = 0
    for x, _ in loader:
        cnt += x.shape[0]
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    out["input_pipeline"] = {"images_per_s": cnt / (ms * 1e-3), "ms_per_epoch": ms, "samples": cnt,
                             "what": "one epoch of a 50k x 32x32x3 uint8 set: seeded permutation gather + Pillow-exact "
                                     "bilinear 32->64 + ToTensor/Normalize -> NCHW fp32, batches of 512"}
    return out


def run_b200(args):
    import torch
    import __graft_entry__ as entry
    entry.build()
    from jck_generation_b200 import _lib, ops, parallel
    from jck_generation_b200.model import DCGAN
    from jck_generation_b200.train.dcgan_trainer import DCGANTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:      # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__), *sys.argv[1:]]
        sys.exit(subprocess.call(cmd))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch

    class _Data:
        def get_data_loader(self):
            return [], None
    targs = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="bench", log_file=0, batch_size=B,
                               num_worker=0, dtype=args.dtype, cuda_graph=0 if args.no_graph else 1, metrics=0,
                               save_path=os.path.join(ROOT, "gpurun_out", "bench_save"))
    torch.manual_seed(12345)
    trainer = DCGANTrainer(targs, DCGAN.Generator(), DCGAN.Discriminator(), _Data())
    comm = trainer.comm
    rank = comm.rank
    step = trainer.step
    gen = torch.Generator().manual_seed(12345 + rank)
    host_real = (torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1).pin_memory()
    real = host_real.to(dev)

    use_graph = trainer.use_graph
    launches_per_step = None
    if use_graph:
        c0 = _lib.launch_count()
        step.capture(B)
        launches_per_step = (_lib.launch_count() - c0) // 3      # 2 warm-up runs + 1 captured run

    def one_step():
        return step.replay(real) if use_graph else step.run(real)

    def timed(fn, n):
        comm.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = _lib.launch_count()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        comm.barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if comm.world_size > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t) / n, _lib.launch_count() - c0

    for _ in range(max(3, args.warmup)):
        one_step()
    sampler = ClockSampler(local)
    sampler.start()
    ms, counted = timed(one_step, args.steps)
    clocks = sampler.summary()
    launches = launches_per_step * args.steps if use_graph else counted
    total_images = B * comm.world_size
    value = total_images / (ms * 1e-3)

    # end to end through the public API: pinned host batch -> H2D -> step -> D2H of the step's scalars.  The
    # trainer's own input pipeline (train/prefetch.py) copies batch i+1 on a side stream while step i runs, as
    # DCGANTrainer.train() does; every timed step still pays one H2D of a full batch and one D2H of its losses.
    from jck_generation_b200.train.prefetch import DevicePrefetcher

    def host_batches():
        while True:
            yield (host_real,)
    feed = iter(DevicePrefetcher(host_batches(), dev))

    def e2e_step():
        (x,) = next(feed)
        s = trainer.train_step(x)
        return s.cpu()
    for _ in range(3):
        e2e_step()
    e2e_ms, _ = timed(e2e_step, args.steps)
    e2e = {"value": total_images / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": host_real.numel() * 4, "d2h_bytes_per_step": 4 * 2 * 4,
           "api": "DCGANTrainer.train_step(real) fed by the trainer's DevicePrefetcher from a pinned host batch "
                  "(H2D of the next batch overlaps the running step); losses read back every step"}

    # dominant kernel roofline: eager pass with CUDA events around every op
    # (weight gradients back on the main stream for this pass: per-kernel times must not include a concurrent kernel)
    pk = peaks()
    side = (step.eg.wgrad_stream, step.ed.wgrad_stream)
    step.eg.wgrad_stream = step.ed.wgrad_stream = None
    with OpTimer(ops, torch) as ot:
        for _ in range(2):
            step.run(real)
        tab = ot.table()
    step.eg.wgrad_stream, step.ed.wgrad_stream = side
    tot = sum(t["ms"] for t in tab.values()) or 1.0
    conv = {k: t for k, t in tab.items() if t["flop"] > 0}
    top = max(conv.items(), key=lambda kv: kv[1]["ms"])
    fam = {}
    for k, t in conv.items():
        f = fam.setdefault(t["op"], {"ms": 0.0, "flop": 0.0, "calls": 0})
        f["ms"] += t["ms"]; f["flop"] += t["flop"]; f["calls"] += t["calls"]
    tk, tv = top
    achieved = tv["flop"] / (tv["ms"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": tk, "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf_sustained"], "traffic": None, "peak_source": pk["source"],
                "avg_launch_ms": tv["ms"] / tv["calls"], "share_of_step": tv["ms"] / tot,
                "step_tensor_frac": FLOP_PER_IMAGE * value / (comm.world_size * pk["tf_sustained"] * 1e12),
                "families": {k: {"share": f["ms"] / tot, "tflops": f["flop"] / (f["ms"] * 1e-3) / 1e12} for k, f in fam.items()}}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tr = json.load(f).get(tk)
        if tr:
            roofline["traffic"] = tr["avg_bytes_per_launch"]
            roofline["traffic_source"] = "profiles/r01_ncu_kernels.md: DRAM read+write bytes per launch (ncu --set full), " \
                                         f"algorithmic {tr['algorithmic_avg_bytes_per_launch']:.3g} B"
    except OSError:
        pass
    if args.kernel_table and rank == 0:
        # per-kernel device time from CUPTI (torch.profiler), warm caches, eager launches
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step.run(real)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot_us = sum(e.device_time_total for e in rows)
        print(f"# kernel table: {tot_us / 3e3:.3f} ms/step of device time over 3 eager steps", file=sys.stderr)
        for e in rows[:45]:
            print(f"# {e.device_time_total / 3e3:8.3f} ms/step {e.count // 3:4d} calls {e.device_time_total / max(e.count, 1):8.1f} us  "
                  f"{e.key[:100]}", file=sys.stderr)
    if args.profile_ops and rank == 0:
        for k, t in sorted(tab.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"# {t['ms'] / 2:9.3f} ms/step  {t['calls'] // 2:4d} calls  {k}", file=sys.stderr)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": comm.world_size, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": total_images,
                       "parallelism": f"dp{comm.world_size}", "cuda_graph": bool(use_graph),
                       "l2": "working set per step (~1.5 GB of activations at 512 images) exceeds the 126 MB L2; no flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    if comm.world_size == 1 and not args.no_cpu_baseline:
        v, cms, cores = time_cpu_port(128, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"oracle port of the reference step, 3 steps of 128 images on {cores} host "
                                          f"threads ({cms:.0f} ms/step)"}
    if comm.world_size == 1 and not args.no_secondary:
        try:
            line["secondary"] = secondary_paths(torch, trainer, dev)
        except Exception as e:                      # the headline never depends on the secondary paths
            line["secondary"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if comm.world_size > 1:
        # tearing down a NCCL communicator that captured CUDA graphs still reference can block for minutes;
        # everything is measured and printed, so leave without the collective teardown
        comm.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
