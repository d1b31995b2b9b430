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


def shared_config(world, batch):
    """`config` of BOTH arms (this framework and --impl reference): the workload, nothing implementation-specific --
    what is specific to a run (CUDA graph, SyncBN transport, the CPU sample) is reported under `run` / `cpu_baseline`."""
    return {"workload": WORKLOAD, "per_gpu_batch": batch, "global_batch": batch * world, "parallelism": f"dp{world}",
            "l2": "working set per step (~1.5 GB of activations at 512 images) exceeds the 126 MB L2; no flush"}


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the side measurements (CGAN step, fp32 step, library bar, FID eval ...)")
    ap.add_argument("--secondary-only", action="store_true", help="(internal) run only the side measurements, print their JSON")
    ap.add_argument("--quick", action="store_true", help="(A/B timing) only the device-timed step: no e2e, rooflines or side paths")
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
def time_cpu_port(batch, steps, warmup, anomaly=False):
    """The reference's step on the host cores (oracle port: the reference's torch operators in the reference's order,
    bit-exact vs the unmodified trainer -- tests/test_oracle_golden.py), all host threads.  `anomaly`: with
    torch.autograd.set_detect_anomaly(True), as the reference's main.py:28 leaves it."""
    import torch
    from oracle import models, steps as osteps
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g, d = models.build("DCGAN", seed=12345)
    og, od = osteps.make_optimizers(g, d, 2e-4)
    real = osteps.make_real(batch, n_steps=1)[0]
    rng = osteps.make_rng(batch, n_steps=1, seed=1)[0]
    prev = torch.is_anomaly_enabled()
    torch.autograd.set_detect_anomaly(bool(anomaly))
    try:
        for _ in range(warmup):
            osteps.dcgan_step(g, d, og, od, real, rng)
        t0 = time.perf_counter()
        for _ in range(steps):
            osteps.dcgan_step(g, d, og, od, real, rng)
        dt = (time.perf_counter() - t0) / steps
    finally:
        torch.autograd.set_detect_anomaly(prev)
    return batch / dt, dt * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every step is ONE per-GPU batch of the GPU arm (512 images, ~1.3 s on 16 host threads): the same batch size the
    # GPU arm's kernels see, so the two arms' `config` are identical; under N > 1 it is a per-GPU-sized sample of the
    # global batch (the reference is a single process)
    sample_batch = args.batch
    v, ms, cores = time_cpu_port(sample_batch, args.steps, max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(args.gpus, args.batch),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle port of train/dcgan_trainer.py:155-189 (torch CPU, {cores} threads), "
                                       f"{args.steps} steps of {sample_batch} images (one per-GPU batch each), anomaly "
                                       f"detection off"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference is pure Python on torch; it has no installable package (no setup.py / pyproject) and "
                    "does not exist on the GPU box, so this arm times the oracle port (bit-exact vs the reference, "
                    "tests/test_oracle_golden.py)"}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------------
# per-op device timing (CUDA events on the launching stream) for the rooflines
# -------------------------------------------------------------------------------------------------------
class OpTimer:
    """Wraps every op of jck_generation_b200.ops with a CUDA-event pair (on the launching stream) and attaches the op's
    ALGORITHMIC work: FLOPs for the convolutions (2 * MACs), bytes for the streaming passes (each tensor read or written
    once: DESIGN.md section 4's per-element figures x the elements of the call)."""

    def __init__(self, ops, torch):
        self.ops, self.torch, self.rec, self.saved = ops, torch, [], {}

    def __enter__(self):
        names = [n for n in dir(self.ops) if callable(getattr(self.ops, n)) and not n.startswith("_") and
                 n not in ("dt", "L", "check", "wgrad_workspace_bytes", "edge_wgrad_workspace_bytes", "img_alloc",
                           "gemm_tc_workspace_bytes")]
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
                self.rec.append((__n, self._work(__n, a, k), e0, e1))
                return r
            setattr(self.ops, n, wrap)
        return self

    def __exit__(self, *exc):
        for n, fn in self.saved.items():
            setattr(self.ops, n, fn)

    @staticmethod
    def _work(name, a, k):
        """(key, flops, bytes) of one call"""
        nb = lambda t: t.numel() * t.element_size()
        if name in ("conv_down", "conv_down_bnbwd"):       # (x_large, w_down, out_small | y_saved ..., Ca, Cb)
            out = a[2] if name == "conv_down" else a[6]
            Ca, Cb = (a[4], a[5]) if name == "conv_down" else (a[8], a[9])
            B, Hs, Ws = out.shape[:3]
            return f"{name}[{Ca}x{Cb}@{Hs}]", 2.0 * B * Hs * Ws * 16 * Ca * Cb, nb(a[0]) + nb(out)
        if name in ("conv_up", "conv_up_bnbwd"):           # (x_small, w_up, out_large | y_saved ..., Ca, Cb)
            out = a[2] if name == "conv_up" else a[6]
            Ca, Cb = (a[4], a[5]) if name == "conv_up" else (a[8], a[9])
            B, Hs, Ws = a[0].shape[:3]
            return f"{name}[{Ca}x{Cb}@{Hs}]", 2.0 * B * Hs * Ws * 16 * Ca * Cb, nb(a[0]) + nb(out)
        if name == "conv_wgrad":                           # (small, large, dw4, workspace, Ca, Cb, accumulate)
            B, Hs, Ws = a[0].shape[:3]
            return f"{name}[{a[4]}x{a[5]}@{Hs}]", 2.0 * B * Hs * Ws * 16 * a[4] * a[5], nb(a[0]) + nb(a[1])
        if name == "bn_act_fwd":                           # (y, ss, a, C, ...): read y, write a
            return f"{name}[{a[3]}ch,{nb(a[0]) >> 20}MiB]", 0.0, nb(a[0]) + nb(a[2])
        if name == "bn_act_bwd_reduce":                    # (da, y, ss, mr, sums, C, ...): read da, y
            return f"{name}[{a[5]}ch,{nb(a[0]) >> 20}MiB]", 0.0, nb(a[0]) + nb(a[1])
        if name == "bn_act_bwd_apply":                     # (da, y, ss, mr, gamma, sums, dy, C, ...): read da, y; write dy
            return f"{name}[{a[7]}ch,{nb(a[0]) >> 20}MiB]", 0.0, nb(a[0]) + nb(a[1]) + nb(a[6])
        if name == "edge_down_img":                        # (img_p4, w, out_small, ...): read image, write out
            return f"{name}[{nb(a[2]) >> 20}MiB]", 0.0, nb(a[0]) + nb(a[2])
        if name in ("edge_up_scatter", "edge_up"):         # (x_small, w, img_p4, Ca): read x, write image
            return f"{name}[{nb(a[0]) >> 20}MiB]", 0.0, nb(a[0]) + nb(a[2])
        if name == "adam":                                 # 4 reads + 3 writes of fp32
            return name, 0.0, 7 * nb(a[0])
        return name, 0.0, 0.0

    def table(self):
        self.torch.cuda.synchronize()
        agg = {}
        for name, (key, fl, by), e0, e1 in self.rec:
            t = agg.setdefault(key, {"op": name, "ms": 0.0, "calls": 0, "flop": 0.0, "bytes": 0.0})
            t["ms"] += e0.elapsed_time(e1)
            t["calls"] += 1
            t["flop"] += fl
            t["bytes"] += by
        return agg


def _ev_time(torch, fn, n):
    """mean device ms of n calls of fn (CUDA events on the current stream, synchronised on both sides)"""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def secondary_paths(torch, trainer, dev, args):
    """Measured beside the headline (device time, CUDA events, after warm-up), single GPU:
    * cgan_step: BASELINE configs[1] (CGAN, batch 256) at the config's shape (1 channel, 10 classes; 28x28 sources are
      resized to 64 by the preprocessor as the reference resizes CIFAR, cgan_data_preprocessor.py:51) and at the reference's
      native shape (3 channels, 100 classes); the step back-propagates the gradient penalty (second order);
    * fp32_step: the headline step in the exact-parity fp32 mode (CUDA-core FMA kernels, the mode the <= 1e-4 tests run);
    * library_bar: the reference's own step UNCHANGED on this GPU through torch eager + cuDNN / cuBLAS (utils.py:4-8
      device='cuda'), true fp32 and torch's default TF32 convolutions;
    * cpu_anomaly_on: the CPU baseline with torch.autograd.set_detect_anomaly(True), as the reference's main.py:28 runs;
    * fid_eval: BASELINE configs[4] on one GPU, scaled to 4096 samples; input_pipeline: the reference's loader on the device."""
    import numpy as np
    from jck_generation_b200 import ops
    out = {}

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:                      # noqa: BLE001 -- the headline never depends on a secondary path
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    def cgan():
        from jck_generation_b200.model import CGAN
        from jck_generation_b200.train.cgan_trainer import CGANTrainer
        res = {}
        for tag, nc, ncls in (("mnist_shape_1ch_10cls", 1, 10), ("reference_shape_3ch_100cls", 3, 100)):
            class _D:
                idx_to_labels = {i: str(i) for i in range(ncls)}

                def get_data_loader(self):
                    return [], None
            B = 256
            a = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="bench_cgan", log_file=0, batch_size=B,
                                   num_worker=0, dtype="bf16", cuda_graph=1, metrics=0,
                                   save_path=os.path.join(ROOT, "gpurun_out", "bench_save"))
            torch.manual_seed(12345)
            tr = CGANTrainer(a, CGAN.Generator(nc=nc, n_classes=ncls), CGAN.Discriminator(nc=nc, n_classes=ncls), _D())
            real = (torch.rand(B, nc, 64, 64) * 2 - 1).to(dev)
            labels = torch.nn.functional.one_hot(torch.randint(0, ncls, (B,)), ncls).to(dev)
            for _ in range(3):
                tr.train_step(real, labels)
            ms = _ev_time(torch, lambda: tr.train_step(real, labels), 20)
            res[tag] = {"ms_per_step": ms, "images_per_s": B / (ms * 1e-3), "batch": B, "dtype": "bf16", "cuda_graph": True}
            del tr
        return res
    guarded("cgan_step", cgan)

    def fp32():
        from jck_generation_b200.model import DCGAN
        from jck_generation_b200.train.dcgan_trainer import DCGANTrainer

        class _D:
            def get_data_loader(self):
                return [], None
        B = args.batch
        a = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="bench_fp32", log_file=0, batch_size=B, num_worker=0,
                               dtype="fp32", cuda_graph=0, metrics=0, save_path=os.path.join(ROOT, "gpurun_out", "bench_save"))
        torch.manual_seed(12345)
        tr = DCGANTrainer(a, DCGAN.Generator(), DCGAN.Discriminator(), _D())
        real = (torch.rand(B, 3, 64, 64) * 2 - 1).to(dev)
        tr.train_step(real)
        ms = _ev_time(torch, lambda: tr.train_step(real), 3)
        return {"ms_per_step": ms, "images_per_s": B / (ms * 1e-3), "batch": B, "dtype": "f32",
                "what": "CUDA-core fp32 FMA kernels (the <= 1e-4 parity mode), eager launches"}
    guarded("fp32_step", fp32)

    def library_bar():
        from oracle import models, steps as osteps
        B, res = args.batch, {}
        keep = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
        try:
            for tag, tf32 in (("fp32", False), ("tf32_convs_torch_default", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = True
                g, d = models.build("DCGAN", seed=12345)
                g, d = g.to(dev), d.to(dev)
                og, od = osteps.make_optimizers(g, d, 2e-4)
                real = osteps.make_real(B, n_steps=1)[0].to(dev)
                rng = {k: v.to(dev) for k, v in osteps.make_rng(B, n_steps=1, seed=1)[0].items()}
                with torch.device(dev):
                    for _ in range(3):
                        osteps.dcgan_step(g, d, og, od, real, rng)
                    ms = _ev_time(torch, lambda: osteps.dcgan_step(g, d, og, od, real, rng), 8)
                res[tag] = {"ms_per_step": ms, "images_per_s": B / (ms * 1e-3)}
                del g, d, og, od
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = keep
        res["what"] = (f"the reference's step (oracle port: its torch operators in its order) on this GPU through torch eager + "
                       f"cuDNN/cuBLAS, {B} images, 5 host syncs per step as the reference's .item() calls")
        return res
    guarded("library_bar", library_bar)

    def anomaly():
        v, ms, cores = time_cpu_port(args.batch, 2, 1, anomaly=True)
        return {"value": v, "unit": UNIT, "ms_per_step": ms, "cores": cores,
                "what": f"CPU baseline with torch.autograd.set_detect_anomaly(True) (reference main.py:28), 2 steps of {args.batch}"}
    guarded("cpu_anomaly_on", anomaly)

    def fid():
        from jck_generation_b200.inception import InceptionV3
        from torchvision import models
        torch.manual_seed(12345)
        net = models.inception_v3(weights=None, aux_logits=True, init_weights=False)
        n, bsz = 4096, 128
        z = torch.randn(n, 100, 1, 1, device=dev)
        res = {}
        for prec in ("bf16", "split"):
            ext = InceptionV3(net.state_dict(), feature="pool3", device=dev, precision=prec)

            def fid_pass():
                feats = []
                with torch.no_grad():
                    for i in range(0, n, bsz):
                        feats.append(ext.forward_generated(trainer.model_g(z[i:i + bsz]).float()))
                return ops.feature_moments(torch.cat(feats).contiguous())
            fid_pass()
            ms = _ev_time(torch, fid_pass, 1)
            res[prec] = {"images_per_s": n / (ms * 1e-3), "ms": ms, "inception_tflops": 11.42e9 * n / (ms * 1e-3) / 1e12}
            del ext
            torch.cuda.empty_cache()
        res["bf16"].update({"samples": n, "feature": "pool3 (2048-d)", "split_precision": res["split"],
                            "what": "G forward + Inception-v3 (94 tcgen05 implicit-GEMM convs, one CUDA graph per 128 images) + "
                                    "2048x2048 covariance, random-init weights; split_precision = the fp32-grade mode (hi + lo "
                                    "bf16 planes, 3 MMAs per product; TFLOP/s counts the algorithmic FLOPs once), Metrics' default"})
        return res["bf16"]
    guarded("fid_eval", fid)

    def pipeline():
        from jck_generation_b200.preprocess.device_pipeline import DeviceImageLoader
        rng = np.random.default_rng(0)
        data = rng.integers(0, 256, (50000, 32, 32, 3), dtype=np.uint8)
        loader = DeviceImageLoader(data, None, 512, 64, [0.5] * 3, [0.5] * 3, shuffle=True)
        for _ in loader:
            pass
        cnt = [0]

        def epoch():
            for x, _ in loader:
                cnt[0] += x.shape[0]
        ms = _ev_time(torch, epoch, 1)
        return {"images_per_s": cnt[0] / (ms * 1e-3), "ms_per_epoch": ms, "samples": cnt[0],
                "what": "one epoch of a 50k x 32x32x3 uint8 set: seeded permutation gather + Pillow-exact bilinear 32->64 + "
                        "ToTensor/Normalize -> NCHW fp32, batches of 512"}
    guarded("input_pipeline", pipeline)
    return out


def make_trainer(torch, args, batch, dtype=None):
    from jck_generation_b200.model import DCGAN
    from jck_generation_b200.train.dcgan_trainer import DCGANTrainer

    class _Data:
        def get_data_loader(self):
            return [], None
    targs = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="bench", log_file=0, batch_size=batch,
                               num_worker=0, dtype=dtype or args.dtype, cuda_graph=0 if args.no_graph else 1, metrics=0,
                               save_path=os.path.join(ROOT, "gpurun_out", "bench_save"))
    torch.manual_seed(12345)
    return DCGANTrainer(targs, DCGAN.Generator(), DCGAN.Discriminator(), _Data())


def run_b200(args):
    import torch
    import __graft_entry__ as entry
    entry.build()
    from jck_generation_b200 import _lib, ops, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:      # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__), *sys.argv[1:]]
        sys.exit(subprocess.call(cmd))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch

    trainer = make_trainer(torch, args, B)
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

    def one_step():
        return step.replay(real) if use_graph else step.run(real)

    for _ in range(max(3, args.warmup)):
        one_step()
    sampler = ClockSampler(local)
    sampler.start()
    ms, counted = timed(one_step, args.steps)
    clocks = sampler.summary()
    launches = launches_per_step * args.steps if use_graph else counted
    total_images = B * comm.world_size
    value = total_images / (ms * 1e-3)

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": comm.world_size, "steps": args.steps,
                              "ms_per_step": ms, "clocks": clocks, "quick": True}), flush=True)
        if comm.world_size > 1:
            comm.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        return

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

    # rooflines: eager pass with CUDA events around every op, everything on ONE stream (weight gradients and the
    # penalty sweep in line: a per-kernel time must not include a concurrent kernel).  The dominant op is the one with the
    # largest total time over ALL ops; the top streaming (HBM-bound) op is reported next to it.
    pk = peaks()
    side = (step.eg.wgrad_stream, step.ed.wgrad_stream, step.gp_stream)
    step.eg.wgrad_stream = step.ed.wgrad_stream = step.gp_stream = None
    with OpTimer(ops, torch) as ot:
        # The host needs 50-100 us per eager launch (ctypes call, tensor-map encode, allocations, two event records), many
        # kernels run 5-40 us: with an empty stream every start event would fire before its kernel has even been launched
        # and the interval would measure the host.  So: time the host side of one eager step, then park the stream behind
        # a device spin 1.5x that long before each measured step -- the whole step is queued before the GPU reaches it and
        # the event pairs bracket back-to-back device work only.
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        step.run(real)
        host_s = time.perf_counter() - h0
        torch.cuda.synchronize()
        ot.rec.clear()
        spin = int(1.5 * host_s * 2.0e9) + 4_000_000
        for _ in range(2):
            torch.cuda._sleep(spin)
            step.run(real)
        tab = ot.table()
    step.eg.wgrad_stream, step.ed.wgrad_stream, step.gp_stream = side
    tot = sum(t["ms"] for t in tab.values()) or 1.0
    traffic_db = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            traffic_db = json.load(f)
    except OSError:
        pass

    def roof(key, t):
        tensor = t["flop"] > 0
        if tensor:
            ach, peak, unit = t["flop"] / (t["ms"] * 1e-3) / 1e12, pk["tf_sustained"], "TFLOP/s"
        else:
            ach, peak, unit = t["bytes"] / (t["ms"] * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
        r = {"bound": "tensor" if tensor else "hbm", "kernel": key, "achieved": ach, "peak": peak, "unit": unit,
             "frac": ach / peak, "traffic": None, "peak_source": pk["source"], "avg_launch_ms": t["ms"] / t["calls"],
             "calls_per_step": t["calls"] // 2, "share_of_step": t["ms"] / tot,
             "algorithmic_per_launch": (t["flop"] if tensor else t["bytes"]) / t["calls"]}
        tr = traffic_db.get(key)
        if tr:
            r["traffic"] = tr["dram_bytes_per_launch"]
            r["traffic_source"] = tr.get("source", "profiles/r02_ncu_kernels.md (ncu --set full, dram__bytes_read + write)")
        return r
    top_key, top_t = max(tab.items(), key=lambda kv: kv[1]["ms"])
    roofline = roof(top_key, top_t)
    roofline["step_tensor_frac"] = FLOP_PER_IMAGE * value / (comm.world_size * pk["tf_sustained"] * 1e12)
    fam = {}
    for k, t in tab.items():
        if t["flop"] > 0 or t["bytes"] > 0:
            f = fam.setdefault(t["op"], {"ms": 0.0, "flop": 0.0, "bytes": 0.0, "calls": 0})
            f["ms"] += t["ms"]; f["flop"] += t["flop"]; f["bytes"] += t["bytes"]; f["calls"] += t["calls"]
    roofline["families"] = {k: ({"share": f["ms"] / tot, "tflops": f["flop"] / (f["ms"] * 1e-3) / 1e12} if f["flop"] > 0 else
                                {"share": f["ms"] / tot, "gbs": f["bytes"] / (f["ms"] * 1e-3) / 1e9}) for k, f in fam.items()}
    hbm_ops = {k: t for k, t in tab.items() if t["flop"] == 0 and t["bytes"] > 0}
    roofline_hbm = roof(*max(hbm_ops.items(), key=lambda kv: kv[1]["ms"])) if hbm_ops else None
    tensor_ops = {k: t for k, t in tab.items() if t["flop"] > 0}
    roofline_tensor = roof(*max(tensor_ops.items(), key=lambda kv: kv[1]["ms"])) if tensor_ops else None

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
            w = (f"{t['flop'] / (t['ms'] * 1e-3) / 1e12:7.0f} TFLOP/s" if t["flop"] > 0 else
                 f"{t['bytes'] / (t['ms'] * 1e-3) / 1e9:7.0f} GB/s" if t["bytes"] > 0 else "")
            print(f"# {t['ms'] / 2:9.3f} ms/step  {t['calls'] // 2:4d} calls  {k:44s} {w}", file=sys.stderr)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": comm.world_size, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": shared_config(comm.world_size, B),
            "run": {"cuda_graph": bool(use_graph), "syncbn_transport": comm.transport,
                    "grad_exchange": ("none (one GPU)" if comm.world_size == 1 else
                                      f"NCCL all-reduce, {len(step.sync_d.buckets)} + {len(step.sync_g.buckets)} buckets started "
                                      "as their gradients become final"),
                    "penalty_sweep_stream": step.gp_stream is not None,
                    "launches_per_step": launches_per_step},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "roofline_hbm": roofline_hbm, "roofline_tensor": roofline_tensor}

    secondary = {}
    if comm.world_size > 1:
        # (1) W ranks == one rank on the global batch, outside the timed region; (2) BASELINE configs[2] as stated:
        # 512 images GLOBAL, i.e. 512 / N per GPU (strong scaling), same step, own CUDA graph
        from jck_generation_b200.train import dp_selfcheck
        try:
            line["dp_check"] = dp_selfcheck.run(comm, per_rank=32)
        except Exception as e:                      # noqa: BLE001
            line["dp_check"] = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]}
        comm.barrier()
        gb = 512
        if gb % comm.world_size == 0:
            b2 = gb // comm.world_size
            tr2 = make_trainer(torch, args, b2)
            real2 = real[:b2].contiguous()
            if tr2.use_graph:
                tr2.step.capture(b2)
            f2 = (lambda: tr2.step.replay(real2)) if tr2.use_graph else (lambda: tr2.step.run(real2))
            for _ in range(max(3, args.warmup)):
                f2()
            ms2, _ = timed(f2, args.steps)
            secondary["strong_scaling"] = {"global_batch": gb, "per_gpu_batch": b2, "ms_per_step": ms2,
                                           "images_per_s": gb / (ms2 * 1e-3),
                                           "what": "BASELINE configs[2]: DCGAN 3x64x64 batch 512 GLOBAL, data parallel with SyncBN"}
    if comm.world_size == 1 and not args.no_cpu_baseline:
        v, cms, cores = time_cpu_port(B, 5, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"oracle port of the reference step, 5 steps of {B} images (the GPU arm's batch) on "
                                          f"{cores} host threads ({cms:.0f} ms/step), anomaly detection off"}
    if comm.world_size == 1 and not args.no_secondary:
        # in a child process: a fault in a side path (they exercise other kernels: CGAN, Inception, fp32 mode) must not take
        # the headline down with it, and their allocations / cudnn autotuning must not disturb it either
        del trainer, step
        torch.cuda.empty_cache()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--secondary-only", "--batch", str(B), "--dtype", args.dtype],
                               capture_output=True, text=True, timeout=900)
            got = [l for l in r.stdout.splitlines() if l.startswith("{")]
            secondary.update(json.loads(got[-1]) if got else {"error": (r.stderr or r.stdout)[-300:]})
        except Exception as e:                      # noqa: BLE001
            secondary["error"] = f"{type(e).__name__}: {e}"[:300]
    if secondary:
        line["secondary"] = secondary
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


def run_secondary_only(args):
    import torch
    import __graft_entry__ as entry
    entry.build()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    trainer = make_trainer(torch, args, args.batch)
    print(json.dumps(secondary_paths(torch, trainer, dev, args)), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.secondary_only:
        run_secondary_only(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
