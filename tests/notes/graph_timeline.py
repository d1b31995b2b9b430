"""Timeline of ONE CUDA-graph replay of the DCGAN step (CUPTI kernel records via torch.profiler): start, duration, stream of
every kernel, idle gaps on the main stream and overlap with the side stream.  Run with JCK_PDL=0 so that a kernel's
duration does not include the time it is parked behind its predecessor.  Usage: python tests/notes/graph_timeline.py [batch]
(under torchrun: the data-parallel step, NCCL / peer-memory kernels included, as rank 0 sees it)."""
import argparse, json, os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
entry.build()
from jck_generation_b200.model import DCGAN
from jck_generation_b200.train.dcgan_trainer import DCGANTrainer
from torch.profiler import ProfilerActivity, profile

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512


class _Data:
    def get_data_loader(self):
        return [], None


args = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="tl", log_file=0, batch_size=B, num_worker=0,
                          dtype="bf16", cuda_graph=1, metrics=0, save_path="/tmp/tl_save")
torch.manual_seed(12345)
tr = DCGANTrainer(args, DCGAN.Generator(), DCGAN.Discriminator(), _Data())
real = (torch.rand(B, 3, 64, 64) * 2 - 1).cuda()
for _ in range(5):
    tr.train_step(real)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(real)
    torch.cuda.synchronize()
if tr.comm.rank != 0:            # under torchrun: rank 0 reports
    tr.comm.barrier()
    os._exit(0)
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
streams = sorted({e["args"].get("stream") for e in ev})
print(f"{len(ev)} device activities, streams {streams}, span {(ev[-1]['ts'] + ev[-1]['dur'] - t0) / 1e3:.3f} ms")
main = max(streams, key=lambda s: sum(1 for e in ev if e["args"].get("stream") == s))
end_main = t0
gap_total = 0.0
for e in ev:
    s = e["args"].get("stream")
    name = e["name"].replace("void jck::(anonymous namespace)::", "").replace("jck::(anonymous namespace)::", "")[:60]
    gap = ""
    if s == main:
        g = e["ts"] - end_main
        if g > 0.5:
            gap = f"  <- idle {g:.1f} us"
            gap_total += g
        end_main = max(end_main, e["ts"] + e["dur"])
    print(f"{(e['ts'] - t0):9.1f} {e['dur']:7.1f} {'M' if s == main else 'S'} {name}{gap}")
print(f"idle on the main stream: {gap_total:.1f} us")
if tr.comm.world_size > 1:
    sys.stdout.flush()
    tr.comm.barrier()
    os._exit(0)
