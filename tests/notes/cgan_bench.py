"""GPU timing of the CGAN G+D step (BASELINE configs[1]-like: 3x64x64, 100 classes, batch 256, bf16), eager launches."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
entry.build()
from jck_generation_b200.model import CGAN
from jck_generation_b200.train.cgan_trainer import CGANTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


class _Data:
    idx_to_labels = {i: str(i) for i in range(100)}
    def get_data_loader(self):
        return [], None


args = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="cganbench", log_file=0, batch_size=B, num_worker=0,
                          dtype="bf16", cuda_graph=int(os.environ.get("GRAPH", "1")), metrics=0, save_path="/tmp/cgan_bench_save")
torch.manual_seed(12345)
tr = CGANTrainer(args, CGAN.Generator(), CGAN.Discriminator(), _Data())
real = (torch.rand(B, 3, 64, 64) * 2 - 1).cuda()
labels = torch.nn.functional.one_hot(torch.randint(0, 100, (B,)), 100).cuda()
for _ in range(3):
    tr.train_step(real, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    s = tr.train_step(real, labels)
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
ms = e0.elapsed_time(e1) / n
print(f"cgan step B={B}: {ms:.3f} ms device ({B / ms * 1e3:.0f} images/s), host wall {wall:.3f} ms/step, scalars {s.flatten().tolist()}")
if "--table" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            tr.train_step(real, labels)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"kernel table: {tot / 3e3:.3f} ms/step device time")
    for e in rows[:25]:
        print(f"  {e.device_time_total / 3e3:8.3f} ms/step {e.count // 3:4d} calls {e.device_time_total / max(e.count, 1):8.1f} us  {e.key[:90]}")
