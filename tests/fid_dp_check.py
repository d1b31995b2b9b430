"""Sharded FID / IS evaluation on real GPUs (BASELINE configs[4]); run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 -m tests.fid_dp_check

W ranks, each extracting the Inception features of n/W generated samples and all-reducing the feature sums / Gram matrix,
must return the score and FID that ONE rank computes on all n samples (same kernels, so the features are bit-identical;
only the order of the fp32 moment sums differs)."""
import os
import sys

import numpy as np
import torch

from jck_generation_b200 import parallel
from jck_generation_b200.metrics import Metrics
from tests.incep_fixture import calibrated_inception


def main():
    comm = parallel.init_from_env()
    rank, W = comm.rank, comm.world_size
    work = f"/tmp/jck_fid_dp_{os.environ.get('MASTER_PORT', '0')}_{rank}"
    os.makedirs(os.path.join(work, "save/iception_v3"), exist_ok=True)
    os.chdir(work)
    torch.save(calibrated_inception(seed=1).state_dict(), "save/iception_v3/loss_bset.pt")
    g = torch.Generator().manual_seed(11)
    real = torch.utils.data.TensorDataset(torch.randn(256, 3, 64, 64, generator=g), torch.zeros(256, dtype=torch.long))
    fake = torch.tanh(torch.randn(512, 3, 64, 64, generator=g))
    ok = True
    for feature in ("logits", "pool3"):
        m = Metrics(real, feature=feature, comm=comm, allow_random_weights=True)
        s1, f1 = m.evaluate_generated(fake)                               # all rows on this rank
        sw, fw = m.evaluate_generated_sharded(parallel.shard_rows(fake, comm))
        scale = float(np.trace(np.cov(m.real_features.astype(np.float64), rowvar=False)))
        es = abs(sw - s1) / abs(s1) if feature == "logits" else 0.0
        ef = abs(fw - f1) / scale
        if rank == 0:
            print(f"{feature}: W={W} score {sw:.6f} vs {s1:.6f} (rel {es:.2e}); fid {fw:.5f} vs {f1:.5f} ({ef:.2e} of tr S)", flush=True)
        ok = ok and es < 1e-5 and ef < 2e-3
    comm.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("fid_dp_check", "OK" if ok else "FAILED", flush=True)
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
