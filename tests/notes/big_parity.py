"""Dump every parity error of a train step at the benchmarked batch sizes (run on the GPU box):

    python -m tests.notes.big_parity [case ...] > gpurun_out/big_parity.json

cases: dcgan_bf16_128 dcgan_bf16_512 dcgan_f32_128 cgan_bf16_256 cgan_f32_64 offinit_bf16_128 envelope_128
The numbers printed here are what tests/test_gpu_big.py's tolerances were set from."""
import json
import sys
import time

import torch

from tests import parity


CASES = {
    "dcgan_bf16_128": lambda: parity.dcgan_step_parity(torch.bfloat16, batch=128),
    "dcgan_bf16_512": lambda: parity.dcgan_step_parity(torch.bfloat16, batch=512),
    "dcgan_f32_128": lambda: parity.dcgan_step_parity(torch.float32, batch=128),
    "cgan_bf16_256": lambda: parity.cgan_step_parity(torch.bfloat16, batch=256),
    "cgan_bf16_256_mnist": lambda: parity.cgan_step_parity(torch.bfloat16, batch=256, nc=1, n_classes=10),
    "cgan_f32_64": lambda: parity.cgan_step_parity(torch.float32, batch=64),
    "offinit_bf16_128_w10": lambda: parity.dcgan_step_parity(torch.bfloat16, batch=128, warm_steps=10),
    "offinit_bf16_128_w30": lambda: parity.dcgan_step_parity(torch.bfloat16, batch=128, warm_steps=30),
    "offinit_f32_128_w10": lambda: parity.dcgan_step_parity(torch.float32, batch=128, warm_steps=10),
    "envelope_128": lambda: parity.autocast_envelope(128),
    "envelope_8": lambda: parity.autocast_envelope(8),
    "envelope_128_w10": lambda: parity.autocast_envelope(128, warm_steps=10),
    "envelope_128_w30": lambda: parity.autocast_envelope(128, warm_steps=30),
}


def main():
    import __graft_entry__ as entry
    entry.build()
    names = sys.argv[1:] or list(CASES)
    out = {}
    for n in names:
        t0 = time.time()
        try:
            out[n] = CASES[n]()
        except Exception as e:      # noqa: BLE001
            out[n] = {"error": f"{type(e).__name__}: {e}"[:500]}
        out[n]["_seconds"] = time.time() - t0
        print(f"# {n}: {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
