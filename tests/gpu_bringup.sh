#!/bin/bash
# First-light script for a GPU box: safe kernels in one process, tensor-core kernels one per process.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m tests.kernel_checks --no-tc > gpurun_out/kernels_simt.log 2>&1
echo "simt exit $?" >> gpurun_out/kernels_simt.log
timeout 900 python -m tests.kernel_checks --only-tc --isolate > gpurun_out/kernels_tc.log 2>&1
echo "tc exit $?" >> gpurun_out/kernels_tc.log
timeout 600 python - > gpurun_out/step_f32.log 2>&1 <<'PY'
import torch
from tests.parity import dcgan_step_parity
e = dcgan_step_parity(torch.float32, batch=4)
for k, v in sorted(e.items(), key=lambda kv: -kv[1]): print(f"{v:.3e} {k}")
PY
echo "f32 step exit $?" >> gpurun_out/step_f32.log
timeout 600 python - > gpurun_out/step_bf16.log 2>&1 <<'PY'
import torch
from tests.parity import dcgan_step_parity
e = dcgan_step_parity(torch.bfloat16, batch=8)
for k, v in sorted(e.items(), key=lambda kv: -kv[1]): print(f"{v:.3e} {k}")
PY
echo "bf16 step exit $?" >> gpurun_out/step_bf16.log
tail -3 gpurun_out/kernels_simt.log; tail -3 gpurun_out/kernels_tc.log; head -5 gpurun_out/step_f32.log; head -5 gpurun_out/step_bf16.log
