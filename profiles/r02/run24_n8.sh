set -x
mkdir -p gpurun_out
export JCK_COMM_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --steps 40 --warmup 5 --quick"
timeout 150 $TR --nproc-per-node 8 --master-port 29541 $B > gpurun_out/r2_n8_q_merge.log 2>&1; grep -h '^{' gpurun_out/r2_n8_q_merge.log | cut -c1-170
NCCL_ALGO=NVLS timeout 150 $TR --nproc-per-node 8 --master-port 29542 $B > gpurun_out/r2_n8_q_merge_nvls.log 2>&1; grep -h '^{' gpurun_out/r2_n8_q_merge_nvls.log | cut -c1-170; tail -2 gpurun_out/r2_n8_q_merge_nvls.log | cut -c1-200
JCK_MERGE_TAIL=0 timeout 150 $TR --nproc-per-node 8 --master-port 29543 $B > gpurun_out/r2_n8_q_nomerge.log 2>&1; grep -h '^{' gpurun_out/r2_n8_q_nomerge.log | cut -c1-170
