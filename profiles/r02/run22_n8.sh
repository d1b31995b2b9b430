set -x
mkdir -p gpurun_out
export JCK_COMM_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 40 --warmup 5 --quick > gpurun_out/r2_n8_quick.log 2>&1; grep -h '^{' gpurun_out/r2_n8_quick.log | cut -c1-200
timeout 300 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_n8_bench.log 2> gpurun_out/r2_n8_bench.err; cut -c1-300 gpurun_out/r2_n8_bench.log; tail -3 gpurun_out/r2_n8_bench.err
timeout 150 $TR --nproc-per-node 4 --master-port 29545 bench.py --gpus 4 --steps 40 --warmup 5 --quick > gpurun_out/r2_n4_quick.log 2>&1; grep -h '^{' gpurun_out/r2_n4_quick.log | cut -c1-200
timeout 120 python bench.py --gpus 1 --steps 40 --warmup 5 --quick > gpurun_out/r2_n8_quick_n1.log 2>&1; grep -h '^{' gpurun_out/r2_n8_quick_n1.log | cut -c1-200
