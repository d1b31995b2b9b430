set -x
mkdir -p gpurun_out
export JCK_COMM_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 bench.py --gpus 2 --steps 40 --warmup 5 --quick > gpurun_out/r2_n2_quick.log 2>&1; grep -h '^{' gpurun_out/r2_n2_quick.log | cut -c1-200
timeout 120 python bench.py --gpus 1 --steps 40 --warmup 5 --quick > gpurun_out/r2_n2_quick_n1.log 2>&1; grep -h '^{' gpurun_out/r2_n2_quick_n1.log | cut -c1-200
timeout 400 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_n2_bench.log 2> gpurun_out/r2_n2_bench.err; cut -c1-300 gpurun_out/r2_n2_bench.log; tail -3 gpurun_out/r2_n2_bench.err
JCK_PDL=0 timeout 200 $TR --master-port 29544 tests/notes/graph_timeline.py 512 > gpurun_out/r2_n2_timeline.log 2>&1; tail -2 gpurun_out/r2_n2_timeline.log
JCK_PDL=0 timeout 200 python tests/notes/graph_timeline.py 512 > gpurun_out/r2_n1_timeline.log 2>&1; tail -2 gpurun_out/r2_n1_timeline.log
