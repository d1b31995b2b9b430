set -x
export JCK_COMM_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29543 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_n2_bench.log 2> gpurun_out/r2_n2_bench.err; cut -c1-300 gpurun_out/r2_n2_bench.log; tail -3 gpurun_out/r2_n2_bench.err
python bench.py --gpus 1 --steps 30 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_n2_bench_n1.log 2>&1; cut -c1-200 gpurun_out/r2_n2_bench_n1.log
timeout 300 $TR --master-port 29544 tests/notes/graph_timeline.py 512 > gpurun_out/r2_n2_timeline_pdl.log 2>&1
