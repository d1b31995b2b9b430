set -x
export JCK_COMM_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 -m tests.comm_check > gpurun_out/r2_n2_comm.log 2>&1; tail -2 gpurun_out/r2_n2_comm.log
timeout 300 $TR --master-port 29542 -m tests.dp_check > gpurun_out/r2_n2_dp.log 2>&1; tail -3 gpurun_out/r2_n2_dp.log
timeout 600 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_n2_bench.log 2> gpurun_out/r2_n2_bench.err; cut -c1-400 gpurun_out/r2_n2_bench.log; tail -3 gpurun_out/r2_n2_bench.err
JCK_PDL=0 timeout 300 $TR --master-port 29544 tests/notes/graph_timeline.py 512 > gpurun_out/r2_n2_timeline.log 2>&1; tail -3 gpurun_out/r2_n2_timeline.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_n2_bench_n1.log 2>&1; cut -c1-200 gpurun_out/r2_n2_bench_n1.log
