set -x
python -m pytest tests/test_gpu_big.py tests/test_gpu_step.py -x -q > gpurun_out/r2_pytest_big.log 2>&1
tail -5 gpurun_out/r2_pytest_big.log
( time python bench.py --steps 20 --warmup 5 --profile-ops ) > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err
JCK_GP_STREAM=0 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench1_nogp.log 2> gpurun_out/r2_bench1_nogp.err
JCK_PDL=0 python tests/notes/graph_timeline.py 512 > gpurun_out/r2_timeline1.log 2>&1
cut -c1-300 gpurun_out/r2_bench1.log; cut -c1-300 gpurun_out/r2_bench1_nogp.log; tail -3 gpurun_out/r2_bench1.err
