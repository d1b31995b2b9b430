set -x
timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_cgan.py tests/test_gpu_big.py -x -q > gpurun_out/r2_pytest_arena.log 2>&1; tail -3 gpurun_out/r2_pytest_arena.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench6.log 2> gpurun_out/r2_bench6.err
cut -c1-300 gpurun_out/r2_bench6.log; tail -3 gpurun_out/r2_bench6.err
JCK_PDL=0 python tests/notes/graph_timeline.py 512 > gpurun_out/r2_timeline3.log 2>&1
grep -c "at::native" gpurun_out/r2_timeline3.log
