set -x
timeout 900 python -m tests.kernel_checks --match wgrad > gpurun_out/r2_kernels_wgrad.log 2>&1
grep -v "^ok" gpurun_out/r2_kernels_wgrad.log | tail -5
timeout 300 python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_wg.log 2>&1
timeout 300 python tests/notes/conv_bench.py 1024 > gpurun_out/r2_convbench_wg1024.log 2>&1
grep "wgrad" gpurun_out/r2_convbench_wg.log gpurun_out/r2_convbench_wg1024.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err
cut -c1-300 gpurun_out/r2_bench5.log; tail -3 gpurun_out/r2_bench5.err
