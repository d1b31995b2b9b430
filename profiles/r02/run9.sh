set -x
for m in bnbwd_up up:c2 up_groups:c2; do timeout 900 python -m tests.kernel_checks --match $m; done > gpurun_out/r2_kernels_winbn.log 2>&1
grep -v "^ok" gpurun_out/r2_kernels_winbn.log | tail -12
timeout 300 python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_winbn.log 2>&1
timeout 300 python tests/notes/conv_bench.py 1024 > gpurun_out/r2_convbench_winbn1024.log 2>&1
grep "c2.*up" gpurun_out/r2_convbench_winbn.log gpurun_out/r2_convbench_winbn1024.log
timeout 600 python -m pytest tests/test_gpu_big.py tests/test_gpu_step.py -x -q > gpurun_out/r2_pytest_winbn.log 2>&1; tail -3 gpurun_out/r2_pytest_winbn.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --profile-ops > gpurun_out/r2_bench4.log 2> gpurun_out/r2_bench4.err
cut -c1-300 gpurun_out/r2_bench4.log; tail -3 gpurun_out/r2_bench4.err
JCK_PDL=0 python tests/notes/graph_timeline.py 512 > gpurun_out/r2_timeline2.log 2>&1
