set -x
for m in bnbwd_up:c2 bnbwd_up:w32 up:c2; do timeout 900 python -m tests.kernel_checks --match $m; done > gpurun_out/r2_kernels_winbn.log 2>&1
grep -v "^ok" gpurun_out/r2_kernels_winbn.log | tail -12
timeout 300 python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_winbn.log 2>&1
timeout 300 python tests/notes/conv_bench.py 1024 > gpurun_out/r2_convbench_winbn1024.log 2>&1
grep "c2.*up" gpurun_out/r2_convbench_winbn.log gpurun_out/r2_convbench_winbn1024.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --profile-ops > gpurun_out/r2_bench4.log 2> gpurun_out/r2_bench4.err
cut -c1-300 gpurun_out/r2_bench4.log; tail -3 gpurun_out/r2_bench4.err
