set -x
for m in up:c2 up:w64 up:w32 up_groups:c2; do timeout 600 python -m tests.kernel_checks --match $m; done > gpurun_out/r2_kernels_win.log 2>&1
grep -v "^ok" gpurun_out/r2_kernels_win.log | tail -12
timeout 300 python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_win.log 2>&1
timeout 300 python tests/notes/conv_bench.py 1024 > gpurun_out/r2_convbench_win1024.log 2>&1
grep "c2.*up" gpurun_out/r2_convbench_win.log gpurun_out/r2_convbench_win1024.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench2.log 2> gpurun_out/r2_bench2.err
cut -c1-300 gpurun_out/r2_bench2.log; tail -3 gpurun_out/r2_bench2.err
