set -x
python -m tests.kernel_checks --isolate --only-tc 2>&1 | grep -v "^ok" > gpurun_out/r2_kernels_tc.log
tail -12 gpurun_out/r2_kernels_tc.log
python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_win.log 2>&1
JCK_UP_WIN=0 python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_nowin.log 2>&1
python tests/notes/conv_bench.py 1024 > gpurun_out/r2_convbench_win1024.log 2>&1
grep "c2.*up" gpurun_out/r2_convbench_win.log gpurun_out/r2_convbench_nowin.log gpurun_out/r2_convbench_win1024.log
python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench2.log 2> gpurun_out/r2_bench2.err
cut -c1-300 gpurun_out/r2_bench2.log; tail -3 gpurun_out/r2_bench2.err
