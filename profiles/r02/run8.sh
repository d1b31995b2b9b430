set -x
export JCK_BN256=1
for m in down:c3 down:c4 up:c4 down_groups up_groups:c4 bnbwd; do timeout 900 python -m tests.kernel_checks --match $m; done > gpurun_out/r2_kernels_bn256.log 2>&1
grep -v "^ok" gpurun_out/r2_kernels_bn256.log | tail -12
timeout 300 python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_bn256.log 2>&1
timeout 300 python tests/notes/conv_bench.py 1536 > gpurun_out/r2_convbench_bn256_1536.log 2>&1
unset JCK_BN256
timeout 300 python tests/notes/conv_bench.py 512 > gpurun_out/r2_convbench_base.log 2>&1
timeout 300 python tests/notes/conv_bench.py 1536 > gpurun_out/r2_convbench_base_1536.log 2>&1
paste -d'\n' gpurun_out/r2_convbench_base.log gpurun_out/r2_convbench_bn256.log | grep -E "c3|c4"
paste -d'\n' gpurun_out/r2_convbench_base_1536.log gpurun_out/r2_convbench_bn256_1536.log | grep -E "c3|c4"
JCK_BN256=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2_bench3.log 2> gpurun_out/r2_bench3.err
cut -c1-300 gpurun_out/r2_bench3.log; tail -3 gpurun_out/r2_bench3.err
