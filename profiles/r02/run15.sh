set -x
mkdir -p gpurun_out
timeout 600 python -m tests.kernel_checks --match bn > gpurun_out/r2_kernels_bn15.log 2>&1; grep -v "^ok" gpurun_out/r2_kernels_bn15.log | tail -8; grep "^ok ('bn" gpurun_out/r2_kernels_bn15.log
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline"
JCK_BN_FRONT=0 JCK_PACK_OVERLAP=0 $B > gpurun_out/r2_bench15_base.log 2> gpurun_out/r2_bench15_base.err
JCK_BN_FRONT=0 $B > gpurun_out/r2_bench15_pack.log 2> gpurun_out/r2_bench15_pack.err
$B > gpurun_out/r2_bench15_front3.log 2> gpurun_out/r2_bench15_front3.err
JCK_BN_OCC=2 $B > gpurun_out/r2_bench15_front2.log 2> gpurun_out/r2_bench15_front2.err
for f in base pack front3 front2; do echo $f; cut -c1-200 gpurun_out/r2_bench15_$f.log | tail -1; tail -2 gpurun_out/r2_bench15_$f.err; done
JCK_BN_FRONT=0 timeout 300 python tests/notes/bn_bench.py 512 > gpurun_out/r2_bnbench15_slab.log 2>&1
timeout 300 python tests/notes/bn_bench.py 512 > gpurun_out/r2_bnbench15_front3.log 2>&1
JCK_BN_OCC=2 timeout 300 python tests/notes/bn_bench.py 512 > gpurun_out/r2_bnbench15_front2.log 2>&1
paste gpurun_out/r2_bnbench15_slab.log gpurun_out/r2_bnbench15_front3.log gpurun_out/r2_bnbench15_front2.log | awk '{print $1,$2,$3,$4,$5,$8,$9,$21,$22,$34,$35}' | head -40
timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_big.py tests/test_input_pipeline.py -x -q -m gpu > gpurun_out/r2_pytest15.log 2>&1; tail -3 gpurun_out/r2_pytest15.log
