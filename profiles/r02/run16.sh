set -x
mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 40 --warmup 5 --quick"
JCK_BN_FRONT=0 $B > gpurun_out/r2_q16_base.log 2>&1
$B > gpurun_out/r2_q16_front3.log 2>&1
JCK_BN_OCC=2 $B > gpurun_out/r2_q16_front2.log 2>&1
JCK_BN_FRONT=0 $B > gpurun_out/r2_q16_base_b.log 2>&1
$B > gpurun_out/r2_q16_front3_b.log 2>&1
grep -h '^{' gpurun_out/r2_q16_*.log | cut -c1-160
for f in gpurun_out/r2_q16_*.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f; done
