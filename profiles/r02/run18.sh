set -x
mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 200 --warmup 10 --quick"
for i in 1 2 3; do
$B > gpurun_out/r2_q18_o3m24_$i.log 2>&1
JCK_BN_OCC=2 $B > gpurun_out/r2_q18_o2m24_$i.log 2>&1
JCK_BN_FRONT_MIN_MB=8 $B > gpurun_out/r2_q18_o3m8_$i.log 2>&1
JCK_BN_FRONT_MIN_MB=8 JCK_BN_OCC=2 $B > gpurun_out/r2_q18_o2m8_$i.log 2>&1
done
JCK_BN_FRONT=0 $B > gpurun_out/r2_q18_slab_1.log 2>&1
for f in gpurun_out/r2_q18_*.log; do echo $f $(grep -o '"ms_per_step": [0-9.]*' $f); done
