set -x
mkdir -p gpurun_out
B="timeout 120 python bench.py --steps 40 --warmup 5 --quick"
$B > gpurun_out/r2_q17_front3.log 2>&1
JCK_BN_OCC=2 $B > gpurun_out/r2_q17_front2.log 2>&1
JCK_BN_FRONT_MIN_MB=8 $B > gpurun_out/r2_q17_min8.log 2>&1
JCK_BN_FRONT_MIN_MB=8 JCK_BN_OCC=2 $B > gpurun_out/r2_q17_min8_occ2.log 2>&1
JCK_BN_FRONT_MIN_MB=48 $B > gpurun_out/r2_q17_min48.log 2>&1
$B > gpurun_out/r2_q17_front3_b.log 2>&1
JCK_BN_OCC=2 $B > gpurun_out/r2_q17_front2_b.log 2>&1
for f in gpurun_out/r2_q17_*.log; do echo $f $(grep -o '"ms_per_step": [0-9.]*' $f); done
# per-step DRAM traffic WITH the kernel-to-kernel L2 hand-off (no cache flush between launches), front on / off
python profiles/one_step.py 2 > gpurun_out/plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches17_warm_front.csv python profiles/one_step.py 2 > gpurun_out/ncu_list_a.log 2>&1
JCK_BN_FRONT=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches17_warm_slab.csv python profiles/one_step.py 2 > gpurun_out/ncu_list_b.log 2>&1
tail -1 gpurun_out/ncu_list_a.log gpurun_out/ncu_list_b.log
