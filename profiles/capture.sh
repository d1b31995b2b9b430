#!/bin/bash
# Round-1 ncu evidence (run under gpurun on ONE B200): launch list of two eager steps, then --set full captures of
# the tensor-core conv kernels, the BatchNorm streams and the image-edge / GEMM / weight-gradient kernels, all taken
# from the SECOND (warm) step, and the CUPTI kernel table.  Every command first runs once without ncu.
set -x
mkdir -p gpurun_out
python profiles/one_step.py 2 > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv \
    python profiles/one_step.py 2 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair -s 27 -c 8 -o gpurun_out/prof_conv_pair_r1 -f \
    python profiles/one_step.py 2 > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bn_act_fwd' -s 16 -c 6 -o gpurun_out/prof_bn_fwd_r1 -f \
    python profiles/one_step.py 2 > gpurun_out/ncu_bn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bn_act_bwd_reduce|bn_act_bwd_apply' -s 22 -c 6 -o gpurun_out/prof_bn_bwd_r1 -f \
    python profiles/one_step.py 2 > gpurun_out/ncu_bn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'edge_up_scatter|edge_down_direct|wgrad_edge_direct|gemm_tc|wgrad_tc' -s 12 -c 10 \
    -o gpurun_out/prof_edge_wgrad_r1 -f python profiles/one_step.py 2 > gpurun_out/ncu_edge.log 2>&1
# per-kernel device time (CUPTI): programmatic dependent launch off, or a kernel's duration includes the time it
# spends parked in griddepcontrol.wait behind its predecessor
JCK_PDL=0 python bench.py --no-cpu-baseline --kernel-table --steps 5 --warmup 3 > gpurun_out/ktable.log 2> gpurun_out/ktable.err
tail -n 3 gpurun_out/ncu_list.log gpurun_out/ncu_conv.log gpurun_out/ncu_edge.log

# Inception-v3 extractor (metrics.py path): launch list of two eager forwards, then --set full of the conv kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_incep_r1.csv \
    python tests/notes/incep_ncu.py > gpurun_out/ncu_incep_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 95 --launch-count 32 -f \
    -o gpurun_out/prof_incep_r1 python tests/notes/incep_ncu.py > gpurun_out/ncu_incep.log 2>&1
