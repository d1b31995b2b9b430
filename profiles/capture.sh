#!/bin/bash
# Round-1 ncu evidence (run under gpurun on ONE B200): launch list of two eager steps, then --set full captures of
# the tensor-core conv kernels and the BatchNorm streams taken from the SECOND (warm) step.
set -x
mkdir -p gpurun_out
python profiles/one_step.py 2 > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv \
    python profiles/one_step.py 2 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair -s 27 -c 8 -o gpurun_out/prof_conv_pair_r1 -f \
    python profiles/one_step.py 2 > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bn_act_bwd_apply|bn_act_fwd|wgrad_tc' -s 34 -c 8 -o gpurun_out/prof_bn_wgrad_r1 -f \
    python profiles/one_step.py 2 > gpurun_out/ncu_bn.log 2>&1
tail -3 gpurun_out/ncu_list.log gpurun_out/ncu_conv.log gpurun_out/ncu_bn.log
