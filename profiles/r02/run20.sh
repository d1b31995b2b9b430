set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_incep.py "tests/test_gpu_step.py::test_metrics_end_to_end" -x -q -m gpu -s > gpurun_out/r2_pytest20.log 2>&1; tail -4 gpurun_out/r2_pytest20.log; grep "split precision" gpurun_out/r2_pytest20.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-ops > gpurun_out/r2_bench20.log 2> gpurun_out/r2_bench20.err
cut -c1-300 gpurun_out/r2_bench20.log; tail -3 gpurun_out/r2_bench20.err
JCK_PDL=0 timeout 200 python bench.py --no-cpu-baseline --no-secondary --kernel-table --steps 5 --warmup 3 > gpurun_out/r2_ktable20.log 2> gpurun_out/r2_ktable20.err
python profiles/one_step.py 2 > gpurun_out/plain.log 2>&1 || exit 1
# --set full of the second (warm) step's D-side kernels: forward of the 1536-image pass, backward of the A+B slice
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bn_act|conv_tc_pair|conv_up_win|wgrad_tc|edge_down_direct' --launch-skip 90 --launch-count 36 -o gpurun_out/prof_step_r2 -f python profiles/one_step.py 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_step_r2.ncu-rep --page raw --csv > gpurun_out/prof_step_r2_raw.csv 2>/dev/null
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/prof_step_r2.ncu-rep); if [ "$sz" -gt 45000000 ]; then rm gpurun_out/prof_step_r2.ncu-rep; fi
du -sh gpurun_out
