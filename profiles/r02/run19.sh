set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest19.log 2>&1; tail -3 gpurun_out/r2_pytest19.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-ops > gpurun_out/r2_bench19.log 2> gpurun_out/r2_bench19.err
cut -c1-300 gpurun_out/r2_bench19.log; tail -3 gpurun_out/r2_bench19.err
python profiles/one_step.py 2 > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bn_act|conv_tc_pair|conv_up_win|wgrad_tc|edge_down_direct|edge_up_scatter|wgrad_edge' --launch-skip 76 --launch-count 76 -o gpurun_out/prof_step_r2 -f python profiles/one_step.py 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
JCK_PDL=0 timeout 200 python bench.py --no-cpu-baseline --no-secondary --kernel-table --steps 5 --warmup 3 > gpurun_out/r2_ktable19.log 2> gpurun_out/r2_ktable19.err
