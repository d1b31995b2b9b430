set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest14.log 2>&1; tail -3 gpurun_out/r2_pytest14.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench14.log 2> gpurun_out/r2_bench14.err
cut -c1-400 gpurun_out/r2_bench14.log; tail -3 gpurun_out/r2_bench14.err
JCK_PDL=0 timeout 300 python bench.py --no-cpu-baseline --no-secondary --kernel-table --steps 5 --warmup 3 > gpurun_out/r2_ktable14.log 2> gpurun_out/r2_ktable14.err
python profiles/one_step.py 2 > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches14.csv \
    python profiles/one_step.py 2 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
