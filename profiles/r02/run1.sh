set -x
python -m tests.kernel_checks --only-big > gpurun_out/r2_kernels_big.log 2>&1
python -m tests.notes.big_parity > gpurun_out/r2_big_parity.json 2> gpurun_out/r2_big_parity.err
python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/r2_bench0.log 2> gpurun_out/r2_bench0.err
JCK_PDL=0 python tests/notes/graph_timeline.py 512 > gpurun_out/r2_timeline0.log 2>&1
tail -3 gpurun_out/r2_kernels_big.log; tail -2 gpurun_out/r2_bench0.log | cut -c1-600
