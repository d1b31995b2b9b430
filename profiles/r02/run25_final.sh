set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest25.log 2>&1; tail -3 gpurun_out/r2_pytest25.log
timeout 100 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r2_smoke25.log 2>&1; tail -2 gpurun_out/r2_smoke25.log
timeout 500 python bench.py --steps 20 --warmup 5 --profile-ops > gpurun_out/r2_bench25.log 2> gpurun_out/r2_bench25.err
cut -c1-260 gpurun_out/r2_bench25.log; grep -c "^#" gpurun_out/r2_bench25.err
