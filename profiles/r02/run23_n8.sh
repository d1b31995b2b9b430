set -x
mkdir -p gpurun_out
export JCK_COMM_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING JCK_PDL=0 timeout 200 $TR --nproc-per-node 8 --master-port 29544 tests/notes/graph_timeline.py 512 > gpurun_out/r2_n8_timeline_full.log 2>&1
grep -v "NCCL INFO" gpurun_out/r2_n8_timeline_full.log > gpurun_out/r2_n8_timeline.log
grep "NCCL INFO" gpurun_out/r2_n8_timeline_full.log | grep -i "nvls\|algo\|proto\|AllReduce\|channels\|Connected" | grep "^.*\[0\]" | head -60 > gpurun_out/r2_n8_nccl_info.log
rm gpurun_out/r2_n8_timeline_full.log
tail -3 gpurun_out/r2_n8_timeline.log; wc -l gpurun_out/r2_n8_nccl_info.log
NCCL_PROTO=Simple timeout 150 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 40 --warmup 5 --quick > gpurun_out/r2_n8_quick_simple.log 2>&1; grep -h '^{' gpurun_out/r2_n8_quick_simple.log | cut -c1-200
