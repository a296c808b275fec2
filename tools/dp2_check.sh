set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${2:-all}" = "all" ]; then
$TR --master-port 29541 tests/dp_gpu_check.py fp32 > gpurun_out/dp${N}_fp32.log 2>&1; grep -E "DP_CHECK|DP_SYNC|Error|error" gpurun_out/dp${N}_fp32.log | tail -5
$TR --master-port 29542 tests/dp_gpu_check.py bf16 > gpurun_out/dp${N}_bf16.log 2>&1; grep -E "DP_CHECK|DP_SYNC|Error|error" gpurun_out/dp${N}_bf16.log | tail -5
fi
$TR --master-port 29543 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_n${N}_native.log 2>&1; tail -c 200 gpurun_out/r2_bench_n${N}_native.log
