#!/bin/bash
# multi-GPU measurements of the round: bash tools/measure_r2_multi.sh N   (run under gpurun --gpus N)
set -x
N=$1
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_dp.py -m gpu -q 2>&1 | tail -5 > $OUT/r2_pytest_dp2.log; cat $OUT/r2_pytest_dp2.log
  $TR --master-port 29551 tools/dp_trace.py > $OUT/r2_trace2.log 2>&1
  $TR --master-port 29552 bench.py --gpus 2 --workload wide --batch 8192 --steps 10 --warmup 3 > $OUT/r2_bench_wide_n2.json 2>$OUT/r2_bench_wide_n2.err
fi
$TR --master-port 29553 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/r2_bench_cfg2_n$N.json 2>$OUT/r2_bench_cfg2_n$N.err
MFVAE_DP_COMM=nccl $TR --master-port 29554 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/r2_bench_cfg2_n${N}_nccl.json 2>$OUT/r2_bench_cfg2_n${N}_nccl.err
if [ "$N" = "4" ]; then
  $TR --master-port 29555 bench.py --gpus 4 --workload wide --steps 10 --warmup 3 > $OUT/r2_bench_wide_n4.json 2>$OUT/r2_bench_wide_n4.err
fi
if [ "$N" = "8" ]; then
  $TR --master-port 29556 bench.py --gpus 8 --workload cfg4 --steps 20 --warmup 3 > $OUT/r2_bench_cfg4_n8.json 2>$OUT/r2_bench_cfg4_n8.err
fi
tail -c 300 $OUT/r2_bench_cfg2_n$N.json
