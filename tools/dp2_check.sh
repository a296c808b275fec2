set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_dp.py -m gpu -q 2>&1 | tail -3; fi
$TR --master-port 29545 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_cfg2_n${N}_split.json 2>gpurun_out/r2_bench_cfg2_n${N}_split.err
if [ "$N" = "2" ]; then MFVAE_DP_BLOCKS=64 $TR --master-port 29546 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_cfg2_n${N}_split_b64.json 2>/dev/null; fi
