set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29543 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_n${N}_try.log 2>&1; tail -c 200 gpurun_out/r2_bench_n${N}_try.log
MFVAE_DP_BLOCKS=32 $TR --master-port 29544 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_n${N}_try_b32.log 2>&1
