set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for b in 16 32; do
  MFVAE_DP_BLOCKS=$b $TR --master-port 29544 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_dpx_n${N}_b$b.log 2>&1
done
