#!/bin/bash
# single-GPU measurements of the round (run under gpurun): every bench workload, the configs[4] microbenchmarks, ncu evidence
set -x
OUT=gpurun_out
python bench.py --steps 50 --warmup 5 > $OUT/r2_bench_cfg2_n1.json 2>$OUT/r2_bench_cfg2_n1.err
for w in ref_dims wide cfg1 cfg4; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > $OUT/r2_bench_${w}_n1.json 2>$OUT/r2_bench_${w}_n1.err
done
python tools/microbench.py --out $OUT/r2_cfg5_microbench.jsonl > $OUT/microbench.log 2>&1
bash tools/profile_r2.sh > $OUT/profile_r2.log 2>&1
tail -3 $OUT/profile_r2.log
