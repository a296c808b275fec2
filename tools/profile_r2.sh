#!/bin/bash
# ncu evidence for one round: (1) launch list of one train step (durations + DRAM bytes, serialised, cold cache),
# (2) one --set full capture each of the fused encoder kernel and of the largest tcgen05 GEMM.  Run under gpurun, 1 GPU.
set -x
OUT=gpurun_out
KRE='regex:^(act_fold|adam|cast_bf16|colsum|emb_grad|enc_bias|enc_fwd|gemm_tc|loss_total|onehot|recon_loss|reparam_kl|stage)'
python tools/one_step.py --steps 6 > $OUT/one_step.log 2>&1 || exit 1
ncu --kernel-name "$KRE" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --launch-skip 174 -c 122 --csv --log-file $OUT/r2_launches.csv python tools/one_step.py --steps 6 > $OUT/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name 'regex:^enc_fwd_kernel' --launch-skip 3 -c 1 -f -o $OUT/enc_fwd \
    python tools/one_step.py --steps 6 > $OUT/ncu_full1.log 2>&1
ncu -i $OUT/enc_fwd.ncu-rep --page raw --csv > $OUT/r2_ncu_full_enc_fwd_raw.csv 2>/dev/null
ncu -i $OUT/enc_fwd.ncu-rep --page details --csv > $OUT/r2_ncu_full_enc_fwd_details.csv 2>/dev/null
rm -f $OUT/enc_fwd.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base demangled --kernel-name 'regex:gemm_tc_kernel<\(int\)256, \(bool\)1, \(bool\)1' --launch-skip 6 -c 2 -f -o $OUT/gemm_wg \
    python tools/one_step.py --steps 6 > $OUT/ncu_full2.log 2>&1
ncu -i $OUT/gemm_wg.ncu-rep --page raw --csv > $OUT/r2_ncu_full_gemm_wgrad_raw.csv 2>/dev/null
ncu -i $OUT/gemm_wg.ncu-rep --page details --csv > $OUT/r2_ncu_full_gemm_wgrad_details.csv 2>/dev/null
rm -f $OUT/gemm_wg.ncu-rep
ls -la $OUT/r2_ncu_full* $OUT/r2_launches.csv
