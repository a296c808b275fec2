#!/bin/bash
# Measured parity figures of the round as JSON lines (profiles/r2_parity.jsonl): every line is one tests/step_check.py run
# (CUDA path through the C ABI against the golden vectors minted from the reference and against the oracle).
OUT=${1:-gpurun_out/r2_parity.jsonl}
: > $OUT
run() { python tests/step_check.py "$@" 2>/dev/null | grep '^STEP_CHECK ' | sed 's/^STEP_CHECK //' >> $OUT; }
run default fp32 simt dropin
run default bf16 tcgen05 dropin
run default bf16 tcgen05 fast
run latent32 bf16 tcgen05 fast
run cfg2_b4096 bf16 tcgen05 fast
run cfg2_b4096 fp32 simt fast
run wide fp32 simt dropin
run wide bf16 tcgen05 dropin
run latent32 fp32 simt jointmse
run latent32 fp32 simt fast eps auto optenc
wc -l $OUT
