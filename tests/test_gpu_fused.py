"""-m gpu: the fused encoder kernels against the layer-by-layer kernels of the same library (same math, same rounding
points), at the benchmark shapes (cfg2: latent 32; reference dims: latent 64), ragged batch sizes (not multiples of the
128-row tile; fewer rows than one tile) and with an explicit agent-index column."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fused_check.py"), *map(str, args)],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("FUSED_CHECK ")]
    assert r.returncode == 0 and line, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads(line[-1][len("FUSED_CHECK "):])
    print(json.dumps(out))
    return out


@pytest.mark.parametrize("latent,batch,extra", [(32, 4096, ""), (64, 1024, ""), (32, 300, ""), (64, 77, ""),
                                                (32, 512, "explicit_idx")])
def test_fused_matches_layerwise(latent, batch, extra):
    o = run(latent, batch, *([extra] if extra else []))
    assert o["finite"]
    # forward: identical operations in identical order -> agreement at fp32 round-off
    assert max(o["fwd_mu"], o["fwd_lv"]) < 1e-5, o
    assert max(o["fwd_rs"], o["fwd_rr"]) < 1e-4, o
    assert max(o["loss_rel"]) < 1e-4, o
    # gradients: split-K / atomic accumulation orders differ between the two schedules
    assert o["grad_rel_median"] < 1e-3 and o["grad_rel_max"] < 2e-2, (o["grad_rel_median"], o["grad_rel_worst"])
    # loss fused into the output-layer epilogue vs the separate loss kernel: same fp32 values enter the same formula
    assert o["lossfuse_loss_rel"] < 1e-5 and o["lossfuse_grad_rel_max"] < 1e-3, {k: v for k, v in o.items() if k.startswith("lossfuse")}
    # default train step: recon_s lives only as bf16 inside the D(recon_s) buffer (one extra bf16 rounding point)
    assert o["r16_loss_rel"] < 1e-4 and o["r16_grad_rel_median"] < 2e-3 and o["r16_grad_rel_max"] < 1e-1, \
        {k: v for k, v in o.items() if k.startswith("r16")}


def run_fold(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fold_check.py"), *map(str, args)],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("FOLD_CHECK ")]
    assert r.returncode == 0 and line, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads(line[-1][len("FOLD_CHECK "):])
    print(json.dumps(out))
    return out


@pytest.mark.parametrize("batch,extra", [(256, ""), (77, ""), (256, "perm_idx")])
def test_folded_column_blocks_match_dense_fp32(batch, extra):
    """csrc/fold.cu in fp32: the regrouped sums are the same mathematics -> agreement at fp32 round-off (a ReLU unit within
    round-off of zero may flip: bounded through the whole-gradient metric and the median over tensors)."""
    o = run_fold("fp32", batch, *([extra] if extra else []))
    assert o["finite"]
    assert max(o["fwd_mu"], o["fwd_lv"], o["fwd_rs"], o["fwd_rr"]) < 1e-5, o
    assert max(o["loss_rel"]) < 1e-5, o
    assert o["grad_rel_median"] < 1e-5 and o["whole_grad_rel"] < 5e-4, o


def test_folded_column_blocks_match_dense_bf16():
    """bf16: the fold moves rounding points (T and the folded bias are rounded once instead of their factors), so the two
    evaluations agree at the bf16 level -- like any two bf16 evaluations of this network (ReLU flips)."""
    o = run_fold("bf16", 4096)
    assert o["finite"]
    assert max(o["fwd_mu"], o["fwd_lv"], o["fwd_rs"]) < 1e-2 and max(o["loss_rel"]) < 2e-3, o
    assert o["whole_grad_cos"] > 0.999, o
