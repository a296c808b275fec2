#!/usr/bin/env python
"""Text timeline of one train step out of tools/dp_trace.py's JSON: every kernel with its start offset, duration and stream,
plus per-stream busy time and the idle gaps on the union of all streams.
    python tools/trace_report.py gpurun_out/dp_trace_n1.json [step_index]"""
import json
import sys


def main(path, which=2):
    ev = [e for e in json.load(open(path)) if e["dur_us"] > 0]
    # steps start at the zero-grad memset / stage kernel: split on the first kernel named *stage_kernel* or onehot
    starts = [i for i, e in enumerate(ev) if "stage_kernel" in e["name"]]
    if len(starts) < which + 2:
        which = max(0, len(starts) - 2)
    lo, hi = starts[which], starts[which + 1]
    # include memsets / kernels launched just before the stage kernel of this step (zero_grads, act fold on aux)
    while lo > 0 and ev[lo - 1]["ts_us"] > ev[starts[which]]["ts_us"] - 30 and "adam" not in ev[lo - 1]["name"]:
        lo -= 1
    step = ev[lo:hi]
    t0 = step[0]["ts_us"]
    streams = sorted({e["stream"] for e in step})
    print(f"step {which}: {len(step)} device activities, span {step[-1]['ts_us'] + step[-1]['dur_us'] - t0:.1f} us, next step starts at {ev[hi]['ts_us'] - t0:.1f} us")
    for e in step:
        col = streams.index(e["stream"])
        print(f"{e['ts_us'] - t0:8.1f} {e['dur_us']:7.1f}  s{col}  {e['name']}")
    for sid in streams:
        busy = sum(e["dur_us"] for e in step if e["stream"] == sid)
        print(f"stream s{streams.index(sid)} ({sid}): busy {busy:.1f} us")
    # union coverage
    iv = sorted((e["ts_us"], e["ts_us"] + e["dur_us"]) for e in step)
    cov, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    gaps = []
    for s, e in iv[1:]:
        if s > cur_e:
            gaps.append((cur_e - t0, s - cur_e)); cov += cur_e - cur_s; cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    cov += cur_e - cur_s
    print(f"union busy {cov:.1f} us; idle gaps > 2 us: " + ", ".join(f"{g:.1f}@{t:.0f}" for t, g in gaps if g > 2))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
