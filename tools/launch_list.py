#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel for profiles/.

    python tools/launch_list.py gpurun_out/launches_r1p.csv profiles/r1p_launches.csv
Also prints every launch of the trace kernels in order (the timed steps are the equal-length ones).
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    trace = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            ns = float(r[ix["Metric Value"]])
        except ValueError:
            continue
        k = r[ix["Kernel Name"]]
        agg.setdefault(k, []).append(ns)
        if "trace_exchange" in k:
            trace.append(ns / 1e6)
    total = sum(sum(v) for v in agg.values())
    with open(sys.argv[2], "w") as f:
        f.write("kernel,launches,total_ms,avg_ms,share_of_process\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f'"{k}",{len(v)},{sum(v) / 1e6:.4f},{sum(v) / len(v) / 1e6:.4f},{sum(v) / total:.4f}\n')
        f.write("# trace-kernel launches in order (ms): " + " ".join(f"{t:.3f}" for t in trace) + "\n")
    print("trace launches (ms):", " ".join(f"{t:.3f}" for t in trace))


if __name__ == "__main__":
    main()
