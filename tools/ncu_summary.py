#!/usr/bin/env python
"""Summarise an ncu report for profiles/: key metrics of the trace kernel + executed SASS mix per ray.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name> --rays 999999870
writes <name>_metrics.csv (selected raw metrics) and <name>_sass_mix.csv (warp instructions per 32 rays by opcode).
"""
import argparse
import collections
import csv
import io
import subprocess

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "lts__t_requests_srcunit_tex_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_prefix")
    ap.add_argument("--rays", type=float, required=True, help="rays traced by the profiled launch")
    a = ap.parse_args()
    rows = ncu_csv(a.report, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(a.out_prefix + "_metrics.csv", "w") as f:
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or ("issue_stalled" in h and "per_issue_active" in h) or h == "Kernel Name":
                f.write(f'{h},{u},"{v}"\n')
    rows = ncu_csv(a.report, "source")
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, thr, n_all = collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]].strip()
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        n, t = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
        tot[op] += n; thr[op] += t; n_all += n
    wr = a.rays / 32.0
    with open(a.out_prefix + "_sass_mix.csv", "w") as f:
        f.write("opcode,warp_instructions,percent,per_32_rays,avg_active_threads\n")
        f.write(f"TOTAL,{n_all},100.0,{n_all / wr:.1f},{sum(thr.values()) / max(n_all, 1):.1f}\n")
        for op, n in tot.most_common():
            f.write(f"{op},{n},{100 * n / n_all:.2f},{n / wr:.2f},{thr[op] / max(n, 1):.1f}\n")
    print(f"{n_all / wr:.1f} warp instructions per 32 rays")


if __name__ == "__main__":
    main()
