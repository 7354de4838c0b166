"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
    python scripts/summarize_ncu.py kernel gpurun_out/prof_trace_r1.ncu-rep profiles/r1_trace_bin.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr, data = None, []
    for r in rows:
        if r[0] == "ID":
            hdr = r
        elif hdr and r[0].isdigit():
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")[:80]
        v = float(d["Metric Value"])
        ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1e-6) * v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`--metrics gpu__time_duration.sum --clock-control none`: "
                "cold-cache, serialised launches; compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {ms:.4f} | {100 * ms / tot:.2f}% | {ms / c:.4f} |\n")
    print(open(dst).read())


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for vals in rows[2:]:
            rec = dict(zip(hdr, zip(units, vals)))
            f.write(f"## `{rec['Kernel Name'][1][:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEEP:
                if k in rec:
                    f.write(f"| {k} | {rec[k][1]} | {rec[k][0]} |\n")
            f.write("\nwarp stall reasons (warps per issue-active cycle):\n\n| reason | value |\n|---|---:|\n")
            st = [(k[len(STALL):].replace("_per_issue_active.ratio", ""), float(v[1]))
                  for k, v in rec.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")]
            for k, v in sorted(st, key=lambda kv: -kv[1])[:8]:
                f.write(f"| {k} | {v:.3f} |\n")
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
