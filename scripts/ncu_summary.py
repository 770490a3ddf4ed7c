"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/ncu_summary.py rep  gpurun_out/prof_env.ncu-rep  profiles/r01_env_points.txt
    python scripts/ncu_summary.py list gpurun_out/launches_r01.csv   profiles/r01_launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__cycles_elapsed.avg.per_second",
]
STALLS = "smsp__average_warps_issue_stalled_"


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
        lines.append("kernel: %s" % d.get("Kernel Name", ("?", ""))[0])
        for k in KEYS:
            if k in d:
                lines.append("  %-70s %s %s" % (k, d[k][0], d[k][1]))
        st = [(h, d[h][0]) for h in hdr if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")]
        st = sorted(((h, float(v)) for h, v in st if v not in ("", "n/a")), key=lambda x: -x[1])[:8]
        lines.append("  top warp stall reasons (avg warps stalled per issue-active cycle):")
        for h, v in st:
            lines.append("    %-66s %.3f" % (h[len(STALLS):-len("_per_issue_active.ratio")], v))
        lines.append("")
    open(out, "w").write("# from %s (ncu --set full --clock-control none --import-source on)\n" % path + "\n".join(lines))


def lst(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# from %s: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n" % path)
        f.write("# %-100s %6s %14s %7s\n" % ("kernel", "n", "total_ns", "share"))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-102s %6d %14.0f %7.4f\n" % (k[:100], a[0], a[1], a[1] / tot))


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
