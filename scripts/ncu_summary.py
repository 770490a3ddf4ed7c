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
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
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


def _num(v, unit=""):
    """ncu prints byte counts with a unit column (byte / Kbyte / Mbyte / Gbyte)"""
    x = float(str(v).replace(",", ""))
    return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3,
                "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)


def js(path, part, units, summary_file, out_json):
    """merge one kernel's counters into profiles/ncu_traffic.json (what bench.py's roofline objects read):
    python scripts/ncu_summary.py json <rep> <part> <units per launch> <committed summary file> <json>"""
    import json
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, unit_row, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, v, u in zip(hdr, vals, unit_row)}

    def get(k):
        return _num(*d[k]) if k in d and d[k][0] not in ("", "n/a") else None
    e = {"file": summary_file, "units": int(units), "kernel": d.get("Kernel Name", ("?", ""))[0][:80],
         "dram_bytes": (get("dram__bytes_read.sum") or 0) + (get("dram__bytes_write.sum") or 0),
         "l1_bytes": get("l1tex__t_bytes.sum"), "lts_bytes": get("lts__t_bytes.sum"),
         "fp64_inst": get("sm__inst_executed_pipe_fp64.sum"), "inst": get("smsp__inst_executed.sum"),
         "fp64_pipe_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
         "l1tex_throughput_pct": get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
         "lts_throughput_pct": get("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
         "dram_throughput_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "lanes_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
         "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
         "registers": get("launch__registers_per_thread"), "kernel_ms_under_ncu": get("gpu__time_duration.sum")}
    e = {k: v for k, v in e.items() if v is not None}
    try:
        allj = json.load(open(out_json))
    except Exception:
        allj = {}
    allj[part] = e
    json.dump(allj, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "json":
        js(*sys.argv[2:7])
    else:
        {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
