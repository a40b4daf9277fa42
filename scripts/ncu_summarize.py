"""Summarise an ncu report (read here, no GPU needed) into profiles/: key metrics JSON,
opcode mix of the executed SASS, and (for a launch-list CSV) per-kernel time shares.

    python scripts/ncu_summarize.py rep gpurun_out/prof.ncu-rep profiles/r01_k1
    python scripts/ncu_summarize.py launches gpurun_out/launches.csv profiles/r01_launches.txt
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.sum",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.sum",
    "smsp__inst_executed_pipe_fp64.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(rep, prefix):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    result = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name",) or h in KEYS or h.startswith("smsp__average_warps_issue_stalled"):
                d[h] = v if h == "Kernel Name" else f"{v} {u}".strip()
        result.append(d)
    src = ncu_csv(rep, "source")
    mix = collections.Counter()
    total = 0
    if len(src) > 2:
        h2 = src[1]
        ia, ie = h2.index("Source"), h2.index("Instructions Executed")
        for r in src[2:]:
            if len(r) <= ie:
                continue
            try:
                n = int(r[ie])
            except ValueError:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia])
            mix[m.group(2) if m else "?"] += n
            total += n
    first = result[0] if result else {}

    def num(key):
        try:
            return float(first.get(key, "0").split()[0].replace(",", ""))
        except Exception:
            return None

    def to_bytes(key):
        s = first.get(key, "")
        parts = s.split()
        if not parts:
            return None
        v = float(parts[0].replace(",", ""))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[1] if len(parts) > 1 else "byte", 1)
        return v * mult

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    summary = {
        "report": rep, "kernel": first.get("Kernel Name"),
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_launch": (rd or 0) + (wr or 0),
        "duration_us_under_ncu": num("gpu__time_duration.sum"),
        "warp_instructions_executed": num("smsp__inst_executed.sum"),
        "metrics": first,
        "sass_opcode_mix_top": [[k, v, round(100.0 * v / max(1, total), 2)] for k, v in mix.most_common(30)],
    }
    json.dump(summary, open(prefix + "_ncu_summary.json", "w"), indent=1)
    print("wrote", prefix + "_ncu_summary.json")


def summarize_launches(csv_path, out_path):
    rows = list(csv.reader(open(csv_path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) > mv:
            name = r[kn].split("(")[0].replace("void ", "")
            v = float(r[mv].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
            agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out_path, "w") as f:
        f.write(f"# per-launch gpu__time_duration from {csv_path} (ncu: cold-cache, serialised; compare SHARES)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:48s} n={len(v):4d} mean={sum(v)/len(v):9.2f} us total={sum(v):10.1f} us share={100*sum(v)/tot:5.1f}%\n")
    print(open(out_path).read())


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        summarize_rep(sys.argv[2], sys.argv[3])
    else:
        summarize_launches(sys.argv[2], sys.argv[3])
