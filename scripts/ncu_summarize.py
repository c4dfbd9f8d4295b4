"""Turn .ncu-rep captures (gpurun_out/) into the tracked summaries under profiles/:

    python scripts/ncu_summarize.py <tag> gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]

writes profiles/<tag>_ncu_selected.json (one record per capture: duration, DRAM bytes, L1TEX
wavefronts, hit rates, occupancy, top stall reasons, ...) and, per capture,
profiles/<name>_ncu_details.txt (`ncu --page details`).  Runs here (no GPU needed)."""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALL = "smsp__average_warp_latency_issue_stalled_"  # ..._<reason>.pct / smsp__average_warps_issue_stalled_*


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return x


def scale(value, unit):
    m = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1,
         "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}
    return value * m[unit] if isinstance(value, float) and unit in m else value


def summarize(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        rec = {"kernel": row[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                rec[k] = scale(num(row[i]), units[i])
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                v = num(row[i])
                if isinstance(v, float):
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = v
        rec["stall_cycles_per_issue_top"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        out.append(rec)
    return out


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    sel = {}
    for rep in reps:
        name = os.path.basename(rep)[:-len(".ncu-rep")]
        sel[name] = summarize(rep)
        det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        with open(os.path.join(REPO, "profiles", name + "_ncu_details.txt"), "w") as f:
            f.write(det)
    path = os.path.join(REPO, "profiles", tag + "_ncu_selected.json")
    old = {}
    if os.path.exists(path):
        with open(path) as f:
            old = json.load(f)
    old.update(sel)
    with open(path, "w") as f:
        json.dump(old, f, indent=1)
    for name, recs in sel.items():
        for r in recs:
            print(name, r["kernel"][:60], "%.4g ms" % (1e3 * r.get("gpu__time_duration.sum", float("nan"))))


if __name__ == "__main__":
    main()
