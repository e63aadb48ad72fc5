#!/usr/bin/env python3
"""Turn ncu output into the text summaries kept under profiles/.

  ncu_summary.py launches  launches.csv                 -> per-kernel totals / shares of an `ncu --metrics gpu__time_duration.sum --csv` log
  ncu_summary.py full      report.ncu-rep [frames]      -> the metrics DESIGN.md quotes, per kernel of an `ncu --set full` report
                                                           (frames = frames the captured launch processed: per-frame figures)
"""
import csv
import io
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    ix = {k: i for i, k in enumerate(h)}
    agg = {}
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = short(r[ix["Kernel Name"]])
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        a = agg.setdefault(k, {"n": 0, "ms": 0.0, "shape": set()})
        a["n"] += 1
        a["ms"] += v
        a["shape"].add((r[ix["Grid Size"]], r[ix["Block Size"]]))
    own = re.compile(r"^(icw::)?(mt_|advance_streams|hb_|chain_|scan_|ns_render|crc_|sincos_leaf|phase_leaf|modal_)")
    ours = {k: a for k, a in agg.items() if own.match(k)}
    tot = sum(a["ms"] for a in ours.values())
    lib = {k: a for k, a in agg.items() if not own.match(k)}
    print(f"launches of icw:: kernels {sum(a['n'] for a in ours.values())}, total {tot:.3f} ms (cold-cache, serialised: compare shares)")
    for k, a in sorted(ours.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"{k:<58} n={a['n']:4d} total={a['ms']:9.3f} ms avg={a['ms'] / a['n']:8.3f} share={a['ms'] / tot:.3f} grid/block={sorted(a['shape'])[:3]}")
    print(f"other launches in the list (torch, input synthesis / parity sampling outside the timed region): {sum(a['n'] for a in lib.values())}, {sum(a['ms'] for a in lib.values()):.3f} ms")


def full(path, frames):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(h)}
    for r in rows[2:]:
        print(short(r[ix["Kernel Name"]]))
        vals = {}
        for k in KEEP:
            if k in ix:
                vals[k] = (r[ix[k]], units[ix[k]])
                print(f"    {k:<92} {r[ix[k]]:>18} {units[ix[k]]}")
        if frames:
            def num(k):
                v, u = vals[k]
                v = float(v.replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            try:
                print(f"    -> instructions per frame {num('smsp__inst_executed.sum') * 32 / frames:.0f} (thread), "
                      f"DRAM bytes per frame {(num('dram__bytes_read.sum') + num('dram__bytes_write.sum')) / frames:.1f}")
            except KeyError:
                pass
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 0.0)
