#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the metrics quoted in DESIGN.md / bench.py."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main(path, title=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title or path}")
    for r in rows[2:]:
        print("kernel", r[hdr.index("Kernel Name")][:70])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:85s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")


def table(path):
    """One line per kernel: duration, DRAM bytes and achieved GB/s, the busiest units, the top stall; then a JSON list."""
    import json

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    col = {k: hdr.index(k) for k in hdr}

    def num(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except Exception:
            return float("nan")

    stalls = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    peak = 6567.1
    try:
        import os
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    recs = []
    print(f"# units: time '{rows[1][col['gpu__time_duration.sum']]}', dram '{rows[1][col['dram__bytes_read.sum']]}'")
    print(f"# {path}: one launch per kernel, ncu --set full --clock-control none; HBM peak {peak} GB/s (MEASURED_PEAKS.json)")
    print(f"{'kernel':44s} {'us':>9s} {'dramMB':>9s} {'GB/s':>7s} {'%peak':>6s} {'dram%':>6s} {'l1tex%':>6s} {'lts%':>6s} {'issue%':>6s} {'fp64%':>6s} {'occ%':>5s} {'regs':>4s}  top stall")
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        dur_us = num(r, "gpu__time_duration.sum") / 1e3  # ns in the raw page
        dram = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
        unit = rows[1][col["dram__bytes_read.sum"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1.0, "KB": 1e3, "MB": 1e6, "GB": 1e9}.get(unit, 1.0)
        dram *= scale
        tunit = rows[1][col["gpu__time_duration.sum"]]
        dur_us = num(r, "gpu__time_duration.sum") * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(tunit, 1e-3)
        gbs = dram / (dur_us * 1e-6) / 1e9 if dur_us > 0 else float("nan")
        top = max(stalls, key=lambda k: num(r, k) if num(r, k) == num(r, k) else -1.0) if stalls else ""
        rec = {"kernel": name, "us": dur_us, "dram_bytes": dram, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak,
               "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               "l1tex_pct": num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
               "lts_pct": num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "issue_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "fp64_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
               "registers": num(r, "launch__registers_per_thread"), "grid": num(r, "launch__grid_size"), "block": num(r, "launch__block_size"),
               "top_stall": top.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
               "top_stall_ratio": num(r, top) if top else float("nan"), "instructions": num(r, "smsp__inst_executed.sum")}
        recs.append(rec)
        print(f"{name[:44]:44s} {dur_us:9.1f} {dram / 1e6:9.2f} {gbs:7.0f} {100 * gbs / peak:6.1f} {rec['dram_pct']:6.1f} {rec['l1tex_pct']:6.1f} {rec['lts_pct']:6.1f} "
              f"{rec['issue_pct']:6.1f} {rec['fp64_pct']:6.1f} {rec['warps_active_pct']:5.1f} {int(rec['registers']):4d}  {rec['top_stall']} ({rec['top_stall_ratio']:.1f})")
    print("JSON " + json.dumps(recs))


if __name__ == "__main__":
    if sys.argv[1] == "--table":
        table(sys.argv[2])
    else:
        main(sys.argv[1], " ".join(sys.argv[2:]))
