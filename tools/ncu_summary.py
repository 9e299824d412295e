#!/usr/bin/env python
"""ncu_summary.py -- turn an `ncu --set full` report into the markdown table kept under profiles/ and
(optionally) profiles/traffic.json (DRAM bytes per assembly = sum over the kernels of one step).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_final_ncu_level6.md [--traffic profiles/traffic.json --alg-bytes N]
"""
import argparse
import csv
import io
import json
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_tma_st.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum",
    "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--title", default="ncu --set full")
    ap.add_argument("--command", default="")
    ap.add_argument("--traffic")
    ap.add_argument("--alg-bytes", type=int, default=0)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [r[ix["Kernel Name"]].replace("grmp::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0] for r in data]
    lines = ["# " + a.title, ""]
    if a.command:
        lines += ["command: `%s`" % a.command, ""]
    lines += ["| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    for m in METRICS:
        if m in ix:
            lines.append("| %s [%s] | " % (m, units[ix[m]]) + " | ".join(r[ix[m]] for r in data) + " |")
    kernels, total = [], 0.0
    for n, r in zip(names, data):
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[ix[m]].replace(",", "")) * SCALE.get(units[ix[m]], 1.0)
        # units are per column, but ncu may pick different prefixes per row: the csv repeats one unit, values already scaled to it
        kernels.append({"kernel": n, "dram_bytes": int(b), "time_us_under_ncu": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) *
                        {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[ix["gpu__time_duration.sum"]], 1.0)})
        total += b
    lines += ["", "DRAM traffic of one assembly (sum over the kernels above): %.3f GB" % (total / 1e9) +
              (" = %.2fx the algorithmic %.3f GB." % (total / a.alg_bytes, a.alg_bytes / 1e9) if a.alg_bytes else ".")]
    if a.note:
        lines += ["", a.note]
    open(a.out, "w").write("\n".join(lines) + "\n")
    if a.traffic:
        json.dump({"dram_bytes_per_launch": int(total), "algorithmic_bytes": a.alg_bytes, "ratio": round(total / a.alg_bytes, 3) if a.alg_bytes else None,
                   "note": "sum over the kernels of one numeric assembly at level 6 (ncu --set full, --clock-control none)", "kernels": kernels},
                  open(a.traffic, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
