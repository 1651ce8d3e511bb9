#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into a few lines of markdown:  python scripts/ncu_summary.py rep [rep...]
  python scripts/ncu_summary.py --json key=rep [key=rep ...] > profiles/traffic_r2.json
writes, per key, the dram read + write bytes of the captured launch (what bench.py reports as roofline.traffic)."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, v, u in zip(hdr, vals, units)} for vals in rows[2:]]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main_json(pairs):
    import json

    out = {}
    for pair in pairs:
        key, rep = pair.split("=", 1)
        d = raw_rows(rep)[0]
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        out[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                    "duration_ms": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(d["gpu__time_duration.sum"][1], 1),
                    "kernel": d["Kernel Name"][0][:120], "grid": d["Grid Size"][0],
                    "tensor_pipe_active_pct": d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("", ""))[0],
                    "note": f"ncu --set full capture {rep} (one launch)"}
    print(json.dumps(out, indent=1))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--json":
        return main_json(sys.argv[2:])
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
            print(f"### {rep}\n")
            print(f"kernel: `{d['Kernel Name'][0][:110]}`  grid {d['Grid Size'][0]} block {d['Block Size'][0]}\n")
            print("| metric | value | unit |\n|---|---|---|")
            for k in WANT:
                if k in d and d[k][0] != "":
                    print(f"| {k} | {d[k][0]} | {d[k][1]} |")
            print()


if __name__ == "__main__":
    main()
