#!/usr/bin/env python
"""ncu_summary.py <report.ncu-rep> ... -> one markdown table row per captured launch (raw page)."""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "time_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", None),
    ("dram__bytes_write.sum", "dram_wr_MB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__shared_mem_per_block_dynamic", "dsmem", 1),
    ("launch__shared_mem_per_block_static", "ssmem", 1),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct", 1),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct", 1),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct", 1),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts", 1),
    ("sm__cycles_elapsed.max", "cycles", 1),
]

def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(out)))
        if len(rd) < 3:
            print(path, "unreadable"); continue
        hdr, units = rd[0], rd[1]
        for row in rd[2:]:
            d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
            name = d.get("Kernel Name", "?").split("(")[0]
            res = {"kernel": name, "grid": d.get("Grid Size"), "block": d.get("Block Size")}
            for m, label, scale in WANT:
                if m not in d or d[m] == "":
                    continue
                if scale is None:
                    res[label] = round(to_bytes(d[m], u[m]) / 1e6, 2)
                else:
                    v = float(d[m].replace(",", ""))
                    if m == "gpu__time_duration.sum":
                        v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u[m], 1e-3)
                        res[label] = round(v, 2)
                    else:
                        res[label] = round(v, 2)
            print(path.split("/")[-1], res)

if __name__ == "__main__":
    main()
