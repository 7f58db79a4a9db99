#!/usr/bin/env python
"""make_ncu_replay.py <commit> <workload> <report.ncu-rep> ...  ->  profiles/ncu_replay.json

The ONE place the bench line's replayed profiler figures come from (bench.py tags them with `source` and
`captured_at_commit`): per kernel, the first captured launch of every `ncu --set full --clock-control none` report —
duration, pipe utilisation, and DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum = roofline.traffic).
Also prints the one-line-per-launch summary (tools/ncu_summary.py format) for profiles/<round>_ncu_full_summary.txt."""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_rd_bytes",
    "dram__bytes_write.sum": "dram_wr_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
}
SCALE = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3,
         "nsecond": 1e-3, "s": 1e6, "second": 1e6}


def short_name(full):
    n = re.sub(r"\(.*", "", full).replace("void ", "")
    n = re.sub(r"^.*::", "", re.sub(r"<.*", "", n))
    return n


def main():
    commit, workload = sys.argv[1], sys.argv[2]
    kernels, lines = {}, []
    for path in sys.argv[3:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(out)))
        if len(rd) < 3:
            print(path, "unreadable", file=sys.stderr)
            continue
        hdr, units = rd[0], rd[1]
        for li, row in enumerate(rd[2:]):
            d, u = dict(zip(hdr, row)), dict(zip(hdr, units))
            full = d.get("Kernel Name", "?")
            res = {"grid": d.get("Grid Size"), "block": d.get("Block Size")}
            for m, label in WANT.items():
                if d.get(m, "") == "":
                    continue
                v = float(d[m].replace(",", "")) * SCALE.get(u[m].lower(), 1)
                res[label] = round(v, 2)
            if "dram_rd_bytes" in res:
                res["dram_bytes"] = res["dram_rd_bytes"] + res.get("dram_wr_bytes", 0)
            name = short_name(full)
            tmpl = re.search(r"<([^()]*)>\(", full)
            res["template_args"] = tmpl.group(1) if tmpl else None
            res["report"] = os.path.basename(path)
            # the first launch of every kernel; further instantiations of the same kernel under "name<args>"
            key = name if name not in kernels or kernels[name]["template_args"] == res["template_args"] else \
                "%s<%s>" % (name, res["template_args"])
            kernels.setdefault(key, res)
            lines.append("%s %s" % (os.path.basename(path), dict(kernel=name + ("<%s>" % res["template_args"] if tmpl else ""), **{k: v for k, v in res.items() if k not in ("template_args", "report")})))
    out = {"_source": "ncu --set full --clock-control none, one launch per kernel after warm-up (tools/gpu_prof_all.sh); "
                      "reports: " + ", ".join(sorted({k["report"] for k in kernels.values()})),
           "_commit": commit, "_workload": workload, "kernels": kernels}
    with open(os.path.join(ROOT, "profiles", "ncu_replay.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
