"""Extracts per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum), duration, tensor-pipe activity and
occupancy from `ncu --page raw --csv` exports and writes profiles/r2_traffic.json (read by bench_support.kernel_probes for the
`roofline.traffic` field).  Usage: python tools/ncu_traffic.py profiles/r2_<kernel>_raw.csv ... """
import csv
import json
import os
import re
import sys

WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "regs",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__inst_executed.sum": "inst_executed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def parse(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    out = {}
    for r in data:
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
        short = re.sub(r"<.*", "", name).replace("lo::", "")
        rec = {"name": name}
        for m, k in WANT.items():
            if m in hdr:
                j = hdr.index(m)
                try:
                    v = float(r[j].replace(",", ""))
                except ValueError:
                    continue
                rec[k] = v * UNIT.get(units[j], 1.0)
        e = out.setdefault(short, {"launches": []})
        e["launches"].append(rec)
    return out


def main():
    tab = {}
    for p in sys.argv[1:]:
        for k, v in parse(p).items():
            tab.setdefault(k, {"launches": [], "source": []})
            tab[k]["launches"] += v["launches"]
            tab[k]["source"].append(os.path.basename(p))
    res = {}
    for k, v in tab.items():
        ls = v["launches"]
        n = len(ls)
        rd = sum(l.get("dram_read", 0) for l in ls) / n
        wr = sum(l.get("dram_write", 0) for l in ls) / n
        res[k] = {"bytes": rd + wr, "dram_read": rd, "dram_write": wr, "launches_profiled": n,
                  "us": sum(l.get("duration", 0) for l in ls) / n, "source": v["source"], "kernel": ls[0]["name"][:160]}
        for f in ("tensor_pipe_pct", "warps_active_pct", "regs", "dram_pct", "issue_active_pct"):
            vals = [l[f] for l in ls if f in l]
            if vals:
                res[k][f] = sum(vals) / len(vals)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
    json.dump(res, open(out, "w"), indent=1)
    for k, v in res.items():
        print("%-34s %8.1f us  read %7.2f MB  write %6.2f MB  tensor %5s  issue %5s  regs %s" % (
            k, v["us"], v["dram_read"] / 1e6, v["dram_write"] / 1e6, "%.1f" % v.get("tensor_pipe_pct", float("nan")),
            "%.1f" % v.get("issue_active_pct", float("nan")), v.get("regs")))


if __name__ == "__main__":
    main()
