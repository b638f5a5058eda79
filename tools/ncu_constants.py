#!/usr/bin/env python3
"""Writes profiles/<out>.json: what only a profiler sees of one kernel launch, from an ncu report (--set full).
usage: tools/ncu_constants.py REPORT.ncu-rep OUT.json "source text" """
import csv, io, json, subprocess, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out, source = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (units[i], vals[i]) for i, h in enumerate(hdr)}

    def num(k):
        return float(m[k][1].replace(",", ""))

    def nbytes(k):
        return num(k) * UNIT[m[k][0]]

    j = {
        "source": source,
        "dram_bytes_per_launch": int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")),
        "l2_throughput_frac": round(num("lts__throughput.avg.pct_of_peak_sustained_elapsed") / 100, 5),
        "issue_slot_util": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100, 5),
        "active_lanes": round(num("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
        "ipc": round(num("sm__inst_executed.avg.per_cycle_active"), 3),
        "l1_hit": round(num("l1tex__t_sector_hit_rate.pct") / 100, 4),
        "l2_hit": round(num("lts__t_sector_hit_rate.pct") / 100, 4),
        "registers": int(num("launch__registers_per_thread")),
        "warps_active_frac": round(num("sm__warps_active.avg.pct_of_peak_sustained_active") / 100, 4),
        "kernel_ms_under_ncu": round(num("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3}[m["gpu__time_duration.sum"][0]], 3),
    }
    json.dump(j, open(out, "w"), indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    main()
