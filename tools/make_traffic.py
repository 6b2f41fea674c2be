#!/usr/bin/env python3
"""profiles/traffic.json entry from an `ncu --set full` report of one bench launch of the headline kernel.

    python tools/make_traffic.py REPORT.ncu-rep CONFIG NBYTES KEY_CSV_PATH "NOTE"

NBYTES = the bytes one launch scans (bench.py's config.bytes_per_gpu); units are converted from ncu's raw page.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "second": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}


def main():
    rep, cfg, nbytes, key_csv, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(h)}

    def val(name):
        return float(v[col[name]].replace(",", "")) * SCALE.get(u[col[name]], 1.0)

    kernel = v[col["Kernel Name"]]
    entry = {
        "kernel": "count_lines_stream_kernel" if "count_lines_stream_kernel" in kernel else kernel.split("(")[0].split()[-1],
        "nbytes": nbytes,
        "gpu_time_ms": round(val("gpu__time_duration.sum"), 6),
        "dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
        "smem_wavefronts_per_launch": int(val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")),
        "smem_bank_conflict_wavefronts_per_launch": int(val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")),
        "sm_cycles_per_launch": int(val("sm__cycles_elapsed.avg")),
        "source": key_csv,
        "note": note,
    }
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            data = json.load(f)
    except Exception:
        data = {}
    data[cfg] = entry
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
