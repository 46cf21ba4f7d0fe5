#!/usr/bin/env python
"""profiles/r2_traffic.json from an `ncu --set full` raw page: the DRAM bytes per launch of the dominant kernel
(fused_update_kernel<AdamW, EMA_DIT, f32>) that bench.py quotes as `roofline.traffic`, tagged with the digest of the
kernel's sources so that bench.py refuses the figure once the code has changed.

    python tools/make_traffic.py profiles/r2_ncu_full_step_kernels_n675M_raw.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {name: hdr.index(name) for name in ("Kernel Name", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                              "gpu__time_duration.sum")}
    picked = [r for r in rows[2:] if "fused_update_kernel<2, 2, 0>" in r[col["Kernel Name"]]]
    if not picked:
        raise SystemExit("no fused_update_kernel<AdamW, EMA_DIT, f32> launch in " + path)
    r = picked[-1]
    rd = float(r[col["dram__bytes_read.sum"]]) * UNITS[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * UNITS[units[col["dram__bytes_write.sum"]]]
    rec = {"kernel": "fused_update_kernel<AdamW, EMA_DIT, f32>", "elements": bench.N3, "dram_bytes_read": rd,
           "dram_bytes_write": wr, "algorithmic_bytes": 36 * bench.N3, "grid": r[col["Grid Size"]],
           "duration_under_ncu": r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]],
           "source": os.path.relpath(path, ROOT), "kernel_source_sha256": bench.kernel_source_digest()}
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main(sys.argv[1])
