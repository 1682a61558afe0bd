"""DRAM traffic of one inflate_batch_kernel launch from an `ncu --set full` capture -> profiles/r02/inflate_traffic.json.
usage: make_traffic_json.py <report.ncu-rep> <algorithmic bytes per launch> "<workload>"
The JSON carries a hash of the kernel's sources; bench.py reports `roofline.traffic` only while that hash still matches."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["inflate_core.h", "kernels.cuh", "simt.h"]


def kernel_hash():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "debigulator_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    rep, alg, workload = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:inflate_batch_kernel"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u, r = rows[0], rows[1], rows[2]

    def get(k):
        v, unit = float(r[h.index(k)].replace(",", "")), u[h.index(k)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    out = {"kernel": "inflate_batch_kernel", "workload": workload, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "traffic_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": alg,
           "kernel_ms_under_ncu": float(r[h.index("gpu__time_duration.sum")]),
           "warp_instructions": float(r[h.index("smsp__inst_executed.sum")].replace(",", "")),
           "kernel_sources": KERNEL_SOURCES, "kernel_sources_sha256_16": kernel_hash(),
           "source": "ncu --set full --clock-control none, one launch; " + os.path.basename(rep)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02", "inflate_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
