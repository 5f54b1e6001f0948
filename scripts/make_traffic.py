#!/usr/bin/env python
"""dev tool: per-unit DRAM traffic of the hot kernels from an `ncu --set full` report -> profiles/traffic.json.
usage: make_traffic.py rep.ncu-rep n_elements n_nodes [source-note]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, Ne, Nn = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# algorithmic bytes per unit: K_e kernel (gather + 4 608 B of K_e), replay, MMA fused assembly (CSR block row 1 944 B +
# per-cluster record and gather program 6 720 B / 16 nodes + 24 B of coordinates)
ALG = {"k_elastic": ("element", Ne, 4832), "k_replay": ("node", Nn, 6798), "k_assemble_hexa8_mma": ("node", Nn, 2388)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
prev = {}
if os.environ.get("TRAFFIC_MERGE") and os.path.exists(path):
    prev = json.load(open(path))
out = {"_doc": "DRAM traffic per launch from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum), stored per unit so that "
               f"bench.py can scale it to the workload it times.  Source: {note} ({Ne} elements, {Nn} nodes)"}
for r in rows[2:]:
    m = re.search(r"(k_\w+<[^>]*>)", r[idx["Kernel Name"]])
    if not m:
        continue
    name = m.group(1).replace(" ", "").replace("(bool)", "").replace("(int)", "")
    if name in out:
        continue
    kind = next((k for k in ALG if name.startswith(k)), None)
    if kind is None:
        continue
    unit, n_units, alg = ALG[kind]
    rd = float(r[idx["dram__bytes_read.sum"]]) * SCALE[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * SCALE[units[idx["dram__bytes_write.sum"]]]
    out[name] = {"read_bytes": rd, "write_bytes": wr, "units": n_units, "unit": unit, "algorithmic_bytes_per_unit": alg,
                 "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(
                     units[idx["gpu__time_duration.sum"]].replace("second", "s").replace("usecond", "us"), 1)}
if prev:  # TRAFFIC_MERGE=1: keep the kernels of earlier captures, add / replace the ones in this report
    doc = prev.get("_doc", "") + " | " + out.pop("_doc")
    prev.update(out)
    prev["_doc"] = doc
    out = prev
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
