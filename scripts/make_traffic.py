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
ALG = {"k_elastic": ("element", Ne, 4832), "k_replay": ("node", Nn, 6798)}
out = {"_doc": "DRAM traffic per launch from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum), stored per unit so that "
               f"bench.py can scale it to the workload it times.  Source: {note} ({Ne} elements, {Nn} nodes)"}
for r in rows[2:]:
    m = re.search(r"(k_\w+<[^>]*>)", r[idx["Kernel Name"]])
    if not m:
        continue
    name = m.group(1).replace(" ", "")
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
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
