#!/usr/bin/env python
"""remap (1 M entries) and mask_except (5 M kept labels) at 1024^3 uint64 with lookup tables of 2 / 4 / 8 slots per entry."""
import json, os, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for spe in ([int(a) for a in sys.argv[1:]] or [2, 4, 8]):
  code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); from fastremap_b200 import api; api.LOOKUP_SLOTS_PER_ENTRY = %d; "
          "sys.argv = ['bench_ops.py', '--ops', 'remap,mask_except', '--reps', '3']; import bench_ops; bench_ops.main()") % (ROOT, os.path.join(ROOT, "benchmarks"), spe)
  out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
  for line in out.splitlines():
    if '"op"' in line and "table build" not in line:
      d = json.loads(line)
      print(json.dumps({"slots_per_entry": spe, "op": d["op"][:40], "ms": round(d["ms"], 3)}), flush=True)
