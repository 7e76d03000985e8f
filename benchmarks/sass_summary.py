#!/usr/bin/env python
"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that show what the kernels are made of
(cuobjdump -sass of the objects behind libfastremap_b200.so; runs without a GPU).
  python benchmarks/sass_summary.py > profiles/sass_summary.txt"""
import glob, os, re, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
COLS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDG.256", r"\bLDG\.E[.\w]*\.256"), ("LDS.128", r"\bLDS\.128"), ("REDG", r"\bREDG"),
        ("ATOMG", r"\bATOMG"), ("ATOMS", r"\bATOMS"), ("STG.128", r"\bSTG\.E[.\w]*\.128"), ("MATCH", r"\bMATCH"), ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"),
        ("CCTL/PREFETCH", r"\bCCTL")]
print("SASS mnemonic counts per kernel of fastremap_b200/libfastremap_b200.so (cuobjdump -sass of csrc/_obj/*.o, sm_100a; round 2, second session).")
print("UBLKCP = cp.async.bulk (1-D TMA bulk copy), SYNCS = mbarrier ops, LDG.256 = 256-bit table-pair loads, REDG/ATOMG = global reductions / atomics,")
print("ATOMS = shared-memory atomics.  No UTC*MMA / HMMA: nothing on this path is a contraction (tensor cores unused by design).\n")
print("%-115s %8s" % ("kernel", "instr") + "".join(" %9s" % c for c, _ in COLS))
for obj in sorted(glob.glob(os.path.join(ROOT, "fastremap_b200", "csrc", "_obj", "*.o"))):
  sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
  name, lines = None, []
  def flush():
    if name is None:
      return
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\((int|bool|unsigned int)\)", "", dem)
    dem = re.sub(r"\(.*", "", dem)
    body = "\n".join(lines)
    instr = len(re.findall(r"/\*[0-9a-f]{4}\*/", body))
    print("%-115s %8d" % (dem[:115], instr) + "".join(" %9d" % len(re.findall(rx, body)) for _, rx in COLS))
  for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
      flush()
      name, lines = m.group(1), []
    else:
      lines.append(ln)
  flush()
