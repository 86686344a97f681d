#!/usr/bin/env python
"""Executed instructions / stall samples per SOURCE line of one kernel, from an ncu report.

    python tools/hot_lines.py <report.ncu-rep> <mangled kernel name> [top]

`ncu --page source` lists per-SASS-instruction counters; `nvdisasm -g` on the cubin of the same build gives the
file:line of every instruction.  The two listings are in the same order, so they are joined by instruction index.
(The library must be the build that was profiled.)
"""
import collections
import os
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

with tempfile.TemporaryDirectory() as td:
    if os.environ.get("WOST_CUBIN"):                    # a specialised (NVRTC) kernel kept by WOST_JIT_CACHE
        cubin = Path(os.environ["WOST_CUBIN"]).resolve()
    else:
        subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("WOST_LIB", str(ROOT / "dcrmontecarlo_b200" / "libwost.so"))], cwd=td, check=True, capture_output=True)
        cubin = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True, check=True).stdout
lines, cur, inside = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        inside = ln.split(".text.")[1].split()[0] == kernel
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s", ln):
        lines.append(cur)

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
assert len(data) == len(lines), (len(data), len(lines), "report and library are different builds")
ix = {h: i for i, h in enumerate(hdr)}
inst, samp, noinst, tinst = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
ops = collections.Counter()
for r, loc in zip(data, lines):
    n = int(r[ix["Instructions Executed"]])
    inst[loc] += n
    tinst[loc] += int(r[ix["Thread Instructions Executed"]])
    samp[loc] += int(r[ix["# Samples"]])
    noinst[loc] += int(r[ix["stall_no_inst"]])
    m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)", r[1])
    ops[m.group(2).split(".")[0] if m else "?"] += n
ti, ts = sum(inst.values()), sum(samp.values())
src = {}
print("total warp instr", ti, " stall samples", ts, " no_inst share %.1f%%" % (100.0 * sum(noinst.values()) / max(ts, 1)),
      " avg active lanes %.2f" % (sum(tinst.values()) / max(ti, 1)))
for loc, n in inst.most_common(top):
    f, l = loc
    if f not in src:
        p = next(iter(ROOT.rglob(f)), None)
        src[f] = p.read_text().splitlines() if p else []
    text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ""
    print("%5.1f%% instr %4.1f lanes %5.1f%% samp  %s:%4d  %s" % (100.0 * n / ti, tinst[loc] / max(n, 1), 100.0 * samp[loc] / ts, f, l, text[:100]))
print({k: round(100.0 * v / ti, 1) for k, v in ops.most_common(25)})
