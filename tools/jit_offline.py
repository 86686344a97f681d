#!/usr/bin/env python
"""Compile the specialised (NVRTC) walk kernel of a scenario WITHOUT a GPU and print registers / spills / code size.

    python tools/jit_offline.py cfg5 [min_blocks]      -> gpurun_out/jit_<name>.cu / .cubin
"""
import re
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dcrmontecarlo_b200 import _native as nat  # noqa: E402
from dcrmontecarlo_b200 import scenarios as sc  # noqa: E402

name = sys.argv[1]
minb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
s = sc.ALL[name]()
out = Path("gpurun_out") / f"jit_{name}"
out.parent.mkdir(exist_ok=True)
sp = {0: 0, 1: 1, 2: 2}[s.sp_mode] if s.delta else 0
g = None if (s.g is None) else s.g
print(nat.jit_offline(dict(g=g, f=s.f, alpha=s.alpha, sigma=s.sigma), neu=s.neumann is not None, src=s.f is not None, delta=s.delta,
                      sp_mode=sp, min_blocks=minb, prefix=out,
                      n_dseg=len(s.dirichlet) - 1, n_nseg=(len(s.neumann) - 1) if s.neumann is not None else 0))
res = subprocess.run(["cuobjdump", "-res-usage", f"{out}.cubin"], capture_output=True, text=True).stdout
print(res.strip().splitlines()[-1])
sass = subprocess.run(["cuobjdump", "-sass", f"{out}.cubin"], capture_output=True, text=True).stdout
n = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", sass, re.M))
print("SASS instructions:", n, "(%.1f KB)" % (n * 16 / 1024), " spill LDL/STL:", len(re.findall(r"\b(LDL|STL)\b", sass)))
