"""Large randomized differential check of the oracle's geometry against the UNMODIFIED reference (build container only).

    python oracle/check_against_reference.py [--n 20000] [--ref /root/reference]

The committed fixtures (tests/golden/geometry.npz, 1 500 queries per scene) pin the oracle anywhere; this script repeats
the comparison here, where the reference can be imported, with many more queries and a different seed, and writes the
outcome to oracle/reference_check_report.json.  Bit-exact means: same float32 bit patterns for distances, silhouette
distances, ray parameters, hit points and normals; same silhouette masks, hit flags and hit segments.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

from gen_golden import import_reference  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    PolyLinesSimple, _, _ = import_reference(args.ref)
    from dcrmontecarlo_b200 import scenarios as sc
    from oracle import wost_oracle as orc

    g = torch.Generator().manual_seed(987654)
    tent = torch.tensor([[0.0, 0.0], [1.0, 1.0], [2.0, 0.0]])
    x = torch.arange(0, 12.0, 0.25)
    topo = torch.stack((x, 0.6 * torch.sin(0.9 * x) + 0.2 * torch.cos(2.3 * x)), dim=-1)
    scenes = {"square2": (sc.square(2.0), 2.2), "circle05": (sc.circle(0.5, 32), 2.0), "tent": (tent, 2.5),
              "topo": (topo, 12.0), "edge": (torch.tensor([[-100.0, 100.0], [100.0, 100.0]]), 110.0)}
    B = args.n
    report = {"queries_per_scene": B, "seed": 987654, "scenes": {}}
    for name, (pts, lim) in scenes.items():
        t0 = time.time()
        poly = PolyLinesSimple(pts)
        q = (torch.rand(B, 2, generator=g) * 2 - 1) * lim
        th = torch.rand(B, generator=g) * 2 * np.pi
        d = torch.stack([torch.cos(th), torch.sin(th)], dim=1)
        r = torch.rand(B, generator=g) * lim * 0.75 + 1e-3
        k = torch.randint(0, len(pts) - 1, (B,), generator=g)
        w = torch.rand(B, generator=g)
        q[::3] = (pts[k] * (1 - w[:, None]) + pts[k + 1] * w[:, None])[::3]       # a third start ON the polyline
        dist, sild = np.empty(B, np.float32), np.empty(B, np.float32)
        ipt, inr, ifound = np.empty((B, 2), np.float32), np.empty((B, 2), np.float32), np.empty(B, bool)
        for i in range(B):
            dist[i] = poly.distance(q[i]).item()
            sild[i] = poly.silhouetteDistance(q[i]).item()
            p_, n_, f_ = poly.intersectPolylines(q[i], d[i], r[i].item())
            ipt[i], inr[i], ifound[i] = p_.numpy(), n_.numpy(), bool(f_)
        P = pts.numpy()
        o_dist = orc.distance(P, q.numpy())
        o_sil = orc.silhouette_distance(P, q.numpy())
        o_pt, o_nr, o_found, _ = orc.intersect(P, q.numpy(), d.numpy(), r.numpy())
        res = {
            "distance_mismatches": int((bits(dist) != bits(o_dist)).sum()),
            "silhouette_distance_mismatches": int((bits(sild) != bits(o_sil)).sum()),
            "hit_flag_mismatches": int((ifound != o_found.astype(bool)).sum()),
            "hit_point_mismatches": int((bits(ipt) != bits(o_pt)).any(axis=1).sum()),
            "hit_normal_mismatches": int((bits(inr)[ifound] != bits(o_nr)[ifound]).any(axis=1).sum()),
            "hits": int(ifound.sum()), "seconds": round(time.time() - t0, 1),
        }
        report["scenes"][name] = res
        print(name, res, flush=True)
    report["all_bit_exact"] = all(v == 0 for s in report["scenes"].values() for k, v in s.items() if k.endswith("mismatches"))
    (ROOT / "oracle" / "reference_check_report.json").write_text(json.dumps(report, indent=1) + "\n")
    print("all bit-exact:", report["all_bit_exact"])


if __name__ == "__main__":
    main()
