"""Planner dry run (tnb_plan_describe: host only, no GPU) over the benchmark configurations: how every contraction of
C2 / C3 / C5 is matricised, which kernel family and tile it takes, and how many waves of the 296 CTA slots (148 SMs x 2)
the launch fills -- the wave-quantisation picture behind DESIGN.md section 4.1 and the N = 8 shard analysis.
usage: python tools/plan_table.py [--md]"""
import argparse
import math
import sys

sys.path.insert(0, ".")
from itensorsgpu_b200 import tn   # noqa: E402

P = tn._lib.plan_describe
F64, C128 = tn._lib.F64, tn._lib.C128


def shapes(chi, d, w, clp=None):
    clp = clp or chi
    return {
        "i   rank3 x rank3": ((chi, d, chi), ("l", "s", "r"), (chi, w, chi), ("l", "a", "lp"), ("s", "r", "a", "lp")),
        "ii  H_eff step 1": ((chi, d, d, chi), ("l", "s1", "s2", "r"), (chi, clp, w), ("l", "lp", "a"), ("s1", "s2", "r", "lp", "a")),
        "iii H_eff step 2": ((d, d, chi, clp, w), ("s1", "s2", "r", "lp", "a"), (w, d, d, w), ("a", "s1", "s1p", "b"), ("s2", "r", "lp", "s1p", "b")),
        "iv  H_eff step 4": ((chi, clp, d, d, w), ("r", "lp", "s1p", "s2p", "c"), (chi, chi, w), ("r", "rp", "c"), ("lp", "s1p", "s2p", "rp")),
    }


def row(name, args, dtype):
    p = P(*args, dtype=dtype)
    full = math.ceil(p["waves"])
    eff = p["waves"] / full if full else 1.0
    flop = 2.0 * p["M"] * p["N"] * p["K"] * (4 if dtype == C128 else 1)
    return (name, "C128" if dtype == C128 else "F64", p["M"], p["N"], p["K"], p["family"], "%dx%dx%d" % p["tile"], p["tiles"],
            "%.2f" % p["waves"], "%.3f" % eff, "%.3g" % flop)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", action="store_true")
    a = ap.parse_args()
    hdr = ("contraction", "dtype", "M", "N", "K", "family", "tile", "CTAs", "waves", "wave efficiency", "flop")
    rows = []
    for chi in (256, 512, 1024, 2048, 4096, 8192):
        for dt in (F64, C128):
            for name, args in shapes(chi, 2, 5).items():
                rows.append(("C2 chi=%d %s" % (chi, name),) + row(name, args, dt)[1:])
    for world in (2, 4, 8):
        for name, args in shapes(4096, 2, 5, 4096 // world).items():
            if name.startswith(("ii ", "iv ")):
                rows.append(("C3 chi=4096 1/%d shard %s" % (world, name),) + row(name, args, F64)[1:])
    for name, args in shapes(8192, 2, 30, 8192 // 4).items():
        rows.append(("C5 chi=8192 w=30 slab 1/4 %s" % name,) + row(name, args, F64)[1:])
    if a.md:
        print("| " + " | ".join(hdr) + " |")
        print("|" + "---|" * len(hdr))
        for r in rows:
            print("| " + " | ".join(str(x) for x in r) + " |")
    else:
        for r in rows:
            print("  ".join(str(x).ljust(w) for x, w in zip(r, (44, 5, 9, 7, 7, 7, 10, 7, 7, 6, 9))))


if __name__ == "__main__":
    main()
