"""Side-by-side per-launch times of two op_table logs (same box, different knobs).  Usage: op_diff.py A.log B.log [min_us]"""
import sys
def load(p):
    rows = []
    for ln in open(p):
        f = ln.split()
        if len(f) >= 6 and f[0] in ("G", "D", "L") and f[2] == "cls":
            rows.append((f[0] + " " + f[1], float(f[4])))
    return rows
a, b = load(sys.argv[1]), load(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
da, db = dict(a), dict(b)
ta = sum(v for _, v in a); tb = sum(v for _, v in b)
print(f"total {ta:.0f} us vs {tb:.0f} us ({tb - ta:+.0f})")
for k, v in a:
    if k in db and abs(db[k] - v) >= thr:
        print(f"{k:<28s} {v:8.1f} -> {db[k]:8.1f}  ({db[k] - v:+6.1f})")
for k, v in b:
    if k not in da:
        print(f"{k:<28s}      new  {v:8.1f}")
for k, v in a:
    if k not in db:
        print(f"{k:<28s} {v:8.1f}  gone")
