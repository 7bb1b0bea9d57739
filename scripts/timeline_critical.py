"""Per-family critical-path increments from a timeline_kernels_*.csv (scripts/timeline_unet.py): with programmatic
dependent launch every kernel starts early and idles until its predecessor completes, so a kernel's cost is
end_i - max(end of earlier kernels), not its duration."""
import collections
import sys

path = sys.argv[1]
nshow = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
for line in open(path).read().splitlines()[1:]:
    p = line.rsplit(",", 3)
    rows.append((p[0], float(p[1]), float(p[2])))
fam = collections.OrderedDict()
prev_end = 0.0
det = []
for name, s, d in rows:
    e = s + d
    inc = max(0.0, e - prev_end)
    prev_end = max(prev_end, e)
    k = name.replace("sdb::", "").replace("void ", "").split("<")[0]
    a = fam.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += inc
    det.append((name, s, d, inc))
tot = sum(a[1] for a in fam.values())
for k, a in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} n={a[0]:4d}  critical-path {a[1]:8.1f} us  avg {a[1] / a[0]:6.2f}  {100 * a[1] / tot:5.1f}%")
print(f"total {tot:.0f} us over {len(rows)} kernels")
for i, (n, s, d, inc) in enumerate(det[:nshow]):
    print(i, n[:44], f"start {s:8.1f} dur {d:6.1f} inc {inc:6.1f}")
