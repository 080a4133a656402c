"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST step."""
import collections
import csv
import sys

path, n_steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[1:]:
    k = r[ik].split("(")[0]
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * 1e3 if r[iu] in ("ms", "msecond") else v
    d.setdefault(k, []).append(v)
tot = 0.0
out = []
for k, v in d.items():
    n = max(len(v) // n_steps, 1)
    out.append((sum(v[-n:]), n, k))
    tot += sum(v[-n:])
for t, n, k in sorted(out, reverse=True):
    print(f"{t:10.1f} us  {100 * t / tot:5.1f}%  x{n:<3d} {k[:90]}")
print(f"{tot:10.1f} us  total of the last step ({n_steps} steps in the capture)")
