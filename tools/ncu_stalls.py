"""Per-instruction warp-stall samples of one kernel from an ncu --set full --import-source on
report: prints the top instructions and the totals per stall reason.
    python tools/ncu_stalls.py report.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
h = rows[hi]
ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data, tot_reason = [], collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    s = int(r[isamp] or 0)
    data.append((s, r))
    for i, c in stall_cols:
        tot_reason[c] += int(r[i] or 0)
tot = sum(s for s, _ in data)
print("total samples", tot)
print("by reason:", [(k, v, f"{100 * v / tot:.1f}%") for k, v in tot_reason.most_common(10)])
by_ex = collections.Counter()
for s, r in data:
    by_ex[int(r[iex] or 0)] += s
print("samples by execution count:", by_ex.most_common(8))
for s, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = {c: int(r[i] or 0) for i, c in stall_cols if int(r[i] or 0) > 0}
    top3 = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{s:6d} {100 * s / tot:5.1f}% ex={r[iex]:>8s} {r[ia].strip()[:64]:64s} {top3}")
