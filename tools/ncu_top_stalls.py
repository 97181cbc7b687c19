"""summarise the SASS-level `--page source --csv` output of an ncu report: top instructions by stall samples, with the dominant
stall reason, plus the totals per stall reason.  usage: ncu -i rep.ncu-rep --page source --csv | python tools/ncu_top_stalls.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
data = []
tot = {c: 0 for c in stall_cols}
total = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    total += n
    st = {c: int(r[idx[c]] or 0) for c in stall_cols}
    for c in stall_cols:
        tot[c] += st[c]
    data.append((n, r[idx["Source"]].strip(), st, r[idx["Instructions Executed"]]))
print("total samples", total)
print("by reason:", ", ".join(f"{c[6:]} {100.0 * v / max(total, 1):.1f}%" for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
for n, src, st, ex in sorted(data, key=lambda d: -d[0])[:N]:
    top = max(st, key=st.get)
    print(f"{100.0 * n / max(total, 1):5.1f}%  {src[:90]:90s}  {top[6:]} ({st[top]})  exec {ex}")
