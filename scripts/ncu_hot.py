"""Top SASS instructions by warp-stall samples, with the source line ncu attributes to each.
    python scripts/ncu_hot.py <file.ncu-rep> [n]"""
import sys, csv, subprocess
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
hdr = rows[hi]
isrc = hdr.index("Source")
isamp = next(i for i, c in enumerate(hdr) if c.startswith("Warp Stall Sampling (All"))
iex = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
body = rows[hi + 1:]
tot = 0; items = []
for k, r in enumerate(body):
    try: s = int(r[isamp])
    except Exception: continue
    tot += s; items.append((s, k, r[isrc], r[iex] if iex is not None else ""))
print("total samples", tot)
for s, k, t, ex in sorted(items, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}%  #{k:5d}  ex={ex:>10s}  {t[:100]}")
