import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
# find header row containing "Source" and a samples column
hdr_i = next(i for i, r in enumerate(rows) if any("Sampling" in c or "Samples" in c for c in r))
hdr = rows[hdr_i]
si = next(i for i, c in enumerate(hdr) if c.strip() in ("Source", "SASS") or c.startswith("Source"))
sm = [i for i, c in enumerate(hdr) if "Samples" in c or "Sampling" in c]
print("columns:", [hdr[i] for i in sm][:4])
agg = []
for r in rows[hdr_i + 1:]:
    try:
        v = float(r[sm[0]].replace(",", "") or 0)
    except Exception:
        continue
    agg.append((v, r[si][:110]))
tot = sum(v for v, _ in agg) or 1
agg.sort(reverse=True)
for v, s in agg[:30]:
    print(f"{100 * v / tot:5.1f}%  {s}")
