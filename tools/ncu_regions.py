"""Group an `ncu --page source --csv` SASS dump into regions of equal execution count / active threads."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
si = hdr.index('Source'); ii = hdr.index('Instructions Executed'); ss = hdr.index('# Samples'); ti = hdr.index('Thread Instructions Executed')
data = [r for r in rows[hi + 1:] if len(r) > max(ii, ss, ti) and r[ii].replace(".", "").isdigit()]
tot = sum(float(r[ii]) for r in data); totS = sum(float(r[ss]) for r in data)
prev = None; start = 0; acc = 0; accS = 0; n = 0
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
for k, r in enumerate(data + [None]):
    if r is not None:
        e = float(r[ii]); t = float(r[ti]) / max(e, 1)
        key = (round(e / 1e3), round(t, 1))
    else:
        key = None
    if prev is not None and key != prev:
        if acc / tot > thr:
            print(f"lines {start:4d}-{k-1:4d} n={n:3d} exec={prev[0]:6d}k thr={prev[1]:4.1f} inst%={acc/tot*100:5.1f} samp%={accS/max(totS,1)*100:5.1f}  first: {data[start][si][:60]}")
        start = k; acc = 0; accS = 0; n = 0
    prev = key
    if r is not None:
        acc += e; accS += float(r[ss]); n += 1
print('total warp-inst', tot)
