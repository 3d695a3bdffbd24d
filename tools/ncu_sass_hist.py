"""Summarise an `ncu --page source --csv` dump: opcode histogram + hottest SASS lines."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
si = hdr.index('Source'); ii = hdr.index('Instructions Executed'); ss = hdr.index('# Samples'); ti = hdr.index('Thread Instructions Executed')
data = [r for r in rows[hi + 1:] if len(r) > max(ii, ss, ti) and r[ii].replace(".","").isdigit()]
tot = sum(float(r[ii]) for r in data); totS = sum(float(r[ss]) for r in data)
print('total warp-inst', tot, 'samples', totS, 'sass lines', len(data))
h = collections.Counter(); hs = collections.Counter()
for r in data:
    toks = r[si].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    h[op] += float(r[ii]); hs[op] += float(r[ss])
for op, n in h.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:10s} inst {n/tot*100:5.1f}%  samples {hs[op]/max(totS,1)*100:5.1f}%")
print('--- hottest by samples')
for r in sorted(data, key=lambda r: -float(r[ss]))[:25]:
    n = float(r[ii])
    print(f"samples {float(r[ss])/max(totS,1)*100:5.1f}% inst {n/tot*100:4.1f}% thr/inst {float(r[ti])/max(n,1):4.1f} | {r[si][:90]}")
