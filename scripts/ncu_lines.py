"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line:
stall samples, warp instructions executed.  usage: ncu_lines.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, agg, hdr = None, {}, None
for r in rows:
    if len(r) == 2 and r[0] == 'File Name':
        fname = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == '':
        continue
    try:
        s = int(r[4])
    except ValueError:
        continue
    st = {h: int(r[i]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not' not in h and r[i].isdigit() and int(r[i]) > 0}
    agg[(fname, int(r[0]))] = (s, r[1].strip()[:100], int(r[7] or 0), st)
tot = sum(v[0] for v in agg.values()); toti = sum(v[2] for v in agg.values())
print('total samples', tot, 'total warp inst', toti)
print('--- by stall samples')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    t3 = sorted(v[3].items(), key=lambda x: -x[1])[:2]
    print("%5.1f%% %5.1f%%i %s:%d | %s | %s" % (100 * v[0] / tot, 100 * v[2] / toti, k[0], k[1], v[1], t3))
