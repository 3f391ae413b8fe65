"""Aggregate the SASS page of an ncu report (`ncu -i X.ncu-rep --page source --csv`) into straight-line segments:
executed warp instructions and stall samples per segment, with the dominant stall reasons.
usage: python scripts/ncu_segments.py report.ncu-rep [min_share_percent]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); isrc = hdr.index('Source')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
segs = []; cur = None
for r in data:
    n = int(r[iI] or 0); s = int(r[iS] or 0)
    st = {hdr[i][6:]: int(r[i] or 0) for i in stall}
    if cur and cur['n'] == n:
        cur['s'] += s; cur['k'] += 1; cur['b'] = r[0][-5:]
        for k, v in st.items(): cur['st'][k] = cur['st'].get(k, 0) + v
        cur['ops'].append(r[isrc].split()[0] if r[isrc].split() else '')
    else:
        cur = {'n': n, 's': s, 'k': 1, 'a': r[0][-5:], 'b': r[0][-5:], 'st': st, 'ops': [r[isrc].strip()[:40]]}
        segs.append(cur)
ti = sum(s['n'] * s['k'] for s in segs); ts = sum(s['s'] for s in segs)
print(rows[0][1]); print('warp instructions %.1fM, samples %d' % (ti / 1e6, ts))
for s in segs:
    if 100.0 * s['s'] / max(ts, 1) >= thresh or 100.0 * s['n'] * s['k'] / max(ti, 1) >= thresh:
        top = sorted(s['st'].items(), key=lambda kv: -kv[1])[:3]
        print('%s..%s exec %8d x %4d = %6.1fM (%4.1f%%)  samples %6d (%4.1f%%)  %s' % (
            s['a'], s['b'], s['n'], s['k'], s['n'] * s['k'] / 1e6, 100.0 * s['n'] * s['k'] / ti, s['s'], 100.0 * s['s'] / ts,
            ' '.join('%s:%d' % kv for kv in top if kv[1])))
