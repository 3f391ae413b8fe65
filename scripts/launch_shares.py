"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel:
launches, mean duration, share of the captured device time; last line = composition of one bench step.
usage: python scripts/launch_shares.py profiles/r01_launches.csv > profiles/r01_launch_shares.txt"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
agg = {}
for r in rows[1:]:
    name = re.sub(r'\(.*$', '', r[ik]).replace('(int)', '').replace('(bool)', '')
    agg.setdefault(name, []).append(float(r[iv].replace(',', '')) / 1000.0)
tot = sum(sum(v) for v in agg.values())
print('# per-kernel device time from %s (ncu, cold-cache, serialised: compare SHARES)' % sys.argv[1])
print('# kernel, launches, mean us, share of all captured launches')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-75s %4d %10.1f %6.1f%%' % (k[:75], len(v), sum(v) / len(v), 100.0 * sum(v) / tot))
step = [(k, sum(v) / len(v)) for k, v in agg.items() if re.search(r'fem_reduce|fem_top|fem_backsub|fem_chunk_|fem_heads_', k)]
k2 = [(k, sum(v) / len(v), len(v)) for k, v in agg.items() if 'lssvr_element_kernel' in k]
if step and k2:
    main = max(k2, key=lambda x: x[2])
    parts = step + [(main[0], main[1])]
    s = sum(p[1] for p in parts)
    print('# one bench step = ' + ', '.join('%s %.1f us (%.0f%%)' % (k.replace('void hfl::', '').replace('hfl::', ''), t, 100 * t / s)
                                           for k, t in parts) + ' ; sum %.1f us' % s)
