"""Per CUDA source line summary of an ncu report (`--import-source on`, built with -lineinfo): executed warp
instructions and stall samples per line, top stall reasons.
usage: python scripts/ncu_lines.py report.ncu-rep [min_share_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = ''; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or r[0] == '': continue
    iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed')
    st = {h[6:]: int(r[i] if r[i] not in ("", "-") else 0) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h and i < len(r)}
    out.append((fname, int(r[0]), r[1].strip()[:90], int(r[iI] if r[iI] not in ("", "-") else 0), int(r[iS] if r[iS] not in ("", "-") else 0), st))
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print('warp instructions %.1fM, samples %d' % (ti / 1e6, ts))
for f, ln, src, n, s, st in out:
    if 100.0 * n / max(ti, 1) >= thresh or 100.0 * s / max(ts, 1) >= thresh:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print('%-22s %4d  instr %6.1fM (%4.1f%%) samples %6d (%4.1f%%) %-40s | %s' % (f, ln, n / 1e6, 100.0 * n / ti, s, 100.0 * s / ts,
              ' '.join('%s:%d' % kv for kv in top if kv[1]), src))
