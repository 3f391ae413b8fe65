"""SASS evidence per kernel of libhfl.so (read here, no GPU): architecture of every cubin, and per kernel the counts of
the mnemonics that show what the kernel is made of - TMA bulk tensor stores (UTMASTG), FP64 arithmetic (DFMA / DMUL /
DADD), hardware reciprocal seeds (MUFU.RCP64H), shared-memory traffic, barriers, shuffles - and the absence of tensor
core instructions (UTC*MMA / HMMA / DMMA: there is no dense contraction on this path).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'hybrid_fem_lssvr_b200', 'libhfl.so')
WANT = ['UTMASTG', 'UTMALDG', 'DFMA', 'DMUL', 'DADD', 'MUFU.RCP64H', 'MUFU', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'SHFL', 'REDUX',
        'ATOM', 'RED', 'UTCMMA', 'HMMA', 'DMMA', 'IMMA']


def main():
    elf = subprocess.run(['cuobjdump', '-lelf', LIB], capture_output=True, text=True).stdout
    print('# cubins in libhfl.so:', ', '.join(sorted(set(re.findall(r'sm_\d+a?', elf)))), '(sm_52: the device runtime objects inside the static cudart; every hfl kernel is sm_100a)')
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(.*', '', cur).replace('void hfl::', '').replace('hfl::', '')
            counts[cur] = collections.Counter()
            continue
        m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            op = m.group(1)
            counts[cur]['total'] += 1
            for w in WANT:
                if op == w or op.startswith(w + '.') or (w == 'MUFU.RCP64H' and op.startswith('MUFU.RCP64H')):
                    counts[cur][w] += 1
    cols = ['total'] + WANT
    print('%-78s' % 'kernel' + ''.join('%12s' % c for c in cols))
    for k, c in counts.items():
        if any(s in k for s in ('lssvr_element_kernel<', 'primal_generic')) and not re.search(r'<(9|8), (16|8|32|0), ', k):
            continue          # one representative M per fine-grid size keeps the table readable
        print('%-78s' % k[:78] + ''.join('%12d' % c[x] for x in cols))
    tc = sum(c['UTCMMA'] + c['HMMA'] + c['DMMA'] + c['IMMA'] for c in counts.values())
    tma = sum(c['UTMASTG'] for c in counts.values())
    print('# tensor-core instructions in the library: %d; UTMASTG (TMA bulk tensor stores) in the library: %d over %d kernels'
          % (tc, tma, sum(1 for c in counts.values() if c['UTMASTG'])))


if __name__ == '__main__':
    main()
