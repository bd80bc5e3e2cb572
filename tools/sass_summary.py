"""profiles/r2_sass_summary.txt: per kernel of libvosprop.so the SASS instruction count and the counts of the mnemonics that
prove (or disprove) a Blackwell-native kernel -- tcgen05.mma = UTC*MMA, TMA = UTMALDG / UBLKCP, TMEM = LDTM / STTM,
packed fp32 = FFMA2 / FMUL2 / FADD2, MUFU.EX2, plus what would betray a legacy path (HMMA = mma.sync, LDGSTS = cp.async).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
LIB = REPO / 'semi-supervised-vos_b200' / 'csrc' / 'libvosprop.so'
WATCH = ['UTCHMMA', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'LDTM', 'STTM', 'SYNCS', 'FFMA2', 'FMUL2', 'FADD2', 'MUFU.EX2', 'FMNMX3', 'HMMA', 'LDGSTS', 'STL', 'LDL']


def main():
    out = subprocess.run(['cuobjdump', '-sass', str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r'\(.*', '', name), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if m and cur is not None:
            op = m.group(1)
            cur['_total'] += 1
            for w in WATCH:
                if op == w or op.startswith(w + '.') or (w == 'MUFU.EX2' and op.startswith('MUFU.EX2')):
                    cur[w] += 1
    print(f'# SASS summary of {LIB.relative_to(REPO)} (cuobjdump -sass, sm_100a); one line per kernel instantiation')
    print('# ' + ' '.join(f'{w:>9s}' for w in ['total'] + WATCH) + '  kernel')
    tot = collections.Counter()
    for name, c in kernels.items():
        print('  ' + ' '.join(f'{c[w]:9d}' for w in ['_total'] + WATCH) + '  ' + name)
        tot.update(c)
    print('# ' + ' '.join(f'{tot[w]:9d}' for w in ['_total'] + WATCH) + '  ALL KERNELS')


if __name__ == '__main__':
    main()
