"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA / TMEM loads and stores / TMA / DSMEM bulk
copies / packed FFMA) in the in-tree library -> profiles/<round>_sass_summary.txt.  CPU only (cuobjdump -sass)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mobileposer_b200', 'lib', 'libmobileposer_b200.so')
rnd = sys.argv[1] if len(sys.argv) > 1 else 'r02'
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
MNEMONICS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'FFMA2', 'FFMA', 'HMMA', 'MUFU', 'STAS', 'CCTL']
counts, arch, cur = collections.OrderedDict(), set(), None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(anonymous namespace\)::', '', cur)
        cur = re.sub(r'\(.*$', '', cur).replace('void mp::', '')
        counts[cur] = collections.Counter()
        continue
    m = re.search(r'arch = (sm_\w+)', line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + '.'):
                counts[cur][mn] += 1
                break
out = [f'SASS mnemonic counts per kernel of mobileposer_b200/lib/libmobileposer_b200.so (cuobjdump -sass), architectures: {sorted(arch)}',
       'UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,',
       'UBLKCP = cp.async.bulk (DSMEM slice exchange), SYNCS = mbarrier ops, STAS = st.async, FFMA2 = packed fp32 FMA', '']
hdr = f'{"kernel":72s}' + ''.join(f'{m:>9s}' for m in MNEMONICS)
out.append(hdr)
tot = collections.Counter()
for k, c in counts.items():
    out.append(f'{k[:72]:72s}' + ''.join(f'{c[m]:9d}' for m in MNEMONICS))
    tot.update(c)
out.append(f'{"TOTAL":72s}' + ''.join(f'{tot[m]:9d}' for m in MNEMONICS))
path = os.path.join(ROOT, 'profiles', f'{rnd}_sass_summary.txt')
open(path, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[-3:]))
print('wrote', path)
