#!/usr/bin/env python
"""SASS opcode histogram of every built object (csrc/build/*.o) -> profiles/<tag>_sass_opcodes.txt.
Shows which kernels carry Blackwell tensor-core / TMA / TMEM instructions (UTCHMMA, UTMALDG, UTMASTG, LDTM, STTM,
UBLKCP, UTCBAR) and which still run on legacy HMMA.  usage: python scripts/sass_histogram.py [tag]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
KEY = ('UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'UTCCP', 'HMMA', 'SYNCS', 'MUFU',
       'FFMA', 'SHFL', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'UCGABAR', 'MAPA', 'ST.ASYNC', 'RED', 'ATOM')
out = [f'# SASS opcode histogram per object, {tag} (cuobjdump -sass; counts are static instructions)', '']
for obj in sorted((ROOT / 'forwardtacotron_b200' / 'csrc' / 'build').glob('*.o')):
    sass = subprocess.run(['cuobjdump', '-sass', str(obj)], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = subprocess.run(['cu++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = fn.replace('ftb::', '').replace('void ', '').replace('(int)', '')
            fn = fn[:fn.rfind('>(') + 1] if '>(' in fn else re.sub(r'\(.*', '', fn)
            per[fn] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and fn:
            op = m.group(1)
            per[fn][op.split('.')[0]] += 1
            if op.startswith('ST.ASYNC') or op.startswith('ST.E.ASYNC'):
                per[fn]['ST.ASYNC'] += 1
    out.append(f'## {obj.name}')
    for fn, c in per.items():
        tot = sum(v for k, v in c.items() if k != 'ST.ASYNC')
        keys = '  '.join(f'{k}={c[k]}' for k in KEY if c.get(k))
        out.append(f'{fn[:90]:90s} total={tot:6d}  {keys}')
    out.append('')
p = ROOT / 'profiles' / f'{tag}_sass_opcodes.txt'
p.write_text('\n'.join(out) + '\n')
print(p, len(out), 'lines')
