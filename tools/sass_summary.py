#!/usr/bin/env python
"""Opcode histogram of every kernel in libgraphrole_b200.so (cuobjdump -sass), written to
profiles/: evidence that the NMF kernel is tcgen05 / TMEM / TMA code (UTCHMMA, LDTM, UTMALDG,
UTMASTG, UTCBAR, SYNCS) and what the gather kernels are made of.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'graphrole_b200', 'csrc', 'libgraphrole_b200.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'UTCATOMSWS', 'SYNCS',
         'ELECT', 'UCGABAR', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'FADD', 'FFMA', 'FMNMX', 'DADD',
         'DFMA', 'DMUL', 'MUFU', 'ATOM', 'RED', 'BAR', 'HMMA', 'IMMA']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(anonymous namespace\)::', '', name)
            name = re.sub(r'\(.*', '', name)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)', line)
        if m and name:
            kernels[name][m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ('UTMALDG', 'UTMASTG', 'LDTM', 'SYNCS', 'LDG', 'STG'):
                kernels[name]['  ' + full] += 1
    lines = ['# cuobjdump -sass graphrole_b200/csrc/libgraphrole_b200.so: instructions per kernel, '
             'selected opcodes', '# (tools/sass_summary.py; sm_100a only -- `cuobjdump -lelf` lists one '
             'cubin per translation unit)', '']
    for k, c in kernels.items():
        total = sum(v for op, v in c.items() if not op.startswith('  '))
        sel = [f'{op}={c[op]}' for op in WATCH if c[op]]
        lines.append(f'{k}')
        lines.append(f'    {total} instructions; ' + ' '.join(sel))
        detail = [f'{op.strip()}={v}' for op, v in sorted(c.items()) if op.startswith('  ')]
        if detail:
            lines.append('    ' + ' '.join(detail))
    text = '\n'.join(lines) + '\n'
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r2_sass_summary.txt')
    open(dst, 'w').write(text)
    print(f'{len(kernels)} kernels -> {dst}')


if __name__ == '__main__':
    main()
