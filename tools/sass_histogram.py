#!/usr/bin/env python
"""SASS opcode histogram of the built CUDA library, per kernel (no GPU needed).

    python tools/sass_histogram.py > profiles/r2_sass_histogram.json

Runs `cuobjdump -sass` on poreplex_b200/libporeplex_b200.so and counts, per kernel, the instructions
that show which hardware units a kernel was written for (B200_PROFILING.md's list): tcgen05 products
(UTCHMMA), TMEM traffic (LDTM / STTM), tensor-pipe barriers (UTCBAR), bulk copies by the TMA unit
(UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = cp.async.bulk.tensor), mbarrier waits (SYNCS), MUFU,
fp64 (DADD / DMUL / DFMA), and the totals.  Template instantiations are listed separately."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'poreplex_b200', 'libporeplex_b200.so')

GROUPS = collections.OrderedDict([
    ('UTCHMMA', r'^UTCHMMA'), ('UTCBAR', r'^UTCBAR'), ('LDTM', r'^LDTM'), ('STTM', r'^STTM'),
    ('UBLKCP', r'^UBLKCP'), ('UTMALDG', r'^UTMALDG'), ('UTMASTG', r'^UTMASTG'), ('SYNCS', r'^SYNCS'),
    ('MUFU', r'^MUFU'), ('F64', r'^D(ADD|MUL|FMA|SETP)'), ('FFMA', r'^FFMA'), ('FMUL_FADD', r'^F(MUL|ADD)\b'),
    ('HMMA_legacy', r'^HMMA'), ('SHFL', r'^SHFL'), ('LDG', r'^LDG'), ('STG', r'^STG'), ('LDS', r'^LDS'),
    ('STS', r'^STS'), ('LDL_STL_spill', r'^(LDL|STL)'), ('BAR', r'^BAR'),
])


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def main():
    if not os.path.exists(LIB):
        sys.exit('build the library first: python __graft_entry__.py')
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    ins = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
    for line in sass.split('\n'):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            # a kernel appears once per cubin that holds it (the linked image and its own object):
            # count its first listing only
            cur = None if m.group(1) in per else per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = ins.match(line)
        if m:
            cur[m.group(1)] += 1
    names = demangle(list(per))
    rows = []
    for k, cnt in per.items():
        total = sum(cnt.values())
        row = collections.OrderedDict(kernel=re.sub(r'^void ', '', names[k]), instructions=total)
        for g, pat in GROUPS.items():
            n = sum(v for op, v in cnt.items() if re.match(pat, op))
            if n:
                row[g] = n
        rows.append(row)
    rows.sort(key=lambda r: -r['instructions'])
    tot = collections.Counter()
    for r in rows:
        for g in GROUPS:
            tot[g] += r.get(g, 0)
    arch = subprocess.run(['cuobjdump', '-lelf', LIB], capture_output=True, text=True).stdout.strip().split('\n')
    json.dump({'library': os.path.relpath(LIB, ROOT), 'elf': arch, 'command': 'cuobjdump -sass',
               'totals': {g: tot[g] for g in GROUPS if tot[g]}, 'kernels': rows}, sys.stdout, indent=1)
    sys.stdout.write('\n')


if __name__ == '__main__':
    main()
