#!/usr/bin/env python
"""Static SASS instruction mix of one kernel of the built library (or of an object file):
`python scratch/sass_mix.py 'k_lnprob<0, 5, 1>' [path.so|.o]` -- counts by mnemonic, fp64 by operand form."""
import collections
import os
import re
import subprocess
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
pat = sys.argv[1]
path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, '..', 'golemflavor_b200', 'lib', 'libgolemflavor_b200.so')
out = subprocess.run('cuobjdump -sass %s | c++filt -p' % path, shell=True, stdout=subprocess.PIPE, text=True).stdout
blocks = re.split(r'\n\s*Function : ', out)
for b in blocks[1:]:
    name = b.split('\n', 1)[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    forms = collections.Counter()
    n = 0
    for line in b.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)\s*(.*?);', line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        base = op.split('.')[0]
        ops[base] += 1
        if base in ('DFMA', 'DMUL', 'DADD'):
            args = m.group(2)
            srcs = args.split(',')[1:]
            kinds = ''.join('c' if 'c[' in a else 'U' if re.search(r'\bUR', a) else 'i' if re.search(r'0x|[0-9]e|\d\.\d|-?\d+$', a.strip()) and 'R' not in a else 'R' for a in srcs)
            forms[base + ':' + kinds] += 1
    print('== %s: %d instructions' % (name[:100], n))
    print('  ' + '  '.join('%s %d' % kv for kv in ops.most_common(28)))
    print('  fp64 forms: ' + '  '.join('%s %d' % kv for kv in sorted(forms.items())))
