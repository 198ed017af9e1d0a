#!/usr/bin/env python
"""Backward branches (loops) of one kernel and the instruction mix inside each: `python scratch/sass_loops.py 'k_lnprob<0, 5, 1>' [lib]`."""
import collections, os, re, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
pat = sys.argv[1]
path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, '..', 'golemflavor_b200', 'lib', 'libgolemflavor_b200.so')
out = subprocess.run('cuobjdump -sass %s | c++filt -p' % path, shell=True, stdout=subprocess.PIPE, text=True).stdout
for b in re.split(r'\n\s*Function : ', out)[1:]:
    name = b.split('\n', 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in b.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print('== %s: %d instructions' % (name[:90], len(ins)))
    for addr, txt in ins:
        m = re.search(r'\bBRA\b.*?0x([0-9a-f]+)', txt)
        if m and int(m.group(1), 16) < addr:
            lo = int(m.group(1), 16)
            body = [t for a, t in ins if lo <= a <= addr]
            ops = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for t in body)
            fp64 = sum(ops[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
            print('  loop 0x%04x..0x%04x: %4d instr, fp64 %3d (DFMA %d DMUL %d DADD %d DSETP %d) MUFU %d F2F %d FFMA %d LDC %d LDL %d STL %d CALL %d' % (
                lo, addr, len(body), fp64, ops['DFMA'], ops['DMUL'], ops['DADD'], ops['DSETP'], ops['MUFU'], ops['F2F'], ops['FFMA'] + ops['FMUL'], ops['LDC'] + ops['LDCU'], ops['LDL'], ops['STL'], ops['CALL']))
