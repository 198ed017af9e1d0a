#!/usr/bin/env python
"""Registers / spills / stack of every kernel of one translation unit: `python scratch/regs.py gf_scan.cu [nvcc flags...]`."""
import re
import subprocess
import sys
import os
HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, '..', 'golemflavor_b200', 'csrc')
src = sys.argv[1]
cmd = ['nvcc', '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xptxas', '-v', '-c',
       os.path.join(CSRC, src), '-o', '/tmp/regs_%s.o' % os.path.basename(src)] + sys.argv[2:]
out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
names = subprocess.run(['c++filt', '-p'], input=out, stdout=subprocess.PIPE, text=True).stdout.splitlines()
cur = None
for line in names:
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        cur = m.group(1)
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
    if m and cur:
        stack, ss, sl = m.groups()
    m = re.search(r'Used (\d+) registers', line)
    if m and cur:
        print('%-60s regs %3s  stack %4s  spill st/ld %4s/%4s' % (cur[:60], m.group(1), stack, ss, sl))
        cur = None
if 'error' in out:
    print(out)
