#!/usr/bin/env python
"""Compile libbayesn_b200 with -Xptxas -v and print registers / spills per kernel (CPU only, no GPU needed).
usage: python scripts/ptxas_summary.py [-o out.so] [filter-substring] [-- extra nvcc flags]"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, 'bayesn_b200', 'csrc')
args = sys.argv[1:]
extra = []
if '--' in args:
    i = args.index('--'); extra = args[i + 1:]; args = args[:i]
out = '/tmp/ptxas_summary.so'
if '-o' in args:
    i = args.index('-o'); out = args[i + 1]; del args[i:i + 2]
flt = args[0] if args else ''
cmd = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++20', '-Xcompiler', '-fPIC', '-shared',
       '-I' + os.path.join(root, 'include'), '-Xptxas', '-v', '-o', out, os.path.join(src, 'bayesn_capi.cu')] + extra
p = subprocess.run(cmd, capture_output=True, text=True)
if p.returncode:
    print(p.stderr[-3000:]); sys.exit(1)
name = None
for line in p.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        spill = ''
        continue
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
    if m:
        spill = f'stack {m.group(1)} spill st/ld {m.group(2)}/{m.group(3)}'
    m = re.search(r'Used (\d+) registers', line)
    if m and name and flt in name:
        print(f'{name:90s} regs {m.group(1):>3s}  {spill}')
