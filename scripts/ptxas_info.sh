#!/bin/bash
# registers / spills of every eval_tape_kernel instantiation (development tool): scripts/ptxas_info.sh [extra nvcc flags]
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -c -o /tmp/kernels_info.o "$(dirname "$0")/../firefly_b200/csrc/kernels.cu" -Xptxas -v "$@" 2>&1 |
python3 -c '
import re, sys
txt = sys.stdin.read()
for m in re.finditer(r"Compiling entry function .(\S+). for.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
    name = m.group(1)
    k = re.search(r"eval_tape_kernelILi(\d+)ELi(\d+)ELb([01])ELb([01])ENS\S*?(FieldB|Field)E", name)
    if k:
        print("K=%s T=%s vec=%s lean=%s %-6s  regs=%s stack=%s spill_st=%s spill_ld=%s" % (*k.groups(), m.group(5), m.group(2), m.group(3), m.group(4)))
'
