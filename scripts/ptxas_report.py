"""Summarise `nvcc -Xptxas -v` output: registers / spills per k_march kernel."""
import re
import sys

t = sys.stdin.read()
pat = re.compile(r"Compiling entry function '([^']+)'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads.*?Used (\d+) registers", re.S)
flt = sys.argv[1] if len(sys.argv) > 1 else ''
for m in pat.finditer(t):
    name = m.group(1)
    if flt and flt not in name:
        continue
    print('%-110s regs %3s stack %4s spill st/ld %4s/%4s' % (name[:110], m.group(5), m.group(2),
                                                           m.group(3), m.group(4)))
