#!/usr/bin/env python
"""Attribute the static SASS of one K1 instantiation in a cubin to source constructs (needs -lineinfo).
Usage: python tools/spec_attrib.py file.cubin NITER LOGL"""
import collections, re, subprocess, sys
cubin, niter, logl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
want = ".text._ZN2eb19render_block_kernelILi%dELi%dE" % (niter, logl)
on, cur, cnt = False, None, collections.Counter()
for line in dis:
    if line.startswith(".text."):
        on = line.startswith(want)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'^\s+/\*[0-9a-f]{4,5}\*/\s+\S', line):
        cnt[cur] += 1
# render_ops.inc case ranges are found from the file itself
ops_src = open(__file__.replace("tools/spec_attrib.py", "elementary_b200/csrc/render_ops.inc")).read().splitlines()
cases = [(i + 1, re.search(r'case (OP_\w+)', l).group(1)) for i, l in enumerate(ops_src) if re.match(r'\s+case OP_\w+', l)]
def op_of(line):
    name = "ops.inc"
    for ln, n in cases:
        if ln <= line: name = n
    return name
b = collections.Counter()
for (f, l), c in cnt.items():
    if f == 'render_ops.inc': b[op_of(l)] += c
    elif f == 'render_kernel.cu': b['kernel.cu:%d' % (l // 25 * 25)] += c
    else: b[str(f)] += c
for k, v in b.most_common(): print(v, k)
print(sum(b.values()), "total")
