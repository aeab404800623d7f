"""Histogram of SASS instructions per interpreter case of one K1 instantiation (instruction-cache footprint analysis).
usage: nvdisasm -g -c render_kernel.sm_100a.cubin > dis.txt; python tools/sass_footprint.py dis.txt render_block_kernelILi8ELi5"""
import collections
import re
import sys

dis, kernel = sys.argv[1], sys.argv[2]
src = open('/root/repo/elementary_b200/csrc/render_kernel.cu').read().split('\n')
label, cur = {}, 'other'
for i, l in enumerate(src, 1):
    m = re.search(r'case (OP_[A-Z0-9_]+)', l)
    if m:
        cur = m.group(1)
    m2 = re.search(r'__device__ .*?(\w+)\(', l)
    if m2 and ('__noinline__' in l or '__forceinline__' in l) and not l.startswith(' ' * 8):
        cur = 'fn:' + m2.group(1)
    if 'tile epilogue' in l:
        cur = 'epilogue'
    if 'state and parameter rows HBM' in l:
        cur = 'prologue'
    if 'state rows shared memory -> HBM' in l:
        cur = 'state_writeback'
    label[i] = cur
hist = collections.Counter()
infn, curline, curfile = False, 0, ''
for l in open(dis):
    if l.startswith('.text.'):
        infn = kernel in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        curfile, curline = m.group(1), int(m.group(2))
        continue
    if re.match(r'\s+/\*[0-9a-f]+\*/\s+[A-Z@]', l):
        hist[label.get(curline, 'other') if curfile.endswith('render_kernel.cu') else 'lib:' + curfile.split('/')[-1]] += 1
print('total', sum(hist.values()))
for k, v in hist.most_common(45):
    print(f'{v:6d} {k}')
