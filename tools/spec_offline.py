#!/usr/bin/env python
"""Offline view of a specialised K1: compile render_kernel.cu against the render program of a canonical graph with nvcc (same flags
NVRTC gets), print registers / stack / static SASS size.  Usage: python tools/spec_offline.py [subsynth32|additive|plumbing] L [outdir]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from elementary_b200 import Runtime, graphs

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "subsynth32"
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    out = sys.argv[3] if len(sys.argv) > 3 else "/tmp/spec"
    os.makedirs(out, exist_ok=True)
    batch = {"subsynth32": graphs.subsynth32, "additive": lambda: graphs.additive64(110.0, 8), "plumbing": graphs.plumbing}[name]()
    rt = Runtime(48000.0, 512, 64, device=-1, tile_width=L)
    assert rt.apply_instructions(batch) == 0
    words = rt.program_words(0)
    # cut at the first OP_END, zero the pointer words (what spec_host.cpp does)
    code, pc = [], 0
    while pc + 8 <= len(words):
        w0 = int(words[pc]); n = 8 + ((w0 >> 8) & 0xFF)
        blk = [int(x) for x in words[pc:pc + n]]; blk[4] = 0; blk[5] = 0
        code += blk; pc += n
        if (w0 & 0xFF) == 0: break
    hdr = "#pragma once\n#include \"rtc_compat.h\"\nnamespace eb {\n__device__ constexpr uint32_t EB_SPEC_CODE[] = {" + ",".join(hex(w) + "u" for w in code) + \
          "};\nconstexpr int EB_SPEC_CODE_LEN = %d;\n}\n#define EB_SPEC_PROGRAM 1\n" % len(code)
    open(os.path.join(out, "eb_spec_program.h"), "w").write(hdr)
    src = os.path.join(ROOT, "elementary_b200", "csrc")
    niter = {32: 8, 16: 8, 8: 8, 4: 4, 2: 4, 1: 1}[L]
    logl = L.bit_length() - 1
    inst = os.path.join(src, "render_kernel.cu")      # the host launchers instantiate every geometry; the one asked for is picked below
    cubin = os.path.join(out, f"spec_{name}_L{L}.cubin")
    cmd = ["nvcc", "-std=c++20", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-lineinfo", "-O3", "-cubin", "-Xptxas", "-v",
           "-I", src, "-I", out, "-include", os.path.join(out, "eb_spec_program.h"), "-diag-suppress=186,68,179,177", inst, "-o", cubin] + sys.argv[4:]
    p = subprocess.run(cmd, capture_output=True, text=True)
    print(p.stderr[-3000:])
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    open(cubin + ".sass", "w").write(sass)
    fn = None; counts = {}
    for line in sass.splitlines():
        if "Function :" in line: fn = line.split(":")[1].strip(); counts[fn] = 0
        elif fn and re.match(r"^\s+/\*[0-9a-f]{4,5}\*/\s+\S", line): counts[fn] += 1
    want = "render_block_kernelILi%dELi%dE" % (niter, logl)
    for line in p.stderr.splitlines():
        if want in line: print(line.strip()[:160]); show = 3
    for k, v in counts.items():
        if want in k: print(v, "SASS instructions", k[:100])

main()
