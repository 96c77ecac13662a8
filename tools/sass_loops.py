#!/usr/bin/env python
"""Offline SASS check: compile one .cu to a cubin for sm_100a and report, for every kernel whose name matches a
pattern, the instruction count of its loops (backward branches) with an opcode histogram of the largest ones.
Usage: tools/sass_loops.py gpusph_b200/csrc/forces.cu 'forces_gather_kernel<(int)1, (bool)1, (bool)0, (bool)0>' [-D...]"""
import collections, re, subprocess, sys, os

def main():
    src, pat = sys.argv[1], sys.argv[2]
    extra = sys.argv[3:]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cubin = os.path.join(root, "build", "sass", os.path.basename(src) + ".cubin")
    os.makedirs(os.path.dirname(cubin), exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math" if False else "-DNOFAST",
           "-I" + os.path.join(root, "include"), "-cubin", "-o", cubin, src] + extra
    cmd = [c for c in cmd if c != "-DNOFAST"]
    subprocess.check_call(cmd)
    out = subprocess.check_output(["cuobjdump", "-sass", cubin], text=True)
    names = subprocess.check_output(["cu++filt"], input=out, text=True)
    kern = None; body = {}
    for line in names.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m: kern = m.group(1); body[kern] = []; continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and kern: body[kern].append((int(m.group(1), 16), m.group(2).strip()))
    for k, ins in body.items():
        if pat not in k: continue
        print("==", k, "instructions:", len(ins))
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index: loops.append((addr_index[tgt], i))
        for s, e in sorted(loops, key=lambda x: x[0] - x[1])[:4]:
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins[s:e + 1])
            print("  loop %04x-%04x: %d instr  " % (ins[s][0], ins[e][0], e - s + 1), dict(ops.most_common()))
            if os.environ.get("SASS_DUMP"):
                for a, t in ins[s:e + 1]: print("      %04x  %s" % (a, t))

main()
