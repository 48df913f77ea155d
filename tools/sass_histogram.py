#!/usr/bin/env python3
"""SASS opcode histograms of the 768-bit pass kernels -> markdown (profiles/sass_opcodes_rNN.md).
usage: python tools/sass_histogram.py [libgpusnarks_b200.so] > profiles/sass_opcodes_r02.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpusnarks_b200", "libgpusnarks_b200.so")
KERNELS = ["ntt768_passILi256ELi2", "ntt768_pass2ILi1E", "ntt768_pass2ILi5E"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout.split("\n")
    funcs, cur = {}, None
    for line in sass:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            op = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", line)
            op = re.sub(r"^@!?U?P\d+\s+", "", op).split()[0].rstrip(";")
            funcs[cur].append(op)
    print("# SASS opcode histograms of the 768-bit pass kernels (cuobjdump -sass libgpusnarks_b200.so, sm_100a, round 2, final build)\n")
    print("`tools/sass_histogram.py`.  IMAD.WIDE.U32(.X) is the 32x32+64 multiply-accumulate the roofline counts; the fixed-operand")
    print("product issues 923 of them (+ 22 for addresses and the small reduction) where the CIOS product issues 1 176.\n")
    for key in KERNELS:
        for name, ops in funcs.items():
            if key in name:
                c = collections.Counter(ops)
                wide = c["IMAD.WIDE.U32"] + c["IMAD.WIDE.U32.X"] + c["IMAD.WIDE"]
                narrow = sum(v for k, v in c.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE"))
                print(f"## `{name}`\n\ntotal {len(ops)} instructions; {wide} wide multiply-accumulates, {narrow} other FMA-pipe integer instructions\n")
                print("| opcode | count |\n|---|---:|")
                for op, n in c.most_common(26):
                    print(f"| {op} | {n} |")
                print()


if __name__ == "__main__":
    main()
