#!/usr/bin/env python3
"""Regenerate tests/golden/*.json from the CPU oracle (oracle/liboracle.so).
Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fieldgen  # noqa: E402
import oracle_lib as O  # noqa: E402
import pyref  # noqa: E402


def main():
    cases = []
    for logn, seed, inverse in [(3, 1, False), (3, 2, True), (10, 3, False), (10, 4, True), (12, 5, False), (16, 6, False), (16, 7, True)]:
        n = 1 << logn
        a = fieldgen.random_elements(n, seed)
        w = fieldgen.omega768(n)
        out = O.fft768(a, w, 3 if n >= 64 else (0 if inverse else -1), inverse=inverse)
        cases.append({"logn": logn, "seed": seed, "inverse": inverse, "sha256": hashlib.sha256(out.tobytes()).hexdigest(),
                      "first": [int(x) for x in out[0]], "input_sha256": hashlib.sha256(a.tobytes()).hexdigest()})
    with open(os.path.join(HERE, "ntt768.json"), "w") as f:
        json.dump({"field": "MNT4-753 Fr", "generator": "tests/golden/make_golden.py", "cases": cases}, f, indent=1)
    cases = []
    for logn, seed, inverse in [(3, 1, False), (10, 2, True), (16, 3, False), (20, 4, False), (22, 5, False), (22, 6, True)]:
        n = 1 << logn
        a = fieldgen.random_u32(n, seed, pyref.P32)
        w = fieldgen.omega32(n)
        out = O.fft32(a, w, pyref.P32, 3 if n >= 64 else (0 if inverse else -1), inverse=inverse)
        cases.append({"logn": logn, "seed": seed, "inverse": inverse, "mod": pyref.P32, "sha256": hashlib.sha256(out.tobytes()).hexdigest(),
                      "first": [int(x) for x in out[:8]], "input_sha256": hashlib.sha256(a.tobytes()).hexdigest()})
    with open(os.path.join(HERE, "ntt32.json"), "w") as f:
        json.dump({"field": "Z/2013265921", "generator": "tests/golden/make_golden.py", "cases": cases}, f, indent=1)


if __name__ == "__main__":
    main()
