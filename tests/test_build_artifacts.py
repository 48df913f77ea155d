"""CPU-side checks of the built artefacts: the C-ABI library loads and exports every symbol
that include/gpusnarks_b200.h declares, fails loudly without a GPU, and the SASS contains what
the design claims (IMAD.WIDE carry chains in the NTT kernel, real multiplies in the probes)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpusnarks_b200", "libgpusnarks_b200.so")


@pytest.fixture(scope="module")
def lib():
    from gpusnarks_b200 import build
    build.build()  # no-op when up to date
    from gpusnarks_b200 import _lib
    return _lib.load()


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "gpusnarks_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from gpusnarks_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared, "gpusnarks_b200/_lib.py SYMBOLS out of sync with the header"


def test_no_cpu_fallback(lib):
    """without a GPU the product path must fail loudly, not compute on the host"""
    import ctypes as C
    cnt = C.c_int()
    lib.gsn_device_count(C.byref(cnt))
    if cnt.value > 0:
        pytest.skip("a GPU is present")
    import gpusnarks_b200 as g
    with pytest.raises(g.GsnError) as e:
        g.Context(0)
    assert e.value.code == 6 and "no CPU fallback" in str(e.value)


def test_product_code_does_not_touch_the_oracle():
    """only tests/, bench.py (cpu_baseline leg) and __graft_entry__.smoke may use oracle/"""
    bad = []
    for base in ("gpusnarks_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".h", ".cuh", ".cu", ".inl", ".cpp")):
                    with open(os.path.join(d, f), errors="ignore") as fh:
                        txt = fh.read()
                    if re.search(r'#include\s+[<"][^">]*oracle|import\s+oracle|from\s+oracle|liboracle|oracle_lib', txt):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


def _sass_histogram(pattern):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    hist, on = {}, False
    for line in out.splitlines():
        if "Function :" in line:
            on = pattern in line
            continue
        if on:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+([A-Z0-9_.]+)", line)
            if m:
                hist[m.group(1)] = hist.get(m.group(1), 0) + 1
    return hist


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_sass_is_what_the_design_claims(lib):
    # one inlined fixed-operand product per kernel: three truncated half products = 323 + 276 + 276 wide multiplies
    # (+ 48 low-only ones), and exactly one copy of it -- the small-tile kernel and the default large-tile variant
    # (which adds 64-bit index arithmetic and the 2 x 24 multiplies of reduce_small)
    for name, hi in (("ntt768_passILi256", 1000), ("ntt768_pass2ILi1", 1150)):
        h = _sass_histogram(name)
        wide = h.get("IMAD.WIDE.U32.X", 0) + h.get("IMAD.WIDE.U32", 0)
        assert 840 <= wide <= hi, (name, h)
        assert h.get("LDS.128", 0) >= 12 and h.get("STS.128", 0) >= 12
        assert not any(k.startswith(("HMMA", "UTC")) for k in h), "no tensor-core instructions expected"
        assert not any(k.startswith("LDL") or k.startswith("STL") for k in h) or name.startswith("ntt768_pass2"), (name, "local memory traffic")
    # the probes must still contain the multiplies they time (ptxas once hoisted them)
    assert _sass_histogram("int32_issue_probeILi2E").get("IMAD.WIDE.U32", 0) > 500
    assert _sass_histogram("int32_issue_probeILi4E").get("IMAD.WIDE.U32.X", 0) > 500
    assert _sass_histogram("int32_issue_probeILi0E").get("IMAD", 0) > 1000
    assert _sass_histogram("int32_issue_probeILi1E").get("IMAD.HI.U32", 0) > 1000
    assert _sass_histogram("int32_issue_probeILi6E").get("DFMA.RZ", 0) >= 64
    mixed = _sass_histogram("int32_issue_probeILi7E")
    assert mixed.get("DFMA.RZ", 0) >= 32 and mixed.get("IMAD.WIDE.U32", 0) >= 32
