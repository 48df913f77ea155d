"""CPU: compile and run the C++ host-side checks (product field headers vs the oracle; the
reference-shaped drop-in driver at least compiles and links against the C ABI)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def test_host_field_types_match_oracle(tmp_path):
    exe = str(tmp_path / "test_host_fields")
    subprocess.check_call([CXX, "-O2", "-fopenmp", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_host_fields.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "host fields ok" in out.stdout, out.stdout + out.stderr


def test_drop_in_driver_compiles_and_links(tmp_path):
    """reference test/main.cpp shape (best_fft<fields::Scalar>(v, omega) then the host FFT) against libgpusnarks_b200.so"""
    from gpusnarks_b200 import build
    build.build()
    exe = str(tmp_path / "test_fft_main")
    subprocess.check_call([CXX, "-O2", "-fopenmp", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_fft_main.cpp"),
                           "-L" + os.path.join(ROOT, "gpusnarks_b200"), "-lgpusnarks_b200", "-Wl,-rpath," + os.path.join(ROOT, "gpusnarks_b200")])
    assert os.path.exists(exe)


def test_reference_main_cpp_compiles_unmodified(tmp_path):
    """the reference's own test/main.cpp (-DFFT) compiles and links against this repo's headers and
    library without edits: same header paths, same template, same field type surface"""
    import pytest
    ref = "/root/reference/test/main.cpp"
    if not os.path.exists(ref):
        pytest.skip("reference tree absent")
    from gpusnarks_b200 import build
    build.build()
    exe = str(tmp_path / "ref_main")
    subprocess.check_call([CXX, "-O2", "-fopenmp", "-std=c++17", "-w", "-DFFT", "-I" + os.path.join(ROOT, "include"), "-I/root/reference/test",
                           ref, "-L" + os.path.join(ROOT, "gpusnarks_b200"), "-lgpusnarks_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "gpusnarks_b200"), "-o", exe])
    assert os.path.exists(exe)


def _build_c_smoke(tmp_path):
    from gpusnarks_b200 import build
    build.build()
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    exe = str(tmp_path / "capi_smoke")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "c", "capi_smoke.c"), "-L" + os.path.join(ROOT, "gpusnarks_b200"), "-lgpusnarks_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "gpusnarks_b200")])
    return exe


def test_c_abi_from_plain_c(tmp_path):
    """include/gpusnarks_b200.h is valid C99 and the library links from C; without a GPU the context refuses to exist"""
    exe = _build_c_smoke(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "capi ok" in out.stdout, out.stdout + out.stderr


def test_device_field_operator_test_compiles(tmp_path):
    """the reference-shaped C++ arithmetic unit test builds against the headers and the C ABI (it runs under -m gpu)"""
    from gpusnarks_b200 import build
    build.build()
    exe = str(tmp_path / "dfot")
    subprocess.check_call([CXX, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "device_field_operator_test.cpp"), "-L" + os.path.join(ROOT, "gpusnarks_b200"),
                           "-lgpusnarks_b200", "-Wl,-rpath," + os.path.join(ROOT, "gpusnarks_b200")])
    assert os.path.exists(exe)


def test_large_tile_kernel_index_model(tmp_path):
    """host replay of the warp-owned 1024-element tile of ntt768_pass2, built from the kernel's own index header
    (csrc/v2_index.h): ownership, enumeration, unit iterations, bank groups, DFT parity for digit widths 1..10"""
    exe = str(tmp_path / "test_v2_index")
    subprocess.check_call([CXX, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "gpusnarks_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_v2_index.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "test_v2_index: ok" in out.stdout, out.stdout + out.stderr


def test_fused_layout_helpers():
    """column / row layouts of the fused four-step plan (block-cyclic column ownership, [k2][r] rows) are bijections
    and agree with the definitions in include/gpusnarks_b200.h"""
    import numpy as np
    from gpusnarks_b200 import fourstep as fs
    for logn, G in [(6, 2), (14, 2), (14, 8), (16, 4), (18, 8)]:
        n = 1 << logn
        a = np.arange(n * 24, dtype=np.uint32).reshape(n, 24)
        log_n1 = fs.split_log_n1(logn)
        n1, n2 = 1 << log_n1, n >> log_n1
        C, R, rb = n2 // G, n1 // G, fs.fused_rank_bit(logn, G)
        cols = [fs.to_column_layout(a, logn, G, r) for r in range(G)]
        assert (fs.from_column_layouts(cols, logn) == a).all()
        for g in (0, G - 1):
            for i1, c in [(0, 0), (1, 1), (n1 - 1, C - 1), (n1 // 2, C // 2 + 1)]:
                i2 = ((c >> rb) << (rb + G.bit_length() - 1)) | (g << rb) | (c & ((1 << rb) - 1))
                assert (cols[g][i1, c] == a[i1 * n2 + i2]).all()
        rows = [fs.to_row_layout(a, logn, G, r) for r in range(G)]
        assert (fs.from_row_layouts(rows, logn) == a).all()
        for h in (0, G - 1):
            for k2, r in [(0, 0), (n2 - 1, R - 1), (n2 // 2, R // 2)]:
                assert (rows[h][k2, r] == a[(h * R + r) + n1 * k2]).all()
