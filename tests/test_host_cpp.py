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


def test_reference_main_cpp_compiles_unmodified_multiexp_mode(tmp_path):
    """the same file WITHOUT -DFFT (test_multiexp and test_multiexp_mnt4753_G1, reference test/main.cpp:89-179): needs
    cuda/multi_exp.h, multiexp<Scalar,Scalar>, multiexp<mnt4753_G1,Scalar>, fields::mnt4753_G1 and its operators"""
    import pytest
    ref = "/root/reference/test/main.cpp"
    if not os.path.exists(ref):
        pytest.skip("reference tree absent")
    from gpusnarks_b200 import build
    build.build()
    exe = str(tmp_path / "ref_main_multiexp")
    subprocess.check_call([CXX, "-O2", "-fopenmp", "-std=c++17", "-w", "-I" + os.path.join(ROOT, "include"), "-I/root/reference/test",
                           ref, "-L" + os.path.join(ROOT, "gpusnarks_b200"), "-lgpusnarks_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "gpusnarks_b200"), "-o", exe])
    assert os.path.exists(exe)


def build_multiexp_driver(out_dir):
    from gpusnarks_b200 import build
    build.build()
    exe = os.path.join(str(out_dir), "test_multiexp_main")
    subprocess.check_call([CXX, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_multiexp_main.cpp"), "-L" + os.path.join(ROOT, "gpusnarks_b200"),
                           "-lgpusnarks_b200", "-Wl,-rpath," + os.path.join(ROOT, "gpusnarks_b200")])
    return exe


def test_multiexp_driver_compiles(tmp_path):
    """the reference-shaped multi-exponentiation driver builds against cuda/multi_exp.h (it runs under -m gpu)"""
    assert os.path.exists(build_multiexp_driver(tmp_path))


def test_host_g1_and_fp2_types_match_the_python_model(tmp_path):
    """fields::mnt4753_G1 / fields::fp2 host arithmetic (include/cuda/device_field.h, include/fields/g1_host.h) against
    tests/g1ref.py: a small C++ program computes k*P + Q and an fp2 product, Python checks the numbers"""
    import random
    import numpy as np
    import g1ref
    import pyref
    src = tmp_path / "g1host.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cuda/device_field.h>
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb");
    fields::mnt4753_G1 P, Q; fields::Scalar k; fields::fp2 a, b;
    if (fread(&P, 288, 1, f) != 1 || fread(&Q, 288, 1, f) != 1 || fread(&k, 96, 1, f) != 1 || fread(&a, 192, 1, f) != 1 || fread(&b, 192, 1, f) != 1) return 2;
    fclose(f);
    fields::mnt4753_G1 R = P * k + Q, Z = P - P, D = P + P;
    fields::fp2 c = a * b;
    f = fopen(argv[2], "wb");
    fwrite(&R, 288, 1, f); fwrite(&Z, 288, 1, f); fwrite(&D, 288, 1, f); fwrite(&c, 192, 1, f);
    fclose(f);
    return 0;
}
''')
    exe = str(tmp_path / "g1host")
    subprocess.check_call([CXX, "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", exe, str(src)])
    rng = random.Random(77)
    q = g1ref.Q
    P, Qp = g1ref.random_point(rng), g1ref.random_point(rng)
    k = rng.randrange(pyref.FR)
    a = (rng.randrange(q), rng.randrange(q))
    b = (rng.randrange(q), rng.randrange(q))

    def pt(Pt):
        return b"".join(np.array(pyref.to_limbs(v), dtype=np.uint32).tobytes() for v in g1ref.to_projective_mont(Pt))

    def fq(v):
        return np.array(pyref.to_limbs(v * pyref.RMONT % q), dtype=np.uint32).tobytes()
    (tmp_path / "in.bin").write_bytes(pt(P) + pt(Qp) + np.array(pyref.to_limbs(k), dtype=np.uint32).tobytes() + fq(a[0]) + fq(a[1]) + fq(b[0]) + fq(b[1]))
    subprocess.check_call([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    out = np.frombuffer((tmp_path / "out.bin").read_bytes(), dtype=np.uint32)
    R, Z, D = (out[i * 72:(i + 1) * 72].reshape(3, 24) for i in range(3))
    c = out[216:264].reshape(2, 24)
    aff = lambda m: g1ref.from_projective_mont(*[pyref.from_limbs(m[j]) for j in range(3)])
    assert aff(R) == g1ref.add(g1ref.mul(k, P), Qp)
    assert aff(Z) is None and aff(D) == g1ref.mul(2, P)
    rinv = pow(pyref.RMONT, -1, q)
    got = (pyref.from_limbs(c[0]) * rinv % q, pyref.from_limbs(c[1]) * rinv % q)
    assert got == ((a[0] * b[0] + 13 * a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)


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


def test_host_copy_pool(tmp_path):
    """the persistent host thread pool of the pageable paths (csrc/host_pool.h): every thread index runs exactly once per
    run(), run() is a barrier, back-to-back runs neither lose nor repeat work, a pitched copy split over it is exact"""
    exe = str(tmp_path / "test_host_pool")
    subprocess.check_call([CXX, "-O2", "-std=c++17", "-Wall", "-pthread", "-I" + os.path.join(ROOT, "gpusnarks_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_host_pool.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "test_host_pool: ok" in out.stdout, out.stdout + out.stderr


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
