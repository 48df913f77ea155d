"""Four-step driver on the GPU: single rank (no exchange) in-process, and -- when the box has
at least two GPUs -- two ranks under torchrun with the NCCL all-to-all."""
import os
import subprocess
import sys

import numpy as np
import pytest

import fieldgen
import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("logn", [4, 9, 12, 16])
def test_single_rank_fourstep_vs_oracle(ctx, logn):
    import torch
    from gpusnarks_b200 import fourstep
    dev = torch.device("cuda", 0)
    n = 1 << logn
    a = fieldgen.random_elements(n, 77 + logn)
    w = fieldgen.omega768(n)
    plan = fourstep.FourStepNTT768(fourstep.CudaBackend(ctx, dev), logn, w)
    x = torch.from_numpy(fourstep.to_column_block(a, logn, 1, 0).view(np.int32)).to(dev)
    x0 = x.clone()
    y = plan.forward(x)
    torch.cuda.synchronize()
    got = fourstep.from_row_blocks([y.cpu().numpy().view(np.uint32)], logn)
    assert (got == O.fft768(a, w, 3 if n >= 64 else -1)).all()
    back = plan.inverse(y)
    torch.cuda.synchronize()
    assert bool((back == x0).all())


def test_two_rank_fourstep_nccl():
    """two ranks under torchrun: NCCL all-to-all exchange and the fused plan (peer stores + arrival flags), both against
    the single-GPU transform and the oracle.  On a one-GPU box the multi-rank product path is still exercised by
    bench.py --gpus N (which bit-compares with the single-GPU transform) and, with one rank, by test_gpu_round2.py."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "run_fourstep_multi.py"), "14", "20", "22"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "FOURSTEP_OK" in out.stdout


def test_cpp_drop_in_driver_runs():
    """the reference-shaped C++ driver (tests/cpp/test_fft_main.cpp = reference test/main.cpp:34-87 with a real
    omega and the corrected host FFT): best_fft<fields::Scalar> / best_fft<dummy_fields::Field> through the
    header shim and the C ABI, device result == host FFT"""
    exe = os.path.join(ROOT, "tests", "cpp", "test_fft_main")
    if not os.path.exists(exe):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-fopenmp", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                               "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_fft_main.cpp"), "-L" + os.path.join(ROOT, "gpusnarks_b200"),
                               "-lgpusnarks_b200", "-Wl,-rpath," + os.path.join(ROOT, "gpusnarks_b200")])
    out = subprocess.run([exe, "16", "20"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("DONE") == 3 and "MISMATCH" not in out.stdout and "Missmatch" not in out.stdout


def test_cpp_best_fft_shards_over_the_visible_gpus():
    """the same reference-shaped C++ driver, with the sharding threshold lowered to 2^18: best_fft<fields::Scalar> then goes
    through gsn_multi_* (one process, every visible GPU, peer access) and must still equal the host FFT"""
    import gpusnarks_b200 as g
    if g.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2): with one device best_fft keeps to gsn_ntt768_host")
    exe = os.path.join(ROOT, "tests", "cpp", "test_fft_main")
    env = dict(os.environ, GSN_MULTI_MIN_LOG_N="18")
    out = subprocess.run([exe, "18", "16"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("DONE") == 3 and "MISMATCH" not in out.stdout and "Missmatch" not in out.stdout
