"""Randomised differential test: many (size, batch, inner stride, direction) geometries of both fields
against the oracle, seeded (reproducible).  Complements the fixed cases with shapes nobody thought of."""
import numpy as np
import pytest

import fieldgen
import oracle_lib as O
import pyref

pytestmark = pytest.mark.gpu


def test_random_geometries_768(ctx):
    rng = np.random.Generator(np.random.PCG64(2024))
    for case in range(40):
        logn = int(rng.integers(0, 13))
        log_r = int(rng.integers(0, 4)) if logn <= 9 else int(rng.integers(0, 2))
        batch = int(rng.choice([1, 1, 2, 3, 5, 8]))
        inverse = bool(rng.integers(0, 2))
        n, R = 1 << logn, 1 << log_r
        if batch * n * R > (1 << 14):
            batch = 1
        a = fieldgen.random_elements(batch * n * R, 5000 + case)
        w = fieldgen.omega768(n)
        d = ctx.device_alloc(a.nbytes)
        try:
            ctx.h2d(d, a)
            ctx.ntt768_device(d, n, w, inverse=inverse, batch=batch, log_r=log_r)
            got = np.empty_like(a)
            ctx.d2h(got, d)
        finally:
            ctx.device_free(d)
        v = a.reshape(batch, n, R, 24)
        g = got.reshape(batch, n, R, 24)
        for b in range(batch):
            for r in range(R):
                exp = O.fft768(np.ascontiguousarray(v[b, :, r]), w, 0 if inverse else -1, inverse=inverse)
                assert (g[b, :, r] == exp).all(), (case, logn, log_r, batch, inverse, b, r)


def test_random_geometries_32(ctx):
    rng = np.random.Generator(np.random.PCG64(2025))
    primes = [(pyref.P32, 31), (998244353, 3), (469762049, 3)]
    for case in range(40):
        mod, gen = primes[int(rng.integers(0, len(primes)))]
        logn = int(rng.integers(0, min(21, pyref.two_adicity(mod) + 1)))
        batch = int(rng.choice([1, 1, 2, 3, 7]))
        inverse = bool(rng.integers(0, 2))
        n = 1 << logn
        if batch * n > (1 << 21):
            batch = 1
        a = fieldgen.random_u32(batch * n, 7000 + case, mod)
        w = pyref.root_of_unity(mod, gen, n) if n > 1 else 1
        d = ctx.device_alloc(max(a.nbytes, 4))
        try:
            ctx.h2d(d, a)
            ctx.ntt32_device(d, n, w, mod, inverse=inverse, batch=batch)
            got = np.empty_like(a)
            ctx.d2h(got, d)
        finally:
            ctx.device_free(d)
        for b in range(batch):
            exp = O.fft32(a[b * n:(b + 1) * n], w, mod, 3 if n >= 64 else (0 if inverse else -1), inverse=inverse)
            assert (got[b * n:(b + 1) * n] == exp).all(), (case, mod, logn, batch, inverse, b)
