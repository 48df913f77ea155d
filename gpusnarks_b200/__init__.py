"""gpusnarks_b200 -- B200 (sm_100a) number-theoretic transform behind the gpusnarks FFT API.

The product is libgpusnarks_b200.so (CUDA kernels + C ABI, include/gpusnarks_b200.h); this
package is the Python mirror of the reference's host interface used by tests and bench.py.
"""
from .ntt import Context, GsnError, FIELD_FR, FIELD_FQ, device_count  # noqa: F401
