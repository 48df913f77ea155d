// ntt32.cuh -- 32-bit prime-field NTT kernels (placeholder; filled in below)
#pragma once
#include <cstdint>
