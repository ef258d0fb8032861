"""Executable version of the argument behind quantize_scan_kernel's arithmetic (csrc/kernels.cu): the fixed-point weight
q = floor(w * 2^e), with e from the largest weight and the particle count (oracle/beluga_oracle.hpp: fixed_point_exponent,
quantize_weight -- std::ldexp), is computed on the device as floor((w * 2^e1) * 2^(e - e1)), e1 = e / 2: two multiplications
by exact powers of two instead of scalbn.  The claims checked here, with exact rationals:

* the two-step product equals ldexp(w, e) for every weight 0 <= w <= wmax, subnormals and huge values included (a power-of-two
  multiplication is exact unless the result is subnormal, and where the first step can lose bits the result is below 1 anyway);
* both factors are normal doubles for every exponent the rule can produce (the kernel builds them from the exponent field);
* n quantised weights never overflow 64 bits: every q < 2^min(52, 62 - ceil_log2 n), so the total stays below 2^62.

CPU only; it models the arithmetic, it does not run the kernel."""
import math
from fractions import Fraction

import numpy as np
import pytest


def ceil_log2(n: int) -> int:
    b = 0
    while (1 << b) < n:
        b += 1
    return b


def fixed_point_exponent(wmax: float, n_total: int) -> int:
    return min(52, 62 - ceil_log2(n_total)) - math.frexp(wmax)[1]


def pow2_double(e: int) -> float:
    """kernels.cu: pow2_double -- the bit pattern (e + 1023) << 52; needs a normal result."""
    assert -1022 <= e <= 1023, e
    return math.ldexp(1.0, e)


def kernel_quantize(w: float, exponent: int) -> int:
    e1 = int(exponent / 2)  # C++ integer division truncates towards zero
    scaled = (w * pow2_double(e1)) * pow2_double(exponent - e1)
    return math.floor(scaled) if scaled > 0.0 else 0  # NaN compares false


def oracle_quantize(w: float, exponent: int) -> int:
    if not (w > 0.0):
        return 0
    return math.floor(Fraction(w) * Fraction(2) ** exponent)  # what floor(ldexp(w, e)) is when ldexp is exact -- checked below


WMAX = [1.0, 4205.0, 1e-12, 1e-300, 5e-324, 2.2250738585072014e-308, 1e300, 1.7976931348623157e308, 0.75, 3.0000000000000004, 2.0 ** -1000,
        2.0 ** 1000]


@pytest.mark.parametrize("n_total", [1, 2, 1000, 1_000_000, 12_500_000, 100_000_000, 1 << 31])
@pytest.mark.parametrize("wmax", WMAX)
def test_two_power_of_two_multiplications_equal_ldexp(wmax, n_total):
    e = fixed_point_exponent(wmax, n_total)
    e1 = int(e / 2)
    assert -1022 <= e1 <= 1023 and -1022 <= e - e1 <= 1023  # both factors are normal doubles
    rng = np.random.default_rng(abs(hash((wmax, n_total))) % (1 << 32))
    ws = [wmax, wmax * 0.5, wmax * (1 - 2 ** -53), math.nextafter(wmax, 0.0), 0.0, 5e-324, 2.2250738585072014e-308]
    ws += list(wmax * rng.random(200)) + list(wmax * 2.0 ** -rng.integers(0, 1100, 100).astype(float))
    limit = 1 << min(52, 62 - ceil_log2(n_total))
    for w in ws:
        if not (0.0 <= w <= wmax):
            continue
        q = kernel_quantize(w, e)
        assert q == oracle_quantize(w, e), (w, e)
        assert q == (math.floor(math.ldexp(w, e)) if w > 0.0 else 0)  # and ldexp itself was exact wherever it matters
        assert q < limit or (w == wmax and q <= limit)
    # the largest weight lands in the top octave of the range: the resolution the exponent rule promises
    assert kernel_quantize(wmax, e) >= limit // 2


def test_weights_that_are_never_selected():
    for w in (0.0, -0.0, -1.0, float("nan"), -float("inf")):
        assert kernel_quantize(w, 10) == 0


@pytest.mark.parametrize("n_total", [1, 3, 1 << 20, (1 << 20) + 1, 100_000_000, 1 << 40])
def test_total_cannot_overflow(n_total):
    per_weight = 1 << min(52, 62 - ceil_log2(n_total))
    assert n_total * per_weight <= 1 << 62
