"""CPU: exhaustive checks of arithmetic shortcuts the kernels rely on."""
from fractions import Fraction

import numpy as np


def _round_f32(x):
    """exact round-to-nearest-even of a Fraction to float32 (returned as a Fraction)."""
    if x == 0:
        return Fraction(0)
    sign = 1 if x > 0 else -1
    x = abs(x)
    e = 0
    while Fraction(2) ** e > x:
        e -= 1
    while Fraction(2) ** (e + 1) <= x:
        e += 1
    q = Fraction(2) ** (e - 23)
    n = x / q
    fl = n.numerator // n.denominator
    rem = n - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return sign * fl * q


def test_census_channel0_lut_is_the_ieee_quotient():
    """ms_fused.cu fills a 256-entry shared-memory table with __fdiv_rn(k, 120) (k <= 120) and
    1.0 elsewhere; channel 0 is a table look-up of the parked census byte.  The exact
    quotient rounded to nearest-even equals NumPy's float32 division (cbmv_generator.py:283
    divides by 120.), and clip(fill, 0, 120)/120 == 1."""
    for c in range(121):
        want = _round_f32(Fraction(c, 120))
        assert float(want) == float(np.float32(c) / np.float32(120.0))
    assert float(np.clip(np.float32(2147483648.0), 0, 120) / np.float32(120.0)) == 1.0


def test_ncc_numerator_is_exact_in_fp32():
    """ms_fused.cu computes 9*P - A_L*A_R in fp32: every operand and the result are
    integers below 2^24 for 3x3 windows of uint8."""
    assert 9 * 9 * 255 * 255 < 2 ** 24 and (9 * 255) ** 2 < 2 ** 24


def test_zsad_never_reaches_the_8192_clip():
    """sum |(L-mL)-(R-mR)| <= sum|L-mL| + sum|R-mR| <= 2 * max zero-mean L1 norm of 25
    bytes = 2 * 2*12*13/25*255 < 8192, so parked ZSAD costs survive the clip unchanged."""
    worst = 2 * (2 * 12 * 13 / 25.0) * 255
    assert worst < 8192
