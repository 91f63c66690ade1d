"""The constants of the hand-written device exp()/sqrt() (csrc/kernels.cuh: fast_exp_nonpos, fast_sqrt_nonneg) are checked
here on the CPU: a wrong table entry or reduction constant would only show up as a ~1e-12 parity drift on the GPU.  The
algorithm is re-stated in exact rational arithmetic on top of the parsed constants to bound its error."""
import glob
import math
import os
import re
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(glob.glob(os.path.join(ROOT, "improving*", "csrc", "kernels.cuh"))[0]).read()
getcontext().prec = 60


def parse_array(name):
    body = re.search(name + r"\[\d+\]\s*=\s*\{(.*?)\};", SRC, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [tok.strip() for tok in body.split(",") if tok.strip()]


def to_float(tok):
    if tok.startswith("0x") or tok.startswith("-0x"):
        return float.fromhex(tok)
    if "/" in tok:                                   # "1.0 / 720" style
        a, b = tok.split("/")
        return float(a) / float(b)
    return float(tok)


def test_exp2_table_is_correctly_rounded():
    tab = [to_float(t) for t in parse_array("g_exp2_tab32")]
    assert len(tab) == 32
    for j, t in enumerate(tab):
        exact = Decimal(2) ** (Decimal(j) / 32)
        assert t == float(exact), j
        # correctly rounded: the error is at most half an ulp
        assert abs(Decimal(t) - exact) <= Decimal(math.ulp(t)) / 2


def test_reduction_constants():
    red = [to_float(t) for t in parse_array("c_exp_red")]
    ln2 = Decimal(2).ln()
    assert red[0] == float(Decimal(32) / ln2)
    assert red[1] == 1.5 * 2.0 ** 52
    hi, lo = red[2], red[3]
    # hi has (at least) 21 trailing zero mantissa bits, so n * hi is exact for |n| < 2^21 (x down to -708 gives |n| < 2^16)
    m, _ = math.frexp(hi)
    assert int(m * 2 ** 53) % (1 << 21) == 0
    assert abs(Decimal(hi) + Decimal(lo) - ln2 / 32) <= Decimal(math.ulp(lo)) / 2      # lo is the correctly rounded remainder
    tay = [to_float(t) for t in parse_array("c_exp_taylor")]
    assert tay == [1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720]


def test_exp_algorithm_error_bound():
    """exp(x) = 2^n 2^(j/32) P6(r) in exact arithmetic on the device constants: the method error (table rounding + truncated
    Taylor series + split ln2) stays below 1.2e-16 relative over the whole argument range; floating-point evaluation adds
    about one more ulp."""
    tab = [Fraction(to_float(t)) for t in parse_array("g_exp2_tab32")]
    red = [to_float(t) for t in parse_array("c_exp_red")]
    rng = np.random.default_rng(0)
    xs = np.concatenate([-rng.random(300) * 40.0, -rng.random(100) * 700.0, [-1e-300, -1e-9, -0.010830, -0.010831, -707.9]])
    worst = 0.0
    for x in xs:
        k = round(x * red[0])                         # ki = n * 32 + j
        n, j = k >> 5, k & 31
        r = Fraction(float(x)) - k * Fraction(red[2]) - k * Fraction(red[3])
        assert abs(r) <= Fraction(1088, 100000)       # |r| <= ln2/64 (+ rounding of k)
        p = 1 + r + r ** 2 / 2 + r ** 3 / 6 + r ** 4 / 24 + r ** 5 / 120 + r ** 6 / 720
        approx = p * tab[j] * Fraction(2) ** n
        exact = Decimal(float(x)).exp()
        rel = abs(Decimal(approx.numerator) / Decimal(approx.denominator) / exact - 1)
        worst = max(worst, float(rel))
    assert worst < 1.2e-16, worst


def test_sqrt_iteration_converges_from_a_20_bit_seed():
    """One coupled Goldschmidt step + one Newton correction from a seed with relative error 2^-20 (MUFU.RSQ64H is better)
    reaches full double precision: the residual error is ~ (3/2 e^2)^2 / 2 ~ 2e-24."""
    for a in (1e-300, 3e-7, 0.5, 2.0, 12345.678, 1e300):
        for e0 in (2.0 ** -20, -(2.0 ** -20)):
            r = (1.0 / math.sqrt(a)) * (1.0 + e0)
            g, h = a * r, 0.5 * r
            e = 0.5 - g * h
            g, h = g + g * e, h + h * e
            g = g + (a - g * g) * h
            assert abs(g / math.sqrt(a) - 1.0) < 4e-16
