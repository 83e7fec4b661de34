"""CPU model of the shortcut the tie pass takes in `beats` (kernels.cu): among candidates of EQUAL prior whose Kahan sums are
both >= 32, the smaller sum is the larger p = pow(B, sigma) * prior — strictly, so neither pow's rounding nor the product with
the prior can merge or reorder two sums that differ by as little as one ulp. Checked here with the host libm (the reference's
own arithmetic, barcode.h:163 / pamld.cpp:73) on sums one ulp apart, where the claim is tightest; below 32 the kernel does
not use the shortcut (reference_power), and the test shows why: there one ulp of sigma no longer always moves p."""
import math

import numpy as np

BASE = math.pow(10.0, -0.1)         # PHRED_PROBABILITY_BASE, phred.h:34


def ordered_fraction(low, high, n, rng):
    sigma = rng.uniform(low, high, size=n)
    nxt = np.nextafter(sigma, np.inf)
    prior = rng.uniform(1e-6, 1.0, size=n)
    strictly = 0
    for s, t, c in zip(sigma, nxt, prior):
        strictly += (math.pow(BASE, s) * c) > (math.pow(BASE, t) * c)
    return strictly / n


def test_one_ulp_of_sigma_orders_p_from_32_on():
    rng = np.random.default_rng(5)
    for low, high in ((32.0, 64.0), (64.0, 128.0), (128.0, 1024.0), (1024.0, 2900.0)):
        assert ordered_fraction(low, high, 20000, rng) == 1.0, (low, high)


def test_below_32_the_power_has_to_be_consulted():
    rng = np.random.default_rng(6)
    # in [1, 8) an ulp of sigma is at most 2^-50: B^sigma moves by less than its own rounding unit for many sums,
    # and the two products come out equal (the reference then keeps the FIRST barcode)
    assert ordered_fraction(1.0, 8.0, 20000, rng) < 0.9
