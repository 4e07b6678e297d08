"""The slot-range emission of the scan kernel (cdw_weights.cu: first_slot) takes ceil(c n / total - u0) as first(c) = min{i : target(i) >= c}
whenever that estimate is further than n 2^-48 from an integer, and walks the exact forward map only otherwise.  This restates both in Python
(exact integers for the forward map, the kernel's double arithmetic for the estimate) and checks, on random and on adversarial inputs, that an
accepted estimate is always the exact answer — the property that keeps systematic ancestors and the multinomial guide table bit-exact."""
import math

import numpy as np


def target(i, u0, inv_n, total):
    """fixed_target(fl(fl(i + u0) * inv_n), total) of cdw_math.cuh, in exact integers."""
    u = float(np.float64(np.float64(i) + np.float64(u0)) * np.float64(inv_n))
    u53 = min(int(u * 9007199254740992.0), 9007199254740991)
    t = (u53 * total) >> 53
    return t if t < total else total - 1


def estimate(c, total, u0, n):
    """(i, accepted) as the kernel computes them."""
    n_over_total = float(np.float64(n) / np.float64(total))
    margin = float(n) * 3.5527136788005009e-15 + 1e-12
    x = float(np.float64(np.float64(c) * np.float64(n_over_total)) - np.float64(u0))
    up = math.ceil(x)
    i = min(max(int(up), 0), n)
    d = up - x
    return i, (margin < d < 1.0 - margin)


def exact_ok(i, c, total, u0, n, inv_n):
    """i == min{j in [0, n] : target(j) >= c} (target is non-decreasing in j)."""
    below = i == 0 or target(i - 1, u0, inv_n, total) < c
    at = i == n or target(i, u0, inv_n, total) >= c
    return below and at


def test_accepted_estimates_are_exact_on_random_inputs():
    r = np.random.RandomState(0)
    accepted = 0
    for trial in range(40000):
        n = int(10 ** r.uniform(0.5, 8.3))
        total = int(r.randint(1 << 30, 1 << 31)) << r.randint(28, 30)          # 2^58 .. 2^61
        u0 = float(r.random_sample()) if trial % 7 else (0.0 if trial % 14 else 1.0 - 2.0 ** -53)
        c = int(r.randint(1, 1 << 31)) * (total >> 31) + int(r.randint(0, 1 << 20))
        c = min(max(c, 1), total)
        i, ok = estimate(c, total, u0, n)
        if ok:
            accepted += 1
            assert exact_ok(i, c, total, u0, n, float(np.float64(1.0) / np.float64(n))), (n, total, u0, c, i)
    assert accepted > 30000       # the exact walk is the exception


def test_accepted_estimates_are_exact_next_to_slot_boundaries():
    """Adversarial: c chosen so that c n / total - u0 sits a few ulps to a few 1e-7 away from an integer, on both sides."""
    r = np.random.RandomState(1)
    checked = 0
    for trial in range(20000):
        n = int(10 ** r.uniform(2, 8.3))
        total = int(r.randint(1 << 30, 1 << 31)) << r.randint(28, 30)
        u0 = float(r.random_sample())
        j = int(r.randint(1, n))
        eps = float(10 ** r.uniform(-16, -6)) * (1 if trial % 2 else -1) * n
        c = int((j + u0 + eps) * total / n)                                      # target of a point just beside slot j's boundary
        c = min(max(c, 1), total)
        i, ok = estimate(c, total, u0, n)
        if ok:
            checked += 1
            assert exact_ok(i, c, total, u0, n, float(np.float64(1.0) / np.float64(n))), (n, total, u0, c, i)
    assert checked > 100
