"""Ancestor selection for the oracle.  TEST INFRASTRUCTURE ONLY.

Two restatements of `MultinomialResampler._resample` (coddiwomple/resamplers.py:254-272):

* `multinomial_float`  — literally what the reference executes: `np.random.choice(N, N, p=softmax(-W))`, which numpy
  implements as `cdf = p.cumsum(); cdf /= cdf[-1]; cdf.searchsorted(random_sample(N), side='right')`.  Fed with the
  uniforms numpy would have drawn, it is bit-identical to the reference (tests/test_oracle_golden.py).
* `ancestors_fixed`    — the exactly-associative form the CUDA kernels implement: weights are quantised to unsigned
  fixed point, the prefix sum and the search run in integers, so the result does not depend on summation order,
  block schedule or GPU count.  It agrees with `multinomial_float` except when a uniform falls inside the rounding
  error of numpy's sequential fp64 cumsum (tests count those; none occur in the committed cases).

Systematic resampling does not exist in the reference; it is defined here by its textbook
form u_i = (u0 + i)/N (evaluated as (u0 + i) * (1/N)) with the same search.  Parity unpinned for it.
"""
import numpy as np
from scipy.special import softmax

_M32 = np.uint64(0xFFFFFFFF)

# fdlibm split of ln 2; k*LN2_HI is exact for |k| < 2^11
_LOG2E = 1.4426950408889634
_LN2_HI = 6.93147180369123816490e-01
_LN2_LO = 1.90821492927058770002e-10
_EXP_CUTOFF = -700.0
_INV_FACT = [1.0 / float(np.prod(np.arange(1, n + 1, dtype=np.float64))) for n in range(0, 14)]


def det_exp(x):
    """exp(x) for x <= 0 built only from correctly-rounded fp64 mul/add/rint (no FMA, no libm) so that the CUDA
    kernel (which uses __dmul_rn/__dadd_rn) and this function agree bit for bit.  ~1 ulp accurate."""
    x = np.asarray(x, dtype=np.float64)
    k = np.rint(x * _LOG2E)
    r = (x - k * _LN2_HI) - k * _LN2_LO
    p = np.full_like(r, _INV_FACT[13])
    for n in range(12, -1, -1):
        p = p * r + _INV_FACT[n]
    ki = k.astype(np.int64)
    scale = ((ki + 1023).clip(1, 2046).astype(np.uint64) << np.uint64(52)).view(np.float64)
    out = p * scale
    return np.where(x < _EXP_CUTOFF, 0.0, out)


def multinomial_float(works, uniforms):
    """np.random.choice(N, len(uniforms), p=softmax(-works)) given the uniforms it would draw."""
    p = softmax(-np.asarray(works, dtype=np.float64))
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(np.asarray(uniforms, dtype=np.float64), side="right")


def systematic_uniforms(u0, n):
    """u_i = (u0 + i) * (1/n) evaluated with one correctly-rounded add and one multiply by the rounded reciprocal
    (deterministic on CPU and GPU, and a multiply instead of a divide per slot on the GPU)."""
    return (np.arange(n, dtype=np.float64) + float(u0)) * (1.0 / float(n))


def quantised_exponentials(works):
    """q_i = rint(det_exp(min W - W_i) * 2^62) as uint64 (<= 2^62): the ONE exponential per particle the kernels evaluate;
    everything downstream (statistics, cdf, search) is integer arithmetic on q.  Returns (q, e, wmin)."""
    w = np.asarray(works, dtype=np.float64)
    wmin = w.min()
    e = det_exp(wmin - w)
    return np.rint(e * 2.0 ** 62).astype(np.uint64), e, float(wmin)


def fixed_point_weights(works):
    """Quantise w_i = exp(-(W_i - min W)) to integers whose total T satisfies 2^59 <= T < 2^61.

    q62_i = rint(e_i 2^62); S1 = sum q62 (exact, Python ints); k = max(0, bitlength(S1) - 60);
    q_i = (q62_i >> k) rounded half up.  Returns (q uint64[N], k).
    """
    q62, _, _ = quantised_exponentials(works)
    s1 = int(q62.sum(dtype=object)) if q62.size < 4096 else sum(int(x) for x in np.add.reduceat(q62.astype(object), np.arange(0, q62.size, 1024)))
    k = max(0, s1.bit_length() - 60)
    if k == 0:
        return q62, 0
    q = ((q62 >> np.uint64(k - 1)) + np.uint64(1)) >> np.uint64(1)
    return q, k


def _mul_shift53(u53, total):
    """floor(u53 * total / 2^53) with u53 < 2^53 and total < 2^61, in 32-bit limbs (no overflow)."""
    u53 = np.asarray(u53, dtype=np.uint64)
    t = np.uint64(total)
    a, b = u53 >> np.uint64(32), u53 & _M32
    c, d = t >> np.uint64(32), t & _M32
    bd, ad, bc, ac = b * d, a * d, b * c, a * c
    mid = (bd >> np.uint64(32)) + (ad & _M32) + (bc & _M32)
    lo = (bd & _M32) | ((mid & _M32) << np.uint64(32))
    hi = ac + (ad >> np.uint64(32)) + (bc >> np.uint64(32)) + (mid >> np.uint64(32))
    return (hi << np.uint64(11)) | (lo >> np.uint64(53))


def fixed_point_targets(uniforms, total):
    u = np.asarray(uniforms, dtype=np.float64)
    u53 = (u * 2.0 ** 53).astype(np.uint64)  # exact: u is a double in [0,1]
    tgt = _mul_shift53(np.minimum(u53, np.uint64(2 ** 53 - 1)), total)
    return np.minimum(tgt, np.uint64(total - 1))


def ancestors_fixed(works, uniforms):
    """searchsorted(cdf_int, floor(u*T), 'right') on the integer cdf — what the CUDA scan + search computes."""
    q, _ = fixed_point_weights(works)
    cdf = np.cumsum(q, dtype=np.uint64)
    total = int(cdf[-1])
    return cdf.searchsorted(fixed_point_targets(uniforms, total), side="right").astype(np.int64)


def ancestors(works, uniforms, kind="multinomial", u0=None):
    if kind == "systematic":
        uniforms = systematic_uniforms(u0 if u0 is not None else uniforms[0], len(works))
    return ancestors_fixed(works, uniforms)


def u128_to_double(v):
    """(hi 2^64 + lo) with two correctly rounded conversions and one add — what the kernel does with its two 64-bit words."""
    hi, lo = v >> 64, v & (2 ** 64 - 1)
    return float(hi) * 18446744073709551616.0 + float(lo)


def canonical_statistics(works):
    """(sum e, sum e^2, nESS, normaliser) as the kernels define them (coddiwomple_b200/csrc/cdw_weights.cu:weight_stats_kernel):
    EXACT integer sums S1 = sum rint(e_i 2^62), S2 = sum rint((e_i e_i) 2^62) (order-free, hence identical for any grid, any
    sharding), converted once: sum_e = double(S1) 2^-62.  With det_exp bit-identical on CPU and GPU, sum e, sum e^2 and nESS
    are bit-reproducible; the normaliser uses the device's log (~1 ulp).
    Returns (wmin, sum_e, sum_e2, nESS, mean) with mean = wmin - log(sum_e) + log(N) (resamplers.py:107, utils.py:41-64)."""
    q62, e, wmin = quantised_exponentials(works)
    n = q62.size
    q2 = np.rint((e * e) * 2.0 ** 62).astype(np.uint64)
    s1 = sum(int(x) for x in q62)
    s2 = sum(int(x) for x in q2)
    sum_e, sum_e2 = u128_to_double(s1) * 2.0 ** -62, u128_to_double(s2) * 2.0 ** -62
    ness = (sum_e * sum_e) / (float(n) * sum_e2)
    mean = wmin - np.log(sum_e) + np.log(float(n))
    return float(wmin), sum_e, sum_e2, float(ness), float(mean)
