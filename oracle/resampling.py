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
form u_i = (u0 + i)/N with the same search.  Parity unpinned for it.
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
    """u_i = (u0 + i)/n evaluated with one correctly-rounded add and one divide (deterministic on CPU and GPU)."""
    return (np.arange(n, dtype=np.float64) + float(u0)) / float(n)


def fixed_point_weights(works):
    """Quantise w_i = exp(-(W_i - min W)) to integers whose total T satisfies 2^59 <= T < 2^61.

    pass 1: q1_i = rint(e_i * 2^S1), S1 = 61 - ceil(log2 N)  -> T1 = sum q1 (exact integer, <= 2^61)
    pass 2: shift s = 60 - bitlength(T1) (>= 0 is enforced), q_i = rint(e_i * 2^(S1+s))
    Returns (q uint64[N], S1 + s).
    """
    w = np.asarray(works, dtype=np.float64)
    n = w.shape[0]
    e = det_exp(-(w - w.min()))
    b = int(np.ceil(np.log2(n))) if n > 1 else 0
    s1 = 61 - b
    q1 = np.rint(e * 2.0 ** s1).astype(np.uint64)
    t1 = int(q1.sum(dtype=np.uint64))
    s = max(0, 60 - t1.bit_length())
    q = np.rint(e * 2.0 ** (s1 + s)).astype(np.uint64)
    return q, s1 + s


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


def canonical_statistics(works):
    """(sum e, sum e^2) of e_i = det_exp(min W - W_i) in the order the one-launch resampling team uses
    (coddiwomple_b200/csrc/cdw_resample.cuh): 1024 virtual threads, element i belongs to virtual thread i mod 1024 and is
    added in index order; the 1024 partials are combined by the fixed binary tree p[k] += p[k + s], s = 512 .. 1.
    With det_exp bit-identical on CPU and GPU this makes sum e bit-reproducible (the device accumulates e^2 with a fused
    multiply-add and uses its own log, so sum e^2, nESS and the normaliser agree to ~1e-15, not bitwise).
    Returns (wmin, sum_e, sum_e2, nESS, mean) with mean = wmin - log(sum_e) + log(N) (resamplers.py:107, utils.py:41-64)."""
    w = np.asarray(works, dtype=np.float64)
    n = w.size
    wmin = w.min()
    e = det_exp(wmin - w)
    V = 1024
    pe, pe2 = np.zeros(V), np.zeros(V)
    for start in range(0, n, V):          # row r of the (n / 1024, 1024) layout is added to the partials in index order
        chunk = e[start:start + V]
        pe[:chunk.size] += chunk
        pe2[:chunk.size] += chunk * chunk
    s = V // 2
    while s > 0:
        pe[:s] += pe[s:2 * s]
        pe2[:s] += pe2[s:2 * s]
        s //= 2
    sum_e, sum_e2 = float(pe[0]), float(pe2[0])
    ness = (sum_e * sum_e) / (float(n) * sum_e2)
    mean = wmin - np.log(sum_e) + np.log(float(n))
    return float(wmin), sum_e, sum_e2, float(ness), float(mean)
