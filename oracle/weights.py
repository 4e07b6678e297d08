"""Weight bookkeeping of the SMC loop on plain arrays.  TEST INFRASTRUCTURE ONLY.

Restates, on struct-of-arrays state, what the reference does on lists of `Particle` objects:
  * `nESS`, `normalized_weights`, `vanilla_floor_threshold`      coddiwomple/utils.py:24-72
  * `TargetFactory.compute_incremental_work`                      coddiwomple/distribution_factories.py:98-131
  * `Resampler.resample` / `_update_particle` / `_null_update_particle`   coddiwomple/resamplers.py:61-200
Checked bit-for-bit against the reference's importable generic layer in tests/test_oracle_golden.py.
"""
import numpy as np
from scipy.special import logsumexp, softmax

from . import resampling


def normalized_weights(works):
    return softmax(-np.asarray(works, dtype=np.float64))  # utils.py:24-37


def nESS(cumulative_works, incremental_works=None):
    """utils.py:41-64: 1 / (N * sum softmax(-(cum+inc))^2)."""
    w_t = np.array(cumulative_works, dtype=np.float64)
    if incremental_works is not None:
        w_t = w_t + incremental_works
    w = normalized_weights(w_t)
    return 1.0 / (np.sum(w ** 2) * len(w))


def vanilla_floor_threshold(observable, floor_threshold=0.5):
    return False if observable > floor_threshold else True  # utils.py:67-72


def incremental_work(u_t, auxiliary_work, last_proposal_work=0.0):
    """distribution_factories.py:124-129 as driven by the SMC loop (openmm/coddiwomple.py:315: the flag is inverted
    in the reference, so `neglect_proposal_work=True` ADDS the last proposal work, which OMMBIP returns as 0.)."""
    return u_t + auxiliary_work + last_proposal_work


def resample_step(cum, inc, aux, labels, uniforms=None, kind="multinomial", always=False, floor_threshold=0.5):
    """One call of `Resampler.resample` (resamplers.py:61-133).

    Returns dict(cum, aux, labels, ancestors, resampled, nESS, mean).  `ancestors[i]` is the slot particle i copies
    its state / aux / root label from (identity when no resampling happens).
    kind: 'multinomial' (uniforms = N draws), 'systematic' (uniforms[0] = u0), 'identity' (the SIS base class).
    """
    cum = np.asarray(cum, dtype=np.float64)
    inc = np.asarray(inc, dtype=np.float64)
    n = len(cum)
    ness = None
    if always:
        do = True  # observable or threshold is None (resamplers.py:90-92)
    else:
        ness = float(nESS(cum, inc))
        do = vanilla_floor_threshold(ness, floor_threshold)
    if not do:
        # _null_update_particle: cum += inc, label repeats (resamplers.py:136-156)
        return dict(cum=cum + inc, aux=np.array(aux, dtype=np.float64), labels=np.array(labels), ancestors=np.arange(n),
                    resampled=False, nESS=ness, mean=None)
    updated = cum + inc
    mean = -logsumexp(-updated) + np.log(n)  # resamplers.py:107
    if kind == "identity":
        anc = np.arange(n)
    elif kind == "float":
        anc = resampling.multinomial_float(updated, uniforms)
    else:
        anc = resampling.ancestors(updated, uniforms, kind=kind)
    # _update_particle (resamplers.py:158-200): aux, state, label from the ancestor; cum += (mean - cum)
    new_cum = cum + (mean - cum)
    return dict(cum=new_cum, aux=np.asarray(aux, dtype=np.float64)[anc], labels=np.asarray(labels)[anc], ancestors=anc,
                resampled=True, nESS=ness, mean=float(mean))


def free_energy(cumulative_works):
    """-log mean exp(-cum): the SMC estimate of the reduced free energy difference (kT)."""
    c = np.asarray(cumulative_works, dtype=np.float64)
    return float(-logsumexp(-c) + np.log(len(c)))
