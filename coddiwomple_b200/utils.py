"""Utility module — same names as coddiwomple/utils.py.

`nESS` and `vanilla_floor_threshold` keep the reference signatures; when `particles` is a device-backed
`ParticleList` (coddiwomple_b200.particles) the observable is computed by the K2 kernels (cdw_weight_update +
cdw_weight_stats) instead of gathering Python floats.
"""
import numpy as np


def add_method(object, function):
    """Attach `function` to one instance under its own name, with the instance as first argument (coddiwomple/utils.py:8-19)."""
    import functools
    setattr(object, function.__name__, functools.partial(function, object))


def dummy_function():
    pass


def normalized_weights(works):
    """softmax(-works) (coddiwomple/utils.py:24-37), evaluated on the device through the K2 kernels."""
    from .kernels import device_normalized_weights
    return device_normalized_weights(works)


def nESS(particles, incremental_works=None, **kwargs):
    """Normalised effective sample size 1 / (N sum w^2), w = softmax(-(cum [+ inc])) (coddiwomple/utils.py:41-64)."""
    from .kernels import device_ness
    from .particles import ParticleList
    if isinstance(particles, ParticleList):
        cum = particles.ensemble.cum
    else:
        cum = np.fromiter((p.cumulative_work for p in particles), dtype=np.float64)
    return device_ness(cum, incremental_works)


def vanilla_floor_threshold(_observable, floor_threshold=0.5, **kwargs):
    """Resample iff the observable has dropped to the floor or below (coddiwomple/utils.py:67-72)."""
    return not (_observable > floor_threshold)
