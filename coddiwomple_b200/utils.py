"""Utility module — same names as coddiwomple/utils.py.

`nESS` and `vanilla_floor_threshold` keep the reference signatures; when `particles` is a device-backed
`ParticleList` (coddiwomple_b200.particles) the observable is computed by the K2 kernels (cdw_weight_update +
cdw_weight_stats) instead of gathering Python floats.
"""
import numpy as np


def add_method(object, function):
    """Bind a function to an object instance (coddiwomple/utils.py:8-19)."""
    from functools import partial
    setattr(object, function.__name__, partial(function, object))


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
        cum = np.array([particle.cumulative_work for particle in particles], dtype=np.float64)
    return device_ness(cum, incremental_works)


def vanilla_floor_threshold(_observable, floor_threshold=0.5, **kwargs):
    """simple floor threshold: resample iff observable <= floor (coddiwomple/utils.py:67-72)."""
    returnable = False if _observable > floor_threshold else True
    return returnable
