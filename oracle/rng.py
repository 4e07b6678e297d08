"""Counter-based noise for the oracle: Philox4x32-10 + Box-Muller.  TEST INFRASTRUCTURE ONLY.

The reference draws its Langevin noise from OpenMM's internal generator (`gaussian` in the CustomIntegrator
program, coddiwomple/openmm/integrators.py:345) and its Maxwell-Boltzmann velocities from
`Context.setVelocitiesToTemperature` (coddiwomple/openmm/propagators.py:137-138); neither stream is reproducible
outside OpenMM, so BAOAB parity with the reference is statistical.  North star: "counter-based Philox noise".
This module restates the published Philox4x32-10 algorithm (Salmon et al., SC'11; Random123 known-answer vectors
are checked in tests/test_oracle_rng.py) with the counter/key layout the CUDA kernels use, so GPU and oracle can be
fed the *same* normals:

    key     = (seed_lo, seed_hi)
    counter = (gid_lo, gid_hi, step, atom | kind << 24)        kind: 0 = O-step, 1 = Maxwell-Boltzmann, 2 = ULA
    4 x uint32 -> u_k = ((x_k >> 8) + 0.5) * 2^-24  -> two Box-Muller pairs in fp32 -> (z0, z1, z2, z3); xyz = z0..z2
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)

KIND_O_STEP = 0
KIND_MAXWELL_BOLTZMANN = 1
KIND_ULA = 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def uniform24(x):
    """((x >> 8) + 0.5) * 2^-24 as float32 — exact in fp32, strictly inside (0, 1)."""
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def box_muller(u1, u2):
    r = np.sqrt(np.float32(-2.0) * np.log(u1, dtype=np.float32), dtype=np.float32)
    ang = np.float32(2.0 * np.pi) * u2
    return r * np.cos(ang, dtype=np.float32), r * np.sin(ang, dtype=np.float32)


def normals(seed, gids, step, n_atoms, kind):
    """Standard normals of shape (P, n_atoms, 3) (float32 values held in fp64) for global particle ids `gids`."""
    gids = np.asarray(gids, dtype=np.uint64)
    atoms = np.arange(n_atoms, dtype=np.uint64)
    c0 = (gids & _MASK)[:, None]
    c1 = (gids >> np.uint64(32))[:, None]
    c3 = (atoms | np.uint64(kind << 24))[None, :]
    x0, x1, x2, x3 = philox4x32_10(c0, c1, np.uint64(step & 0xFFFFFFFF), c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = box_muller(uniform24(x0), uniform24(x1))
    z2, _ = box_muller(uniform24(x2), uniform24(x3))
    return np.stack([z0, z1, z2], axis=-1).astype(np.float64)


def uniforms53(seed, n, stream=0):
    """n uniforms in [0,1) with 53 random bits each (Philox, counter = index); used for resampling tests/bench."""
    idx = np.arange((n + 1) // 2, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10(idx & _MASK, idx >> np.uint64(32), np.uint64(stream), np.uint64(0x5EED),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    a = ((x0.astype(np.uint64) << np.uint64(32)) | x1.astype(np.uint64)) >> np.uint64(11)
    b = ((x2.astype(np.uint64) << np.uint64(32)) | x3.astype(np.uint64)) >> np.uint64(11)
    out = np.empty(2 * len(idx), dtype=np.float64)
    out[0::2] = a.astype(np.float64) * 2.0 ** -53
    out[1::2] = b.astype(np.float64) * 2.0 ** -53
    return out[:n]
