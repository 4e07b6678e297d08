"""The SMC annealing loop on arrays.  TEST INFRASTRUCTURE ONLY.

Restates `mcmc_smc_resampler` (coddiwomple/openmm/coddiwomple.py:238-333, hot loop :312-331) with the per-particle
Python objects replaced by struct-of-arrays state, in the order the reference executes it:

    init   aux = -u_0(x_0), cum = 0                       distribution_factories.py:180-220 (equip_initial_sample)
    loop   t += 1                                          coddiwomple.py:314
           inc = u_t(x_{t-1}) + aux + 0                    :315   (compute_incremental_work, proposal work is 0.)
           x, v <- BAOAB(x, v; lambda_t), n_steps = 1      :316-317 (velocities re-drawn only at t == 1)
           resample(cum, inc) with nESS <= 0.5             :318
           if t == T-1: stop                               :321-323
           aux = -u_t(x_t)                                 :326-327

Weights use the PRE-move configuration while the copy made on resampling moves the POST-move (x, v); aux for the
copied state is recomputed from the copied state, so gathering -u_t(x_t) with the state is equivalent.
"""
import numpy as np

from . import langevin, rng, weights
from .potential import kT_kj_per_mol


def lambda_table(parameter_sequence, names):
    return np.array([[float(p[k]) for k in names] for p in parameter_sequence], dtype=np.float64)


def anneal(system, x0, parameter_sequence, masses=None, constraints=None, temperature=300.0, dt=0.002, gamma=1.0,
           n_steps=1, tol=1e-6, seed=0, resample_seed=1, resampler="multinomial", always_resample=False,
           floor_threshold=0.5, gid0=0, trace=None, proposal="baoab"):
    """Run one anneal.  `system.energy_forces(x, lambdas)`; x0 (P, A, 3) nm (fp32-representable values).

    Returns dict(x, v, cum, aux, labels, resampled_at, nESS_trace, free_energy).
    """
    kT = kT_kj_per_mol(temperature)
    masses = system.masses if masses is None else np.asarray(masses, dtype=np.float64)
    constraints = system.constraints if constraints is None else constraints
    x = np.array(x0, dtype=np.float64)
    P, A = x.shape[0], x.shape[1]
    gids = np.arange(gid0, gid0 + P)
    v = np.zeros_like(x)
    cum = np.zeros(P)
    labels = np.arange(P)
    T = len(parameter_sequence)
    aux = -system.energy_forces(x, parameter_sequence[0], want_forces=False)[0] / kT
    resampled_at, ness_trace = [], []
    for t in range(1, T):
        lam = parameter_sequence[t]
        if proposal == "ula":  # Euler-Maruyama proposal + generalised weight (oracle/ula.py)
            from . import ula
            out = ula.ula(lambda xx: system.energy_forces(xx, lam), x, masses, constraints, dt, kT, n_steps, seed, gids, step0=t * n_steps, tol=tol)
            inc = np.where(out["nan"], weights.incremental_work(out["u_first"] / kT, aux),
                           ula.incremental_work(out["u_last"] / kT, aux, out["proposal_work"]))
            out["v"] = v
        else:
            out = langevin.baoab(lambda xx: system.energy_forces(xx, lam), x, v, masses, constraints, dt, gamma, kT, n_steps,
                                 seed, gids, step0=t * n_steps, tol=tol, randomize_velocities=(t == 1))
            inc = weights.incremental_work(out["u_first"] / kT, aux)
        x, v = out["x"], out["v"]
        aux_next = -out["u_last"] / kT
        # revert-on-NaN keeps the pre-move state; its aux is -u_t(x_{t-1})
        if out["nan"].any():
            aux_next = np.where(out["nan"], -out["u_first"] / kT, aux_next)
        if resampler == "multinomial":
            u = rng.uniforms53(resample_seed, P, stream=t)
        elif resampler == "systematic":
            u = rng.uniforms53(resample_seed, 1, stream=t)
        else:
            u = None
        step = weights.resample_step(cum, inc, aux_next, labels, uniforms=u, kind=resampler, always=always_resample,
                                     floor_threshold=floor_threshold)
        anc = step["ancestors"]
        cum, aux, labels = step["cum"], step["aux"], step["labels"]
        if step["resampled"]:
            x, v = x[anc], v[anc]
            resampled_at.append(t)
        ness_trace.append(step["nESS"])
        if trace is not None:
            trace.append(dict(t=t, inc=inc, cum=cum.copy(), ancestors=anc.copy(), resampled=step["resampled"], nESS=step["nESS"]))
    return dict(x=x, v=v, cum=cum, aux=aux, labels=labels, resampled_at=resampled_at, nESS_trace=ness_trace,
                free_energy=weights.free_energy(cum))
