"""Unadjusted Langevin (Euler-Maruyama) proposal and its generalised SMC log-weight.  TEST INFRASTRUCTURE ONLY.

Restates SURVEY.md §8(f) rank 2:

    x' = x + tau * beta * F_t(x) + sqrt(2 tau) * xi,   tau_a = D_a dt^2,  D_a = kT / (2 m_a)
                                                                       coddiwomple/openmm/integrators.py:676-680, 829-838
    (optionally) constrain positions                                   coddiwomple/openmm/integrators.py:741-745
    proposal_work = log K(x'|x) - log L(x|x')
                  = -|x' - x - tau beta F_t(x)|^2 / (4 tau) + |x - x' - tau beta F_t(x')|^2 / (4 tau)
                                                                       coddiwomple/openmm/integrators.py:864-888
    incremental work = u_t(x') - u_{t-1}(x) + proposal_work            coddiwomple/distribution_factories.py:98-131
    i.e. -log w of  log w = log gamma_t(x') + log L(x|x') - log gamma_{t-1}(x) - log K(x'|x)
                                                                       troubleshooting/csmc.py:137-191

The toy functions of troubleshooting/csmc.py use unit mass, kT = 1 and tau = dt_csmc / 2: that is dt = sqrt(dt_csmc)
here.  `logw_three_ways` restates the three equivalent formulas of troubleshooting/test_csmc.py:119-162; the golden
fixture tests/golden/csmc_ula.json holds the reference's own outputs for them.

Precision contract as oracle/langevin.py: positions live on the fp32 grid, arithmetic is fp64.
"""
import numpy as np

from . import rng
from .langevin import _shake, r32

KIND_ULA = 2


def taus(masses, dt, kT):
    """tau_a = kT dt^2 / (2 m_a)  [nm^2]."""
    return kT * dt * dt / (2.0 * np.asarray(masses, dtype=np.float64))


def ula(energy_forces, x, masses, constraints, dt, kT, n_steps, seed, gids, step0, tol=1e-6):
    """n_steps Euler-Maruyama moves under one target.  energy_forces(x) -> (U[P] kJ/mol, F[P, A, 3] kJ/mol/nm).

    Returns dict(x, u_first, u_last, proposal_work, nan); proposal_work is dimensionless (sum over the moves)."""
    x0 = r32(x)
    x = x0.copy()
    P, A = x.shape[0], x.shape[1]
    masses = np.asarray(masses, dtype=np.float64)
    inv_m = 1.0 / masses
    tau = taus(masses, dt, kT)[None, :, None]
    drift_c = tau / kT
    U, F = energy_forces(x)
    u_first = U.copy()
    proposal = np.zeros(P)
    for s in range(n_steps):
        xi = rng.normals(seed, gids, step0 + s, A, KIND_ULA)
        x_old = x
        r0 = [(x[:, i].astype(np.float32) - x[:, j].astype(np.float32)).astype(np.float64) for (i, j, _) in constraints]
        x = r32(x_old + (drift_c * F + np.sqrt(2.0 * tau) * xi))
        if constraints:
            dummy = np.zeros_like(x)
            _shake(x, dummy, r0, inv_m, constraints, tol, 0.0)
        d = x - x_old - drift_c * F
        fwd = -((d * d) / (4.0 * tau)).sum((1, 2))
        U, F = energy_forces(x)
        d = x_old - x - drift_c * F
        bwd = -((d * d) / (4.0 * tau)).sum((1, 2))
        proposal += fwd - bwd
    nan = ~np.isfinite(U + proposal) if n_steps > 0 else np.zeros(P, dtype=bool)
    if nan.any():
        x[nan] = x0[nan]
    return dict(x=x, u_first=u_first, u_last=U.copy(), proposal_work=proposal, nan=nan)


def incremental_work(u_new, aux, proposal_work):
    """work_inc = u_t(x_t) - u_{t-1}(x_{t-1}) + proposal_work (distribution_factories.py:98-131)."""
    return u_new + aux + proposal_work


# ---------------------------------------------------------------------------------------------- csmc.py toy restatement
def toy_force(potential_grad, x, parameter):
    return -potential_grad(x, parameter)


def logw_three_ways(x, x_new, pot_old_x, pot_new_xnew, force_new_x, force_new_xnew, dt_csmc):
    """(manual, generalised, closed-form ULA) log-weights; equal in exact arithmetic (test_csmc.py:119-162).

    manual / generalised: log gamma_t(x') + log L - log gamma_{t-1}(x) - log K with Gaussian kernels of mean
    x + tau F, covariance 2 tau I, tau = dt/2 (csmc.py:31-64, 137-191); closed form: csmc.py:137-172."""
    tau = dt_csmc / 2.0
    n = x.size

    def log_normal(y, mu):
        return -0.5 * n * np.log(2.0 * np.pi * 2.0 * tau) - ((y - mu) ** 2).sum() / (4.0 * tau)

    lf = log_normal(x_new, x + tau * force_new_x)
    lb = log_normal(x, x_new + tau * force_new_xnew)
    manual = -pot_new_xnew + lb + pot_old_x - lf
    generalised = (-pot_new_xnew) + lb - (-pot_old_x) - lf
    closed = (pot_old_x + 0.5 * x.dot(force_new_x) + (dt_csmc / 8.0) * (force_new_x ** 2).sum() - pot_new_xnew
              - 0.5 * x_new.dot(force_new_xnew) - (dt_csmc / 8.0) * (force_new_xnew ** 2).sum()
              - 0.5 * x_new.dot(force_new_x) + 0.5 * x.dot(force_new_xnew))
    return manual, generalised, closed
