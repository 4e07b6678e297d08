"""Quadratic twisting of the Euler-Maruyama proposal (controlled SMC).  TEST INFRASTRUCTURE ONLY.

Restates the twisting algebra of troubleshooting/csmc.py (the script SURVEY.md §8(f) rank 2 names):

    psi_t(x) = exp(-phi_t(x)),  phi_t(x) = x^T A_t x + b_t^T x + c_t + d_t
    Theta_t  = (I + 2 dt A_t)^-1                                               Theta_t                        csmc.py:460-478
    f_t(x)   = x + (dt / 2) F_t(x)                                             f_t / EL_mu_sigma              csmc.py:480-500, 32-64
    x' ~ N(Theta (f - dt b), dt Theta)                                         twisted_forward_proposal       csmc.py:503-528
    log K_t(psi_t)(x) = 0.5 log det Theta + (f - dt b)^T Theta (f - dt b) / (2 dt) - f^T f / (2 dt) - c - d
                                                                               twisted_forward_log_normalizer csmc.py:531-561
    log w_t = log w_t^uncontrolled + log K_t(psi_t)(x_{t-1}) + phi_t(x_t)      twisted_t_log_weight           csmc.py:593-623
    log w_0 = log Z_0 + phi_0(x_0)                                             twisted_zeroth_log_weight      csmc.py:564-591
    twisted Gaussian-mixture initial proposal and its normaliser               csmc.py:336-458

`reference_normalizer=True` evaluates the quadratic form of log K_t(psi_t) with Theta^-1, which is what csmc.py:556 literally computes
(`mahalanobis(f, dt*b, np.linalg.inv(theta))` - scipy's third argument is the matrix of the quadratic form).  The Gaussian integral
int N(x'; f, dt I) psi_t(x') dx' has Theta there (checked by quadrature in tests/test_twisted.py); the two coincide for A_t = 0, the only
regime the reference's own test runs (troubleshooting/test_csmc.py:165-226).  Both variants are pinned against outputs of the reference's
functions executed in the build container (tests/golden/make_csmc_twisted_fixture.py -> tests/golden/csmc_twisted.json).

`twisted_ula` is the batched move for molecular systems with the device kernel's parameterisation: per-atom time scale s_a = 2 tau_a
(tau_a = kT dt^2 / (2 m_a), coddiwomple/openmm/integrators.py:676-680), DIAGONAL A_t, one table a[3A] b[3A] c d per iteration.
Precision contract as oracle/langevin.py: positions live on the fp32 grid, arithmetic is fp64.
"""
import numpy as np

from . import rng
from .langevin import _shake, r32
from .ula import KIND_ULA, taus


# ------------------------------------------------------------------------------------------------ csmc.py toy restatement (dense A)
def theta_t(A, dt):
    """(I + 2 dt A)^-1 (csmc.py:460-478)."""
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    return np.linalg.inv(np.eye(A.shape[0]) + 2.0 * dt * A)


def twisted_forward_moments(theta, f, dt, b):
    """Mean and covariance of the twisted forward proposal (csmc.py:503-528)."""
    return theta @ (np.asarray(f) - dt * np.asarray(b)), dt * theta


def twisted_forward_log_normalizer(theta, f, b, c, d, dt, reference_normalizer=False):
    """log K_t(psi_t)(x) (csmc.py:531-561); see the module docstring for `reference_normalizer`."""
    f, b = np.asarray(f, dtype=np.float64), np.asarray(b, dtype=np.float64)
    g = f - dt * b
    M = np.linalg.inv(theta) if reference_normalizer else theta
    return 0.5 * np.log(np.linalg.det(theta)) + (g @ M @ g) / (2.0 * dt) - f.dot(f) / (2.0 * dt) - c - d


def phi(x, A, b, c, d=0.0):
    x = np.asarray(x, dtype=np.float64)
    return x @ np.atleast_2d(A) @ x + np.asarray(b).dot(x) + c + d


def twisted_t_log_weight(x, uncontrolled_logw, log_forward_normalizer, A, b, c, d):
    """csmc.py:593-623."""
    return uncontrolled_logw + log_forward_normalizer + phi(x, A, b, c, d)


def twisted_zeroth_log_weight(x0, A0, b0, c0, log_pi0_normalizer):
    """csmc.py:564-591."""
    return log_pi0_normalizer + phi(x0, A0, b0, c0)


def gmm_log_zeta(Sigma_tilde, Sigma, mu, b0, c0):
    """log int N(x; mu, Sigma) psi_0(x) dx of one mixture component (csmc.py:432-458)."""
    Si = np.linalg.inv(Sigma)
    m = Si @ mu - b0
    return (-0.5 * np.log(np.linalg.det(Sigma)) + 0.5 * np.log(np.linalg.det(Sigma_tilde)) + 0.5 * m @ Sigma_tilde @ m - 0.5 * mu @ Si @ mu - c0)


def twisted_gmm_components(alphas, mus, Sigmas, A0, b0, c0):
    """Unnormalised log mixing weights and covariances of the twisted mixture (csmc.py:336-374)."""
    Sigmas = np.asarray(Sigmas, dtype=np.float64)
    tilde = np.linalg.inv(np.linalg.inv(Sigmas) + 2.0 * np.asarray(A0, dtype=np.float64))
    log_zetas = np.array([gmm_log_zeta(st, s, m, np.asarray(b0), c0) for st, s, m in zip(tilde, Sigmas, np.asarray(mus, dtype=np.float64))])
    return np.log(np.asarray(alphas, dtype=np.float64)) + log_zetas, tilde


def twisted_gmm_moments(Sigma_tilde, Sigma, mu, b0):
    """Mean / covariance of one twisted component (csmc.py:376-415)."""
    return Sigma_tilde @ (np.linalg.inv(Sigma) @ mu - b0), Sigma_tilde


def twisted_gmm_lognormalizer(log_alpha_tildes):
    """csmc.py:417-430."""
    return np.logaddexp.reduce(log_alpha_tildes)


# ------------------------------------------------------------------------------------------------ the kernel's parameterisation
def pack_twist(a, b, c=0.0, d=0.0):
    """a[A, 3], b[A, 3], c, d -> the 6 A + 2 table of cdw_langevin_params.twist."""
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape and a.size % 3 == 0
    return np.concatenate([a, b, [float(c), float(d)]])


def twisted_ula(energy_forces, x, masses, constraints, dt, kT, n_steps, seed, gids, step0, twist, tol=1e-6, reference_normalizer=False):
    """n_steps twisted Euler-Maruyama moves under one target; `twist` = pack_twist(...) (None: oracle.ula.ula's move).

    Returns dict(x, u_first, u_last, proposal_work, twist_logk, twist_phi, nan): proposal_work (dimensionless) already holds
    -(log K_t(psi_t)(x) + phi_t(x')) of every move, i.e. incremental work = u_t(x') + aux + proposal_work as in oracle.ula."""
    x0 = r32(x)
    x = x0.copy()
    P, A = x.shape[0], x.shape[1]
    masses = np.asarray(masses, dtype=np.float64)
    inv_m = 1.0 / masses
    tau = taus(masses, dt, kT)[None, :, None]
    drift_c = tau / kT
    if twist is not None:
        tw = np.asarray(twist, dtype=np.float64)
        ta, tb, tc, td = tw[:3 * A].reshape(1, A, 3), tw[3 * A:6 * A].reshape(1, A, 3), tw[6 * A], tw[6 * A + 1]
    U, F = energy_forces(x)
    u_first = U.copy()
    proposal, logk_sum, phi_sum = np.zeros(P), np.zeros(P), np.zeros(P)
    for s in range(n_steps):
        xi = rng.normals(seed, gids, step0 + s, A, KIND_ULA)
        x_old = x
        r0 = [(x[:, i].astype(np.float32) - x[:, j].astype(np.float32)).astype(np.float64) for (i, j, _) in constraints]
        if twist is None:
            x = r32(x_old + (drift_c * F + np.sqrt(2.0 * tau) * xi))
        else:
            f = x_old + drift_c * F
            sc = 2.0 * tau
            th = 1.0 / (1.0 + 2.0 * sc * ta)
            g = f - sc * tb
            x = r32(th * g + np.sqrt(sc * th) * xi)
            quad = g * g / th if reference_normalizer else g * g * th
            logk = (0.5 * np.log(th) + (quad - f * f) / (2.0 * sc)).sum((1, 2)) - tc - td
        if constraints:
            dummy = np.zeros_like(x)
            _shake(x, dummy, r0, inv_m, constraints, tol, 0.0)
        dlt = x - x_old - drift_c * F
        fwd = -((dlt * dlt) / (4.0 * tau)).sum((1, 2))
        U, F = energy_forces(x)
        dlt = x_old - x - drift_c * F
        bwd = -((dlt * dlt) / (4.0 * tau)).sum((1, 2))
        proposal += fwd - bwd
        if twist is not None:
            ph = ((ta * x + tb) * x).sum((1, 2)) + tc + td
            proposal -= logk + ph
            logk_sum += logk
            phi_sum += ph
    nan = ~np.isfinite(U + proposal) if n_steps > 0 else np.zeros(P, dtype=bool)
    if nan.any():
        x[nan] = x0[nan]
    return dict(x=x, u_first=u_first, u_last=U.copy(), proposal_work=proposal, twist_logk=logk_sum, twist_phi=phi_sum, nan=nan)


# ------------------------------------------------------------------------------------------------ dense A_t on a molecular system
def dense_tables(A, b, c, d, s):
    """The per-iteration arrays of cdw_dense_twist (include/cdw.h) from A_t [D, D], b_t [D], c_t, d_t and the per-coordinate variances s [D] of
    the uncontrolled kernel N(f, diag s):  Sigma~ = (S^-1 + 2 A)^-1, theta = Sigma~ S^-1 (csmc.py:460-478 when S = dt I), the lower Cholesky
    factor of Sigma~, beta = Sigma~ b, logk0 = 0.5 log det theta + 0.5 b^T Sigma~ b - c - d, cd = c + d."""
    A, b, s = np.asarray(A, dtype=np.float64), np.asarray(b, dtype=np.float64), np.asarray(s, dtype=np.float64)
    sig = np.linalg.inv(np.diag(1.0 / s) + 2.0 * A)
    sig = 0.5 * (sig + sig.T)
    theta = sig / s[None, :]
    beta = sig @ b
    return dict(theta=theta, chol=np.linalg.cholesky(sig), A=A, b=b, beta=beta, logk0=float(0.5 * np.linalg.slogdet(theta)[1] + 0.5 * b @ beta - c - d),
                cd=float(c + d), sigma=sig)


def dense_log_normalizer_direct(f, A, b, c, d, s):
    """log int N(x'; f, diag s) exp(-phi(x')) dx' written as the reference writes it (csmc.py:531-561 with S in place of dt I): the difference of
    two large quadratic forms.  Used to check the cancellation-free form the kernel evaluates."""
    Sinv = np.diag(1.0 / np.asarray(s))
    sig = np.linalg.inv(Sinv + 2.0 * np.asarray(A))
    m = Sinv @ f - b
    return 0.5 * np.linalg.slogdet(sig @ Sinv)[1] + 0.5 * m @ sig @ m - 0.5 * f @ Sinv @ f - c - d


def twisted_ula_dense(energy_forces, x, masses, constraints, dt, kT, n_steps, seed, gids, step0, tables, tol=1e-6):
    """oracle.ula.ula with the dense twist `tables` = dense_tables(...) (the kernel's arithmetic: lane_twisted_move_dense, cdw_lane.cuh)."""
    x0 = r32(x)
    x = x0.copy()
    P, A = x.shape[0], x.shape[1]
    D = 3 * A
    masses = np.asarray(masses, dtype=np.float64)
    inv_m = 1.0 / masses
    tau = taus(masses, dt, kT)[None, :, None]
    drift_c = tau / kT
    th, ch, Am, b, beta = (np.asarray(tables[k], dtype=np.float64) for k in ("theta", "chol", "A", "b", "beta"))
    U, F = energy_forces(x)
    u_first = U.copy()
    proposal = np.zeros(P)
    for s_ in range(n_steps):
        xi = rng.normals(seed, gids, step0 + s_, A, KIND_ULA).reshape(P, D)
        x_old = x
        r0 = [(x[:, i].astype(np.float32) - x[:, j].astype(np.float32)).astype(np.float64) for (i, j, _) in constraints]
        f = (x_old + drift_c * F).reshape(P, D)
        u, w = f @ th.T, f @ Am.T
        logk = tables["logk0"] - (u * w).sum(1) - u @ b
        x = r32(((u - beta[None]) + xi @ ch.T).reshape(P, A, 3))
        if constraints:
            dummy = np.zeros_like(x)
            _shake(x, dummy, r0, inv_m, constraints, tol, 0.0)
        dlt = x - x_old - drift_c * F
        fwd = -((dlt * dlt) / (4.0 * tau)).sum((1, 2))
        U, F = energy_forces(x)
        dlt = x_old - x - drift_c * F
        bwd = -((dlt * dlt) / (4.0 * tau)).sum((1, 2))
        xf = x.reshape(P, D)
        ph = ((xf @ Am.T) * xf).sum(1) + xf @ b + tables["cd"]
        proposal += fwd - bwd - (logk + ph)
    nan = ~np.isfinite(U + proposal) if n_steps > 0 else np.zeros(P, dtype=bool)
    if nan.any():
        x[nan] = x0[nan]
    return dict(x=x, u_first=u_first, u_last=U.copy(), proposal_work=proposal, nan=nan)
