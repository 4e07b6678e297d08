"""Controlled-SMC helpers on the host — the initial step of `twisted_smc` (troubleshooting/csmc.py:625-747).

The per-iteration twists run inside the propagate kernel (`engine.TwistSequence`, include/cdw.h).  The ZEROTH step is different: when the first
target pi_0 is a Gaussian mixture, the initial twist psi_0 = exp(-(x^T A_0 x + b_0^T x + c_0)) turns it into another Gaussian mixture the particles
are drawn from exactly, with the initial log-weight log w_0 = log Z_0 + phi_0(x_0) (csmc.py:336-458, 564-591).  That is a one-off draw of P
vectors, done here in numpy exactly like the reference draws its initial frames on the host (openmm/coddiwomple.py:302-304); the works go to
`SMCEngine.initialize(..., initial_works=-log_w0)`."""
import numpy as np


def twisted_gmm_components(alphas, mus, Sigmas, A0, b0, c0):
    """Unnormalised log mixing weights log(alpha_j zeta_j) and covariances (Sigma_j^-1 + 2 A_0)^-1 of the twisted mixture (csmc.py:336-374, 432-458)."""
    alphas, mus, Sigmas = np.asarray(alphas, dtype=np.float64), np.asarray(mus, dtype=np.float64), np.asarray(Sigmas, dtype=np.float64)
    A0, b0 = np.asarray(A0, dtype=np.float64), np.asarray(b0, dtype=np.float64)
    if mus.shape != (alphas.shape[0], b0.shape[0]) or Sigmas.shape != (alphas.shape[0], b0.shape[0], b0.shape[0]):
        raise ValueError("alphas (M,), mus (M, N), Sigmas (M, N, N), A0 (N, N), b0 (N,)")
    tilde = np.linalg.inv(np.linalg.inv(Sigmas) + 2.0 * A0)
    log_zeta = np.empty(alphas.shape[0])
    for j in range(alphas.shape[0]):
        Si = np.linalg.inv(Sigmas[j])
        m = Si @ mus[j] - b0
        log_zeta[j] = (-0.5 * np.linalg.slogdet(Sigmas[j])[1] + 0.5 * np.linalg.slogdet(tilde[j])[1] + 0.5 * m @ tilde[j] @ m
                       - 0.5 * mus[j] @ Si @ mus[j] - c0)
    return np.log(alphas) + log_zeta, tilde


def twisted_gmm_initial_sample(n, alphas, mus, Sigmas, A0, b0, c0=0.0, rng=None):
    """n draws of the twisted Gaussian-mixture initial proposal and their initial log-weights (csmc.py:376-430, 564-591).

    Returns (x0 (n, N), log_w0 (n,)): log w_0 = logsumexp_j log(alpha_j zeta_j) + x0^T A_0 x0 + b_0^T x0 + c_0.  With A_0 = 0, b_0 = 0, c_0 = 0 the
    draws are plain samples of the mixture and log w_0 = 0 (the reference's consistency check, csmc.py:585-590)."""
    rng = np.random.default_rng() if rng is None else rng
    log_at, tilde = twisted_gmm_components(alphas, mus, Sigmas, A0, b0, c0)
    log_norm = np.logaddexp.reduce(log_at)
    comp = rng.choice(len(log_at), size=int(n), p=np.exp(log_at - log_norm))
    mus, Sigmas, A0, b0 = (np.asarray(v, dtype=np.float64) for v in (mus, Sigmas, A0, b0))
    x0 = np.empty((int(n), b0.shape[0]))
    for j in range(len(log_at)):
        idx = np.nonzero(comp == j)[0]
        if idx.size:
            mean = tilde[j] @ (np.linalg.inv(Sigmas[j]) @ mus[j] - b0)
            x0[idx] = mean + rng.standard_normal((idx.size, b0.shape[0])) @ np.linalg.cholesky(tilde[j]).T
    log_w0 = log_norm + np.einsum("ni,ij,nj->n", x0, A0, x0) + x0 @ b0 + c0
    return x0, log_w0
