#!/usr/bin/env python
"""Generate tests/golden/csmc_twisted.json by RUNNING the twisting functions of the reference's controlled-SMC script
(troubleshooting/csmc.py:336-747).  Build container only (needs /root/reference, scipy, torch for csmc.py's import).

* `forward`: seeded calls of Theta_t, f_t, twisted_forward_proposal, twisted_forward_log_normalizer, twisted_t_log_weight with DENSE
  non-zero A, b, c in 1, 2 and 3 dimensions (the reference's values, including what csmc.py:556 literally computes).
* `gmm`: twisted_gmm_components, compute_twisted_gmm_lognormalizer, twisted_gmm_proposal, twisted_zeroth_log_weight.
* `runs`: twisted_smc (csmc.py:625-747) on the 1-D non-default potential of troubleshooting/test_csmc.py:18-24 — one run with zero twists
  (the uncontrolled regime of test_csmc.py:165-226) and one with constant non-zero twists: trajectory and incremental log-weights.
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/troubleshooting"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import csmc  # noqa: E402
from test_csmc import default_potential, nondefault_potential  # noqa: E402


def main():
    rs = np.random.RandomState(20261020)
    forward = []
    for dim in (1, 1, 2, 3, 3):
        for pot_name, pot in (("default", default_potential), ("nondefault", nondefault_potential)):
            Mx = rs.randn(dim, dim)
            A = 0.4 * Mx @ Mx.T + 0.1 * np.eye(dim)
            b = rs.randn(dim)
            c, d = float(rs.randn()), 0.0
            dt = float(0.1 + 0.8 * rs.rand())
            lam = float(rs.rand())
            x = rs.randn(dim) * 0.7
            theta = csmc.Theta_t(x, A, dt)
            f = csmc.f_t(x, pot, lam, dt)
            np.random.seed(int(rs.randint(1 << 30)))
            x_fwd, logp_fwd = csmc.twisted_forward_proposal(theta, f, dt, b)
            logk = csmc.twisted_forward_log_normalizer(theta, f, b, c, d, dt)
            unc = float(rs.randn())
            logw = csmc.twisted_t_log_weight(x_fwd, unc, logk, A, b, c, d)
            forward.append(dict(dim=dim, potential=pot_name, A=A.tolist(), b=b.tolist(), c=c, d=d, dt=dt, parameter=lam, x=x.tolist(), theta=theta.tolist(),
                                f=np.asarray(f).tolist(), x_forward=np.asarray(x_fwd).tolist(), logp_forward=float(logp_fwd), log_forward_normalizer=float(logk),
                                uncontrolled_logw=unc, twisted_logw=float(logw)))
    gmm = []
    for dim, comps in ((1, 2), (2, 3), (3, 2)):
        alphas = rs.rand(comps); alphas /= alphas.sum()
        mus = rs.randn(comps, dim)
        Sig = np.array([(lambda m: 0.5 * m @ m.T + 0.3 * np.eye(dim))(rs.randn(dim, dim)) for _ in range(comps)])
        Mx = rs.randn(dim, dim)
        A0 = 0.3 * Mx @ Mx.T
        b0 = rs.randn(dim)
        c0 = float(rs.randn())
        log_alpha_tildes, Sig_tilde = csmc.twisted_gmm_components(alphas, mus, Sig, A0, b0, c0)
        lognorm = csmc.compute_twisted_gmm_lognormalizer(log_alpha_tildes)
        np.random.seed(int(rs.randint(1 << 30)))
        x0, logpdf = csmc.twisted_gmm_proposal(log_alpha_tildes, Sig_tilde, mus, Sig, b0)
        logw0 = csmc.twisted_zeroth_log_weight(x0, A0, b0, c0, lognorm)
        gmm.append(dict(dim=dim, alphas=alphas.tolist(), mus=mus.tolist(), Sigmas=Sig.tolist(), A0=A0.tolist(), b0=b0.tolist(), c0=c0,
                        log_alpha_tildes=np.asarray(log_alpha_tildes).tolist(), Sigma_tildes=np.asarray(Sig_tilde).tolist(), lognormalizer=float(lognorm),
                        x0=np.asarray(x0).tolist(), logpdf=float(logpdf), log_w0=float(logw0)))
    runs = []
    lam_seq = np.linspace(0, 1, 10)
    gmm_unc = dict(alphas=np.array([0.5, 0.5]), mus=np.array([[0.0], [0.0]]), Sigmas=np.array([[[1.0]], [[1.0]]]))
    for name, A_t, b_t, c_t in (("zero", 0.0, 0.0, 0.0), ("constant", 0.35, -0.2, 0.1)):
        T = len(lam_seq) - 1
        fns = dict(A=[(lambda x, v=A_t: np.array([[v]])) for _ in range(T)], b=[(lambda x, v=b_t: np.array([v])) for _ in range(T)],
                   c=[(lambda v=c_t: v) for _ in range(T)])
        init = dict(A=np.array([[0.0]]), b=np.array([0.0]), c=0.0)
        np.random.seed(11 if name == "zero" else 12)
        traj, logw = csmc.twisted_smc(nondefault_potential, lam_seq, 0.5, init, fns, None, gmm_unc)
        runs.append(dict(name=name, A=A_t, b=b_t, c=c_t, dt=0.5, parameter_sequence=lam_seq.tolist(), trajectory=[np.asarray(x).tolist() for x in traj],
                         log_incremental_weights=[float(np.asarray(w).reshape(-1)[0]) for w in logw]))
    out = dict(source="troubleshooting/csmc.py:336-747 run from /root/reference", forward=forward, gmm=gmm, runs=runs)
    with open(os.path.join(HERE, "csmc_twisted.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote csmc_twisted.json:", len(forward), "forward cases,", len(gmm), "gmm cases,", len(runs), "runs")


if __name__ == "__main__":
    main()
