#!/usr/bin/env python
"""Generate tests/golden/csmc_ula.json by RUNNING the reference's toy ULA / SIS code (troubleshooting/csmc.py).

Build container only (needs /root/reference, scipy, torch for csmc.py's import).  What is recorded:

* `moves`: seeded single ULA moves through `uncontrolled_ULA_proposal_utility` (csmc.py:96-134) under the two toy
  potentials of troubleshooting/test_csmc.py:12-24, with the three log-weights of test_csmc.py:119-162
  (`compute_generalized_logw`, `compute_ULA_logw`, manual) as the reference computes them.
* `sis`: one seeded ensemble of `SIS(..., metropolize=False)` runs (csmc.py:197-262) on the non-default potential:
  final cumulative log-weights and the EXP free energy (test_csmc.py:176-226 expects -0.34 +- 0.1).
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
    np.random.seed(20260101)
    moves = []
    for name, pot, dim in (("default", default_potential, 1), ("nondefault", nondefault_potential, 1), ("default", default_potential, 3),
                           ("nondefault", nondefault_potential, 6)):
        for _ in range(6):
            dt = float(np.random.rand())
            p_old, p_new = float(np.random.rand()), float(np.random.rand())
            x = np.random.randn(dim) * np.sqrt(0.5)
            if dim == 1:
                x_f, lf, lb = csmc.uncontrolled_ULA_proposal_utility(x, pot, dt, p_new)
            else:  # the utility's `if not x_forward` only handles scalars: call its two halves the same way
                from scipy.stats import multivariate_normal
                mu, S = csmc.EL_mu_sigma(x, pot, dt, p_new)
                x_f = multivariate_normal.rvs(mean=mu, cov=S)
                lf = multivariate_normal.logpdf(x_f, mu, S)
                mu_b, S_b = csmc.EL_mu_sigma(x_f, pot, dt, p_new)
                lb = multivariate_normal.logpdf(x, mu_b, S_b)
            lg_old, lg_new = -pot(x, p_old), -pot(x_f, p_new)
            f_x, f_xf = csmc.compute_force(x, pot, p_new), csmc.compute_force(x_f, pot, p_new)
            moves.append(dict(potential=name, dim=dim, dt=dt, param_old=p_old, param_new=p_new, x=x.tolist(), x_forward=np.asarray(x_f).tolist(),
                              logp_forward=float(lf), logp_backward=float(lb), force_x=f_x.tolist(), force_x_forward=f_xf.tolist(),
                              logw_manual=float(lg_new + lb - lg_old - lf),
                              logw_generalized=float(csmc.compute_generalized_logw(lg_old, lg_new, lf, lb)),
                              logw_ula=float(csmc.compute_ULA_logw(x, np.asarray(x_f), f_x, f_xf, -lg_old, -lg_new, dt))))
    np.random.seed(7)
    lam = np.linspace(0, 1, 10)
    from scipy.stats import multivariate_normal
    logws = []
    for _ in range(400):
        x = np.array([multivariate_normal.rvs(0.0, 1.0)])
        lw, _, _ = csmc.SIS(x, nondefault_potential, lam, 1.0, metropolize=False)
        logws.append(float(np.sum(lw)))
    w = -np.array(logws)
    fe = float(-(np.logaddexp.reduce(-w) - np.log(len(w))))
    out = dict(source="troubleshooting/csmc.py + troubleshooting/test_csmc.py run from /root/reference", moves=moves,
               sis=dict(potential="nondefault", dt=1.0, n_lambda=10, n_particles=400, free_energy=fe, expected=-0.34, tolerance=0.1,
                        cumulative_logw_mean=float(np.mean(logws)), cumulative_logw_std=float(np.std(logws))))
    with open(os.path.join(HERE, "csmc_ula.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote csmc_ula.json; SIS free energy", fe)


if __name__ == "__main__":
    main()
