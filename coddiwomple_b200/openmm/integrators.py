"""OpenMM Integrator module — `OMMLI` with the reference's constructor (coddiwomple/openmm/integrators.py:19-465).

The reference builds a CustomIntegrator program from the splitting string; here the "V R O R V" (BAOAB) program is
the fused K4 kernel and any other string of V, R, O sub-steps with an optional Metropolised block `{ }` runs the kernel's
programmable body, so the class only carries the numbers and the work accumulators.  Multiple-time-step splittings are
refused (the reference itself raises on MTS, :412).
"""
import numpy as np

from ..engine import LangevinParameters
from . import units


class OMMLI(LangevinParameters):
    """Langevin integrator from a splitting string of V, R, O sub-steps and at most one Metropolised block "{ ... }" that must hold
    every R and V (openmm/integrators.py:207-276, 363-444).  "V R O R V" is the fused BAOAB body; other strings run the programmable
    body (cdw_langevin_params.splitting).  Force-group (multiple-time-step) V steps are refused like the reference does (:412)."""

    def __init__(self, temperature=300.0, collision_rate=1.0, timestep=0.001, splitting="V R O R V", constraint_tolerance=1e-6, **kwargs):
        splitting = " ".join(splitting.upper().split())
        self._sanity_check(splitting)
        super().__init__(temperature=units.strip(temperature), collision_rate=units.strip(collision_rate), timestep=units.strip(timestep),
                         constraint_tolerance=constraint_tolerance, splitting=splitting)
        self._splitting = splitting
        self._metropolized_integrator = "{" in splitting
        self.ntrials = self.naccept = self.nreject = 0
        self.reset()

    @staticmethod
    def _sanity_check(splitting):
        """openmm/integrators.py:219-276: only R, V, O, { and }; one un-nested block; no shadow-work generating step outside it."""
        toks = splitting.split()
        for tok in toks:
            if tok[0] == "V" and len(tok) > 1:
                raise NotImplementedError("force group splitting is currently disabled")   # as the reference (integrators.py:412)
            if tok not in ("V", "R", "O", "{", "}"):
                raise ValueError(f"Invalid step name '{tok}' used; valid step names are R V O {{ }}")
        if "{" in toks or "}" in toks:
            if toks.count("{") != 1 or toks.count("}") != 1:
                raise ValueError("There can only be one Metropolized region.")
            lo, hi = toks.index("{"), toks.index("}")
            if lo > hi or any(t in ("R", "V") for t in toks[:lo] + toks[hi + 1:]):
                raise ValueError("Shadow work generating steps found outside the Metropolization block")

    def record_trials(self, trial_counts):
        """Add the (trials, accepted) pairs a launch returned to the integrator's counters (the reference's ntrials / naccept /
        nreject globals, integrators.py:432-443)."""
        t, a = int(trial_counts[..., 0].sum()), int(trial_counts[..., 1].sum())
        self.ntrials += t
        self.naccept += a
        self.nreject += t - a

    @property
    def acceptance_rate(self):
        return self.naccept / self.ntrials if self.ntrials else float("nan")

    # work accumulators, in kT (integrators.py:160-161, 200-205)
    def reset_proposal_work(self):
        self.proposal_work = 0.0

    def reset_shadow_work(self):
        self.shadow_work = 0.0

    def reset(self):
        self.reset_proposal_work()
        self.reset_shadow_work()

    @property
    def a(self):
        return float(np.exp(-self.collision_rate * self.timestep))

    @property
    def b(self):
        return float(np.sqrt(1.0 - np.exp(-2.0 * self.collision_rate * self.timestep)))

    def _get_energy_with_units(self, variable_name, dimensionless=False):
        work = getattr(self, variable_name)
        return work if dimensionless else work * self.kT

    def get_proposal_work(self, dimensionless=True):
        return self._get_energy_with_units("proposal_work", dimensionless)

    def get_shadow_work(self, dimensionless=True):
        return self._get_energy_with_units("shadow_work", dimensionless)


class OMMLIAIS(OMMLI):
    """Annealed-importance-sampling Langevin integrator — coddiwomple/openmm/integrators.py:468-665.

    Splitting "P I H N S V R O R V": before every BAOAB step the protocol advances (I: iteration += 1,
    fractional_iteration = iteration / n_steps; H: lambda <- f(fractional_iteration)) and the state work accumulates
    new_pe - old_pe at fixed coordinates (P, N, S).  That is exactly the incremental work the fused propagate kernel
    computes (inc = u_k(x_{k-1}) - u_{k-1}(x_{k-1})), so AIS is the SMC loop with resampling switched off; the Lepton
    update functions are lowered on the host to a [n_steps + 1][n_lambda] table (openmm/lepton.py).
    """

    def __init__(self, openmm_pdf_update_functions, n_steps, temperature=300.0, collision_rate=1.0, timestep=0.001, splitting="P I H N S V R O R V",
                 constraint_tolerance=1e-6, **kwargs):
        if openmm_pdf_update_functions is None:
            raise Exception(f"the {self.__class__.__name__} class requires update functions")
        if (n_steps < 0) or (n_steps != int(n_steps)):
            raise Exception('n_steps must be an integer >= 0')
        if " ".join(splitting.upper().split()) != "P I H N S V R O R V":
            raise NotImplementedError(f"splitting {splitting!r}: only 'P I H N S V R O R V' is implemented")
        self._openmm_pdf_update_functions = dict(openmm_pdf_update_functions)
        self._n_steps = int(n_steps)
        self._function_parameters = list(self._openmm_pdf_update_functions.keys())
        super().__init__(temperature=temperature, collision_rate=collision_rate, timestep=timestep, splitting="V R O R V", constraint_tolerance=constraint_tolerance)
        self._splitting = splitting

    def parameter_sequence(self, start_parameters):
        """lambda_0 = the pdf_state's current parameters (the driver resets them to the end state), then f(k / n_steps)."""
        from .lepton import protocol_table
        seq = [dict(start_parameters)]
        for row in protocol_table(self._openmm_pdf_update_functions, self._n_steps):
            full = dict(start_parameters)
            full.update(row)
            seq.append(full)
        return seq

    def reset(self):
        super().reset()
        self.state_work = 0.0
        self.iteration = 0.0

    def get_state_work(self, dimensionless=True):
        return self._get_energy_with_units("state_work", dimensionless)

    def set_state_work(self, work):
        self.state_work = float(work)

    def set_proposal_work(self, work):
        self.proposal_work = float(work)


class OpenMMEulerMaruyamaIntegrator(LangevinParameters):
    """Generalized Euler-Maruyama (unadjusted Langevin) integrator — coddiwomple/openmm/integrators.py:666-888.

        x_(k+1) = x_k + tau * force(x_k | parameters) * beta + sqrt(2 tau) * R,   tau = D dt^2,  D = kT / (2 m)   (:676-680, 829-838)
        proposal_work = log q(x_new | x_old) - log q(x_old | x_new)                                                (:864-888)

    followed, when the system has constraints, by the position-constraint projection of the CustomIntegrator flavour
    (:741-745).  The move and both kernel densities are computed by the fused kernel (CDW_FLAG_ULA); this class
    carries the numbers.  There is no collision rate and there are no velocities.
    """

    def __init__(self, timestep=0.001, temperature=300.0, constraint_tolerance=1e-6, mass_vector=None, **kwargs):
        super().__init__(temperature=units.strip(temperature), collision_rate=0.0, timestep=units.strip(timestep), constraint_tolerance=constraint_tolerance)
        self.mass_vector = None if mass_vector is None else np.asarray(units.strip(mass_vector), dtype=np.float64).reshape(-1)
        self.reset()

    def reset(self):
        self.proposal_work = 0.0

    def tau(self, masses=None):
        """tau_a = kT dt^2 / (2 m_a)  [nm^2]."""
        m = self.mass_vector if masses is None else np.asarray(masses, dtype=np.float64)
        return self.kT * self.timestep ** 2 / (2.0 * m)

    def get_proposal_work(self, dimensionless=True):
        return self.proposal_work if dimensionless else self.proposal_work * self.kT
