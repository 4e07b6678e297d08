"""OpenMM Integrator module — `OMMLI` with the reference's constructor (coddiwomple/openmm/integrators.py:19-465).

The reference builds a CustomIntegrator program from the splitting string; here the "V R O R V" (BAOAB) program is
the fused K4 kernel, so the class only carries the numbers and the work accumulators.  Other splittings
(Metropolised `{ }`, multiple-time-step, g-BAOAB) are out of scope (the reference itself raises on MTS, :412).
"""
import numpy as np

from ..engine import LangevinParameters
from . import units


class OMMLI(LangevinParameters):
    def __init__(self, temperature=300.0, collision_rate=1.0, timestep=0.001, splitting="V R O R V", constraint_tolerance=1e-6, **kwargs):
        if " ".join(splitting.upper().split()) != "V R O R V":
            raise NotImplementedError(f"splitting {splitting!r}: only 'V R O R V' (BAOAB) is implemented by the fused kernel")
        super().__init__(temperature=units.strip(temperature), collision_rate=units.strip(collision_rate), timestep=units.strip(timestep),
                         constraint_tolerance=constraint_tolerance)
        self._splitting = splitting
        self._metropolized_integrator = False
        self.reset()

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
