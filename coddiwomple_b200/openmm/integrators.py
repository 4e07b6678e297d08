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
