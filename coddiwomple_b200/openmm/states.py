"""OpenMM State Adapter module — `OpenMMParticleState`, `OpenMMPDFState` with the reference's interface
(coddiwomple/openmm/states.py:25-136), evaluated by the K1 kernel instead of an OpenMM Context."""
import numpy as np
import torch

from .. import _lib
from ..engine import DeviceSystem, as_spec, positions_to_rows
from ..states import ParticleState, PDFState
from ..system import kT_in_md_units
from ..testsystems import HO_PREFIX, RELATIVE_ALCHEMICAL_PARAMETERS
from . import units


class AlchemicalState:
    """Stand-in for openmmtools.alchemy.AlchemicalState's role here: names the global parameters a PDFState exposes."""
    parameter_names = ("lambda_sterics", "lambda_electrostatics", "lambda_bonds", "lambda_angles", "lambda_torsions")

    @classmethod
    def exposed(cls, spec):
        return [n for n in cls.parameter_names if n in spec.lambda_names]


class RelativeAlchemicalState(AlchemicalState):
    """perses.annihilation.lambda_protocol.RelativeAlchemicalState's nine parameters (openmm/coddiwomple.py:24)."""
    parameter_names = RELATIVE_ALCHEMICAL_PARAMETERS


class HarmonicAlchemicalState(AlchemicalState):
    """coddiwomple/tests/utils.py:10-24."""
    parameter_names = (f"{HO_PREFIX}_x0", f"{HO_PREFIX}_U0")


class OpenMMParticleState(ParticleState):
    """positions (A,3) nm, velocities (A,3) nm/ps, box vectors (unused in vacuum) — openmm/states.py:25-41."""

    def __init__(self, positions, velocities=None, box_vectors=None, **kwargs):
        self.positions = np.array(units.strip(positions), dtype=np.float64)
        v = units.strip(velocities)
        self.velocities = None if v is None else np.array(v, dtype=np.float64)
        self.box_vectors = units.strip(box_vectors)


class OpenMMPDFState(PDFState):
    """beta*U(x; lambda) of an OpenMM System, on the GPU (openmm/states.py:45-136).

    system: OpenMM System XML text, a SystemSpec, or an openmm.System when OpenMM is installed.
    """

    def __init__(self, system, alchemical_composability=AlchemicalState, temperature=300.0, pressure=1.0, **kwargs):
        self.device_system = system if isinstance(system, DeviceSystem) else DeviceSystem(as_spec(system))
        self.spec = self.device_system.spec
        self.temperature = float(units.strip(temperature))
        self.pressure = units.strip(pressure)
        if self.pressure is not None:
            raise NotImplementedError("only the vacuum path (pressure=None, openmm/coddiwomple.py:124-143) is in scope")
        self.kT = kT_in_md_units(self.temperature)
        self.beta = 1.0 / self.kT
        self.is_periodic = False
        self._exposed = alchemical_composability.exposed(self.spec)
        self._parameters = {n: float(self.spec.lambda_defaults[self.spec.lambda_names.index(n)]) for n in self._exposed}
        self._lam = torch.from_numpy(self.spec.lambda_vector(self._parameters)).cuda()
        self.lib = self.device_system.lib

    def set_parameters(self, parameters):
        assert set(self._parameters.keys()) == set(parameters.keys()), f"the parameter keys supplied do not match the internal parameter names"
        for key, val in parameters.items():
            self._parameters[key] = float(val)
        self._lam.copy_(torch.from_numpy(self.spec.lambda_vector(self._parameters)))

    def get_parameters(self):
        return dict(self._parameters)

    @property
    def lambda_device(self):
        return self._lam

    def reduced_potential(self, particle_state):
        """Single state -> float, as in the reference; a (P, A, 3) array or float4-row tensor -> array of P values."""
        if isinstance(particle_state, OpenMMParticleState):
            return float(self.reduced_potential_batch(particle_state.positions[None])[0])
        out = self.reduced_potential_batch(particle_state)
        return out

    def reduced_potential_batch(self, positions):
        rows = positions if torch.is_tensor(positions) else positions_to_rows(positions)
        P = rows.shape[0]
        u = torch.empty(P, dtype=torch.float64, device=rows.device)
        _lib.check(self.lib.cdw_reduced_potential(self.device_system.handle, _lib.ptr(rows), P, _lib.ptr(self._lam), self.beta, _lib.ptr(u),
                                                 _lib.current_stream()), "cdw_reduced_potential")
        return u if torch.is_tensor(positions) else u.cpu().numpy()
