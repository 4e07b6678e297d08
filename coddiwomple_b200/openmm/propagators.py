"""OpenMM Propagator Adapter module — `OMMBIP` with the reference's `apply` signature
(coddiwomple/openmm/propagators.py:21-261), executed by the fused BAOAB kernel (cdw_baoab)."""
import ctypes as C
import logging

import numpy as np
import torch

from .. import _lib
from ..engine import positions_to_rows
from ..propagators import Propagator
from .states import OpenMMParticleState

_logger = logging.getLogger("openmm_propagators")


class OMMBIP(Propagator):
    """Generalized OpenMM Base Integrator Propagator.

    `apply` keeps the reference's semantics: the particle state is modified in place; `apply_pdf_to_context=True`
    pushes the pdf_state's current parameters into the propagation "context" (here: a device lambda vector owned by
    the propagator), False leaves the context's parameters as they were (tests/test_openmm.py:276-281);
    `returnable_key` selects which accumulated work is returned (default: proposal_work = 0.).
    Batched form: `apply_batch(pos_rows, vel_rows, ...)` on (P, A, 4) device tensors.
    """

    def __init__(self, openmm_pdf_state, integrator, context_cache=None, reassign_velocities=False, n_restart_attempts=0, seed=None):
        super().__init__()
        self.pdf_state = openmm_pdf_state
        self.integrator = integrator
        self.reassign_velocities = reassign_velocities
        self.n_restart_attempts = n_restart_attempts
        self.lib = openmm_pdf_state.lib
        # the propagation context's own copy of the global parameters (openmmtools ContextCache in the reference)
        self._context_lambda = openmm_pdf_state.lambda_device.clone()
        self._seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self._step_counter = 0
        self._gid_counter = 0
        self.integrator.reset()

    def _get_context_parameters(self):
        lam = self._context_lambda.cpu().numpy()
        return {n: float(lam[i]) for i, n in enumerate(self.pdf_state.spec.lambda_names)}

    def _get_global_integrator_variables(self):
        i = self.integrator
        return dict(a=i.a, b=i.b, shadow_work=i.shadow_work, proposal_work=i.proposal_work, kT=i.kT)

    def context_reduced_potential(self, particle_state):
        """beta*U in the propagation context (the reference reads it with context.getState(getEnergy=True))."""
        rows = positions_to_rows(particle_state.positions[None])
        u = torch.empty(1, dtype=torch.float64, device="cuda")
        _lib.check(self.lib.cdw_reduced_potential(self.pdf_state.device_system.handle, _lib.ptr(rows), 1, _lib.ptr(self._context_lambda),
                                                 self.pdf_state.beta, _lib.ptr(u), _lib.current_stream()), "cdw_reduced_potential")
        return float(u.item())

    def apply_batch(self, pos_rows, vel_rows, n_steps=1, reset_integrator=False, apply_pdf_to_context=False, randomize_velocities=False,
                    track_works=True, gid0=None):
        """In-place BAOAB on device rows.  Returns dict(u_first, u_last, shadow, proposal, nan) device tensors (kT)."""
        if reset_integrator:
            self.integrator.reset()
        if apply_pdf_to_context:
            self._context_lambda.copy_(self.pdf_state.lambda_device)
        P = pos_rows.shape[0]
        d = pos_rows.device
        out = dict(u_first=torch.empty(P, dtype=torch.float64, device=d), u_last=torch.empty(P, dtype=torch.float64, device=d),
                   shadow=torch.zeros(P, dtype=torch.float64, device=d), proposal=torch.zeros(P, dtype=torch.float64, device=d),
                   nan=torch.zeros(P, dtype=torch.int32, device=d))
        flags = (_lib.CDW_FLAG_RANDOMIZE_VELOCITIES if (randomize_velocities or self.reassign_velocities) else 0) | \
                (_lib.CDW_FLAG_TRACK_WORKS if track_works else 0)
        trials = torch.zeros(P, 2, dtype=torch.int32, device=d) if getattr(self.integrator, "splitting_word", 0) else None
        out["trials"] = trials
        prm = self.integrator.as_struct(flags, n_steps=n_steps, trial_counts=trials)
        if gid0 is None:
            gid0 = self._gid_counter
        _lib.check(self.lib.cdw_baoab(self.pdf_state.device_system.handle, _lib.ptr(pos_rows), _lib.ptr(vel_rows), P, _lib.ptr(self._context_lambda),
                                     C.byref(prm), C.c_uint64(self._seed), C.c_uint64(self._step_counter), C.c_int64(gid0), _lib.ptr(out["u_first"]),
                                     _lib.ptr(out["u_last"]), _lib.ptr(out["shadow"]), _lib.ptr(out["proposal"]), _lib.ptr(out["nan"]),
                                     _lib.current_stream()), "cdw_baoab")
        self._step_counter += max(int(n_steps), 1)
        # NaN restart loop (openmm/propagators.py:134-190): a replica whose move produced a non-finite energy was left at its
        # pre-move state by the kernel; it is integrated again (fresh noise: the step counter has moved on), up to
        # n_restart_attempts times.  Whatever is still flagged after the last attempt stays reverted, as in the reference.
        out["attempts"] = 0
        for attempt in range(int(self.n_restart_attempts)):
            bad = torch.nonzero(out["nan"]).flatten()
            if bad.numel() == 0:
                break
            _logger.warning(f"Potential energy is NaN after {attempt} attempts of integration with move {self.__class__.__name__} Attempting a restart...")
            sub_pos, sub_vel = pos_rows[bad].contiguous(), vel_rows[bad].contiguous()
            n = int(bad.numel())
            sub = {k: torch.zeros(n, dtype=out[k].dtype, device=d) for k in ("u_first", "u_last", "shadow", "proposal", "nan")}
            _lib.check(self.lib.cdw_baoab(self.pdf_state.device_system.handle, _lib.ptr(sub_pos), _lib.ptr(sub_vel), n, _lib.ptr(self._context_lambda),
                                         C.byref(prm), C.c_uint64(self._seed), C.c_uint64(self._step_counter), C.c_int64(gid0), _lib.ptr(sub["u_first"]),
                                         _lib.ptr(sub["u_last"]), _lib.ptr(sub["shadow"]), _lib.ptr(sub["proposal"]), _lib.ptr(sub["nan"]),
                                         _lib.current_stream()), "cdw_baoab")
            self._step_counter += max(int(n_steps), 1)
            pos_rows[bad] = sub_pos
            vel_rows[bad] = sub_vel
            for k in sub:
                out[k][bad] = sub[k]
            out["attempts"] = attempt + 1
        if trials is not None and hasattr(self.integrator, "record_trials"):
            self.integrator.record_trials(trials.cpu().numpy())
        return out

    def apply(self, particle_state, n_steps=1, reset_integrator=False, apply_pdf_to_context=False, returnable_key=None,
              randomize_velocities=False, **kwargs):
        assert isinstance(particle_state, OpenMMParticleState)
        pos = positions_to_rows(particle_state.positions[None])
        vel = positions_to_rows(particle_state.velocities[None]) if particle_state.velocities is not None else torch.zeros_like(pos)
        out = self.apply_batch(pos, vel, n_steps=n_steps, reset_integrator=reset_integrator, apply_pdf_to_context=apply_pdf_to_context,
                               randomize_velocities=randomize_velocities)
        if int(out["nan"].item()):
            _logger.warning(f"Potential energy is NaN after {out['attempts']} attempts of integration with move {self.__class__.__name__}")
        else:
            self.integrator.shadow_work += float(out["shadow"].item())
            self.integrator.proposal_work += float(out["proposal"].item())
        particle_state.positions = pos[0, :, :3].double().cpu().numpy()
        particle_state.velocities = vel[0, :, :3].double().cpu().numpy()
        if returnable_key is not None:
            try:
                proposal_work = getattr(self.integrator, f"get_{returnable_key}_work")(dimensionless=True)
            except Exception as e:
                raise Exception(f"{e}; the returnable key {returnable_key} is not a valid method in {self.integrator.__class__.__name__}")
        else:
            proposal_work = 0.
        return particle_state, proposal_work


class OMMAISP(OMMBIP):
    """Annealed Importance Sampling propagator — coddiwomple/openmm/propagators.py:264-344.

    `apply` anneals the FULL protocol (fractional_iteration 0 -> 1) for one particle state and records the state-work
    trajectory; `apply_batch` does the same for P particle states at once (one replica per team of GPU lanes) and
    returns the (P, n_records) work matrix in kT."""

    def __init__(self, openmm_pdf_state, integrator, record_state_work_interval=None, context_cache=None, reassign_velocities=False,
                 n_restart_attempts=0, seed=None):
        super().__init__(openmm_pdf_state, integrator, context_cache=context_cache, reassign_velocities=reassign_velocities,
                         n_restart_attempts=n_restart_attempts, seed=seed)
        pdf_state_parameters = list(self.pdf_state.get_parameters().keys())
        function_parameters = self.integrator._function_parameters
        assert set(pdf_state_parameters) == set(function_parameters), f"the pdf_state parameters ({pdf_state_parameters}) is not equal to the function parameters ({function_parameters})"
        self._record_state_work_interval = record_state_work_interval
        self._state_works = {}
        self._state_works_counter = 0

    @property
    def state_works(self):
        return self._state_works

    def apply_batch(self, positions, n_steps=None, seed=None):
        """positions: (P, A, 3) nm.  Returns (works[P][n_records] in kT, engine)."""
        from ..engine import SMCEngine
        n_steps = self.integrator._n_steps if n_steps is None else int(n_steps)
        if n_steps != self.integrator._n_steps:
            raise ValueError("n_steps must equal the integrator's protocol length")
        seq = self.integrator.parameter_sequence(self.pdf_state.get_parameters())
        P = np.asarray(positions).shape[0]
        interval = self._record_state_work_interval
        self._seed_counter = getattr(self, "_seed_counter", 0) + 1
        eng = SMCEngine(self.pdf_state.device_system, P, seq, langevin=self.integrator, resampler="multinomial", floor_threshold=0.0,
                        seed=(self._seed if seed is None else int(seed)) + 7919 * self._seed_counter, record_works_every=(interval or n_steps))
        eng.anneal(positions)
        works = eng.work_history()  # (n_records, P), first row = 0
        return works.T, eng

    def apply(self, particle_state, n_steps=1, reset_integrator=False, apply_pdf_to_context=False, returnable_key=None, randomize_velocities=False,
              **kwargs):
        if reset_integrator:
            self.integrator.reset()
        works, eng = self.apply_batch(particle_state.positions[None], n_steps=n_steps)
        self._current_state_works = [float(w) for w in works[0]]
        self._state_works[self._state_works_counter] = list(self._current_state_works)
        self._state_works_counter += 1
        self.integrator.state_work = self._current_state_works[-1]
        particle_state.positions = eng.ensemble.positions[0].double().cpu().numpy()
        particle_state.velocities = eng.ensemble.velocities[0].double().cpu().numpy()
        self.pdf_state.set_parameters({k: eng.parameter_sequence[-1][k] for k in self.pdf_state.get_parameters()})
        return particle_state, 0.


class OMMAISVP(OMMAISP):
    """Verbose variant in the reference (coddiwomple/openmm/propagators.py:348-391); same arithmetic."""


class OMMAISPR(OMMAISP):
    """Trajectory-writing variant in the reference (coddiwomple/openmm/propagators.py:394-478); writing is out of scope,
    the work bookkeeping is identical."""

    def __init__(self, openmm_pdf_state, integrator, record_state_work_interval=None, reporter=None, trajectory_write_interval=1, context_cache=None,
                 reassign_velocities=False, n_restart_attempts=0, seed=None):
        super().__init__(openmm_pdf_state, integrator, record_state_work_interval=record_state_work_interval, context_cache=context_cache,
                         reassign_velocities=reassign_velocities, n_restart_attempts=n_restart_attempts, seed=seed)
        self._reporter = reporter


class OMMEMP(OMMBIP):
    """OpenMM Euler-Maruyama Propagator: `apply` proposes x' ~ K(.|x) with the unadjusted Langevin kernel of the
    pdf_state's current target and returns the proposal work log K(x'|x) - log L(x|x')
    (coddiwomple/openmm/integrators.py:780-888; the generalised weight of troubleshooting/csmc.py:174-191 is then
    `TargetFactory.compute_incremental_work`'s u_t(x_t) + aux + proposal_work, distribution_factories.py:98-131)."""

    def apply_batch(self, pos_rows, vel_rows=None, n_steps=1, reset_integrator=False, apply_pdf_to_context=False, gid0=None, **kwargs):
        if reset_integrator:
            self.integrator.reset()
        if apply_pdf_to_context:
            self._context_lambda.copy_(self.pdf_state.lambda_device)
        P = pos_rows.shape[0]
        d = pos_rows.device
        out = dict(u_first=torch.empty(P, dtype=torch.float64, device=d), u_last=torch.empty(P, dtype=torch.float64, device=d),
                   proposal=torch.zeros(P, dtype=torch.float64, device=d), nan=torch.zeros(P, dtype=torch.int32, device=d))
        prm = self.integrator.as_struct(_lib.CDW_FLAG_ULA, n_steps=n_steps)
        if gid0 is None:
            gid0 = self._gid_counter
        _lib.check(self.lib.cdw_smc_propagate(self.pdf_state.device_system.handle, _lib.ptr(pos_rows), None, _lib.ptr(pos_rows), None, P, None, None, None,
                                             None, _lib.ptr(self._context_lambda), C.byref(prm), C.c_uint64(self._seed), C.c_uint64(self._step_counter),
                                             C.c_int64(gid0), None, None, None, None, _lib.ptr(out["u_first"]), _lib.ptr(out["u_last"]), None,
                                             _lib.ptr(out["proposal"]), _lib.ptr(out["nan"]), None, _lib.current_stream()), "cdw_smc_propagate")
        self._step_counter += max(int(n_steps), 1)
        return out

    def apply(self, particle_state, n_steps=1, reset_integrator=False, apply_pdf_to_context=False, **kwargs):
        assert isinstance(particle_state, OpenMMParticleState)
        pos = positions_to_rows(particle_state.positions[None])
        out = self.apply_batch(pos, n_steps=n_steps, reset_integrator=reset_integrator, apply_pdf_to_context=apply_pdf_to_context)
        work = 0.0
        if not int(out["nan"].item()):
            work = float(out["proposal"].item())
            self.integrator.proposal_work += work
        particle_state.positions = pos[0, :, :3].double().cpu().numpy()
        return particle_state, work
