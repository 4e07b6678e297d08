"""Device-resident SMC state and the fused annealing loop.

This is the B200 replacement for the body of the reference's hot loop
(`mcmc_smc_resampler`, coddiwomple/openmm/coddiwomple.py:312-331): where the reference walks a Python list of
`Particle` objects and calls OpenMM once per particle, all particles live on the GPU as struct-of-arrays and every
iteration is a fixed set of kernel launches with no host synchronisation (the resampling decision nESS <= threshold is
taken on the device): for MCMC moves ONE propagate launch on the main stream (cdw_smc_lookahead) with the weight
statistics / resampling of the same iteration beside it on a side stream; for the Euler-Maruyama proposal (optionally
twisted: TwistSequence) propagate and resample back to back.  PyTorch is used for allocation, streams and
`torch.distributed` only; all arithmetic is in libcdw_b200.so (include/cdw.h).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .system import SystemSpec, kT_in_md_units, spec_from_xml


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.CdwError("coddiwomple_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class DeviceSystem:
    """One replica's force field on the device (cdw_system handle). Replaces the OpenMM Context of
    OpenMMPDFState (coddiwomple/openmm/states.py:84-86)."""

    def __init__(self, system):
        _require_cuda()
        if isinstance(system, DeviceSystem):
            raise TypeError("already a DeviceSystem")
        self.spec = as_spec(system)
        self.lib = _lib.load()
        desc, keep = _lib.make_desc(self.spec)
        self.handle = C.c_void_p()
        _lib.check(self.lib.cdw_system_create(C.byref(desc), C.byref(self.handle)), "cdw_system_create")
        del keep
        self.n_atoms = self.spec.n_atoms
        self.n_lambda = self.spec.n_lambda

    def choose_team(self, n_particles_total):
        """Pin the lanes per replica from the size of the WHOLE ensemble (all ranks of a sharded run pass the same number, so a sharded
        run and a single-GPU run of the same ensemble still agree bit for bit).  Ensembles that fill the GPU several times over run
        teams of 2 when seven 64-thread CTAs of them fit an SM (measured, fluorobenzene hybrid, 1e5 replicas: 271 us per iteration
        against 304 us with teams of 4; the 22-atom system only fits three such CTAs and stays at 4: 745 against 984 us)."""
        team = 0
        if self.n_atoms >= 8 and int(n_particles_total) >= 65536 and 7 * int(self.lib.cdw_system_smem_bytes(self.handle, 2)) <= 227 * 1024:
            team = 2
        _lib.check(self.lib.cdw_system_set_team(self.handle, team), "cdw_system_set_team")
        return int(self.lib.cdw_system_team(self.handle))

    def lambda_table(self, parameter_sequence, device="cuda"):
        """[T][n_lambda] fp64 device table for a list of {name: value} dicts (reference: parameter_sequence)."""
        rows = [self.spec.lambda_vector(p) for p in parameter_sequence]
        return torch.from_numpy(np.stack(rows)).to(device)

    def register_protocol(self, lambda_table, kT):
        """Precompute the effective term tables of every row of `lambda_table` ([T][n_lambda] fp64 device tensor):
        propagate calls whose lambda pointer is a row of it then skip the per-CTA table build
        (cdw_system_register_protocol).  The tensor is kept alive here and must not be modified afterwards."""
        if self.n_lambda == 0:
            return
        self._protocol = lambda_table
        _lib.check(self.lib.cdw_system_register_protocol(self.handle, _lib.ptr(lambda_table), int(lambda_table.shape[0]), float(kT),
                                                        _lib.current_stream()), "cdw_system_register_protocol")

    def __del__(self):
        try:
            if self.handle:
                self.lib.cdw_system_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def as_spec(system):
    """Accept a SystemSpec, an OpenMM System XML string, or (when OpenMM is installed) an openmm.System."""
    if isinstance(system, SystemSpec):
        return system
    if isinstance(system, str):
        return spec_from_xml(system)
    if hasattr(system, "spec") and isinstance(system.spec, SystemSpec):
        return system.spec
    try:  # an openmm.System object: serialise it (cannot be exercised where OpenMM is absent)
        try:
            from openmm import XmlSerializer
        except ImportError:
            from simtk.openmm import XmlSerializer
        return spec_from_xml(XmlSerializer.serialize(system))
    except ImportError as e:
        raise TypeError(f"cannot interpret {type(system)} as a system: pass OpenMM System XML or a SystemSpec") from e


def positions_to_rows(positions, device="cuda", pinned=False):
    """(P, A, 3) host array (nm) -> (P, A, 4) float32 rows."""
    x = np.asarray(positions, dtype=np.float32)
    if x.ndim == 2:
        x = x[None]
    rows = torch.zeros(x.shape[0], x.shape[1], 4, dtype=torch.float32, pin_memory=pinned)
    rows[..., :3] = torch.from_numpy(np.ascontiguousarray(x))
    return rows.to(device, non_blocking=pinned) if device is not None else rows


class LangevinParameters:
    """OMMLI's numbers in MD units (coddiwomple/openmm/integrators.py:75-131, 147-169)."""

    STEP_CODES = {"V": _lib.CDW_STEP_V, "R": _lib.CDW_STEP_R, "O": _lib.CDW_STEP_O, "{": _lib.CDW_STEP_OPEN, "}": _lib.CDW_STEP_CLOSE}

    def __init__(self, temperature=300.0, collision_rate=1.0, timestep=0.002, constraint_tolerance=1e-6, n_steps=1, splitting="V R O R V"):
        self.temperature = float(temperature)
        self.collision_rate = float(collision_rate)
        self.timestep = float(timestep)
        self.constraint_tolerance = float(constraint_tolerance)
        self.n_steps = int(n_steps)
        self.splitting_word = self.encode_splitting(splitting)

    @classmethod
    def encode_splitting(cls, splitting):
        """The splitting string as cdw_langevin_params.splitting (4 bits per sub-step); 0 for "V R O R V", which has its own fused body."""
        toks = splitting.upper().split()
        if toks == ["V", "R", "O", "R", "V"]:
            return 0
        if len(toks) > 15:
            raise ValueError("a splitting may hold at most 15 sub-steps")
        word = 0
        for k, tok in enumerate(toks):
            word |= cls.STEP_CODES[tok] << (4 * k)
        return word

    @property
    def kT(self):
        return kT_in_md_units(self.temperature)

    def as_struct(self, flags=0, n_steps=None, trial_counts=None, twist=None):
        """`twist`: this iteration's row of a TwistSequence table (device tensor of 6 A + 2 doubles) or None."""
        dense = twist if isinstance(twist, _lib.DenseTwist) else None
        prm = _lib.LangevinParams(self.timestep, self.collision_rate, self.kT, self.constraint_tolerance,
                                  self.n_steps if n_steps is None else int(n_steps), int(flags), int(getattr(self, "splitting_word", 0)),
                                  None if trial_counts is None else trial_counts.data_ptr(),
                                  None if (twist is None or dense is not None) else twist.data_ptr(), None if dense is None else C.addressof(dense))
        prm._keep = dense   # the descriptor is read when the call is made
        return prm


class TwistSequence:
    """Per-iteration quadratic twists psi_t = exp(-(x^T A_t x + b_t^T x + c_t + d_t)) of the Euler-Maruyama proposal — the `twisting_functions`
    / `twisting_parameters` of troubleshooting/csmc.py:625-747 for position-independent A_t, b_t.

    Diagonal A_t:  `TwistSequence(a, b, c, d)` with a, b of shape (T-1, A, 3) (1/nm^2, 1/nm); row t - 1 of the device table holds
                   a[3A] b[3A] c d of iteration t (cdw_langevin_params.twist).
    Full A_t:      `TwistSequence.dense(A, b, c, d)` with A of shape (T-1, 3A, 3A) (symmetric), b (T-1, 3A); the per-iteration matrices of
                   cdw_dense_twist (Theta, the Cholesky factor of the proposal covariance, ...) are computed on the host when the engine
                   binds the sequence to a system and a time step.
    reference_normalizer=True (diagonal form only) evaluates the forward log normaliser the way csmc.py:556 literally does (include/cdw.h)."""

    def __init__(self, a, b, c=None, d=None, reference_normalizer=False, device=None):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        if a.ndim != 3 or a.shape[2] != 3 or a.shape != b.shape:
            raise ValueError("a and b must both have shape (n_iterations, n_atoms, 3)")
        n = a.shape[0]
        c = np.zeros(n) if c is None else np.asarray(c, dtype=np.float64).reshape(n)
        d = np.zeros(n) if d is None else np.asarray(d, dtype=np.float64).reshape(n)
        self.n_iterations, self.n_atoms = n, a.shape[1]
        self.reference_normalizer = bool(reference_normalizer)
        self.is_dense = False
        self.host = np.concatenate([a.reshape(n, -1), b.reshape(n, -1), c[:, None], d[:, None]], axis=1)
        self.table = None if device is None else torch.from_numpy(self.host).to(device)

    @classmethod
    def dense(cls, A, b, c=None, d=None):
        A, b = np.asarray(A, dtype=np.float64), np.asarray(b, dtype=np.float64)
        if A.ndim != 3 or A.shape[1] != A.shape[2] or A.shape[1] % 3 or b.shape != A.shape[:2]:
            raise ValueError("A must have shape (n_iterations, 3 n_atoms, 3 n_atoms) and b (n_iterations, 3 n_atoms)")
        if not np.allclose(A, np.swapaxes(A, 1, 2)):
            raise ValueError("A_t must be symmetric")
        self = cls.__new__(cls)
        n = A.shape[0]
        self.n_iterations, self.n_atoms = n, A.shape[1] // 3
        self.reference_normalizer = False
        self.is_dense = True
        self.A, self.b = A, b
        self.c = np.zeros(n) if c is None else np.asarray(c, dtype=np.float64).reshape(n)
        self.d = np.zeros(n) if d is None else np.asarray(d, dtype=np.float64).reshape(n)
        self.table = None
        return self

    @staticmethod
    def dense_tables(A, b, c, d, s):
        """One iteration's cdw_dense_twist arrays from A_t, b_t, c_t, d_t and the per-coordinate variances s of the uncontrolled kernel:
        (theta, chol, beta, logk0, cd).  Raises if S^-1 + 2 A_t is not positive definite."""
        prec = np.diag(1.0 / s) + 2.0 * A
        try:
            np.linalg.cholesky(prec)
        except np.linalg.LinAlgError:
            raise ValueError("S^-1 + 2 A_t must be positive definite (the twisted proposal is a Gaussian; Theta_t, troubleshooting/csmc.py:460-478)")
        sig = np.linalg.inv(prec)
        sig = 0.5 * (sig + sig.T)
        theta = sig / s[None, :]
        beta = sig @ b
        logk0 = 0.5 * np.linalg.slogdet(theta)[1] + 0.5 * b @ beta - c - d
        return theta, np.linalg.cholesky(sig), beta, float(logk0), float(c + d)

    def bind(self, masses, langevin, device, n_particles):
        """Check the sequence against a system / time step and put its tables on `device` (idempotent for the same arguments)."""
        masses = np.asarray(masses, dtype=np.float64)
        if masses.shape[0] != self.n_atoms:
            raise ValueError(f"the twist is for {self.n_atoms} atoms, the system has {masses.shape[0]}")
        s_atom = langevin.kT * langevin.timestep ** 2 / masses          # s = 2 tau_a = kT dt^2 / m_a
        if not self.is_dense:
            self.check(masses, langevin)
            return self.to(device)
        D, n = 3 * self.n_atoms, self.n_iterations
        s = np.repeat(s_atom, 3)
        th, ch, be, self._logk0, self._cd = np.empty((n, D, D)), np.empty((n, D, D)), np.empty((n, D)), [], []
        for t in range(n):
            th[t], ch[t], be[t], k0, cd = self.dense_tables(self.A[t], self.b[t], self.c[t], self.d[t], s)
            self._logk0.append(k0); self._cd.append(cd)
        up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(device)
        self._dev = dict(theta=up(th), chol=up(ch), A=up(self.A), b=up(self.b), beta=up(be))
        self._scratch = torch.zeros(int(n_particles), 2 * D, dtype=torch.float64, device=device)
        self.table = self._dev["theta"]
        return self

    def to(self, device):
        if self.table is None or self.table.device != torch.device(device):
            self.table = torch.from_numpy(self.host).to(device)
        return self

    def check(self, masses, langevin):
        """1 + 2 s a > 0 for every coordinate (s = 2 tau_a = kT dt^2 / m_a): the twisted proposal must stay a proper Gaussian."""
        s = langevin.kT * langevin.timestep ** 2 / np.asarray(masses, dtype=np.float64)
        A = self.n_atoms
        if s.shape[0] != A:
            raise ValueError(f"the twist is for {A} atoms, the system has {s.shape[0]}")
        if ((1.0 + 2.0 * s[None, :, None] * self.host[:, :3 * A].reshape(-1, A, 3)) <= 0.0).any():
            raise ValueError("1 + 2 s a must be positive for every coordinate (Theta_t, troubleshooting/csmc.py:460-478)")

    def row(self, t):
        """Iteration t (1-based) -> what LangevinParameters.as_struct takes: the device row (diagonal) or a cdw_dense_twist descriptor."""
        if not self.is_dense:
            return self.table[t - 1]
        v = self._dev
        return _lib.DenseTwist(v["theta"][t - 1].data_ptr(), v["chol"][t - 1].data_ptr(), v["A"][t - 1].data_ptr(), v["b"][t - 1].data_ptr(),
                               v["beta"][t - 1].data_ptr(), self._logk0[t - 1], self._cd[t - 1], self._scratch.data_ptr())

    @property
    def flags(self):
        return _lib.CDW_FLAG_TWIST_REFERENCE_NORMALIZER if self.reference_normalizer else 0


class ParticleEnsemble:
    """All particles of one SMC run (or one rank's shard of them) as device-resident struct-of-arrays.

    Mirrors the per-particle fields of `Particle` (coddiwomple/particles.py:11-53): cumulative_work -> cum[P],
    auxiliary_work -> aux[P], ancestry[-1] -> label[P], state -> pos/vel[P][A] float4 rows, iteration -> scalar.
    pos/vel/aux/label (and the force rows that travel with a particle) are double-buffered: the resampling copy of the
    reference (resamplers.py:123-129) is the gather-on-load of the next propagate call.

    Two loops share this state:
      * look-ahead (MCMC moves, the default): `equip_lookahead` + `iterate_lookahead` — every launch also produces the next
        iteration's incremental works, and the resampling runs on a side stream beside the next launch (cdw_smc_lookahead,
        cdw_smc_weigh_resample);
      * sequential (Euler-Maruyama proposal, work tracking): `propagate_and_weigh` + `resample_fused`.
    """

    def __init__(self, system, n_particles, gid0=0, device=None, record_history=False, static_cache=True):
        _require_cuda()
        self.system = system if isinstance(system, DeviceSystem) else DeviceSystem(system)
        self.lib = self.system.lib
        self.P = int(n_particles)
        self.A = self.system.n_atoms
        self.gid0 = int(gid0)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        d, P, A = self.device, self.P, self.A
        self.pos = torch.zeros(2, P, A, 4, dtype=torch.float32, device=d)
        self.vel = torch.zeros(2, P, A, 4, dtype=torch.float32, device=d)
        self.aux = torch.zeros(2, P, dtype=torch.float64, device=d)
        self.label = torch.arange(P, dtype=torch.int32, device=d).repeat(2, 1) + self.gid0
        self.cum = torch.zeros(P, dtype=torch.float64, device=d)
        self.inc = torch.zeros(P, dtype=torch.float64, device=d)
        self.W = torch.zeros(P, dtype=torch.float64, device=d)
        # ancestors are double-buffered: the launch of iteration t gathers through anc_bufs[(t - 1) % 2] while the resampling of
        # iteration t, which runs beside it, writes anc_bufs[t % 2]
        self.anc_bufs = [torch.arange(P, dtype=torch.int32, device=d) for _ in range(2)]
        self.anc_cur = 0
        self.nan_flags = torch.zeros(P, dtype=torch.int32, device=d)
        self.wmin_key = torch.zeros(1, dtype=torch.int64, device=d)
        self.uniforms = torch.zeros(max(P, 1), dtype=torch.float64, device=d)
        self.cdf = torch.zeros(P + 2, dtype=torch.int64, device=d)[:P]
        self.ws_bytes = int(self.lib.cdw_weight_workspace_bytes(P))
        self.workspace = torch.zeros((self.ws_bytes + 7) // 8, dtype=torch.int64, device=d)
        # rows that travel with a particle besides (x, v): the look-ahead loop keeps the TOTAL forces F(x; lambda_next) there,
        # the sequential loop the forces / energy of the lambda-independent terms (include/cdw.h: cdw_static_cache)
        self.static_cache = bool(static_cache)
        self.fs = torch.zeros(2, P, A, 4, dtype=torch.float32, device=d)
        self.us = torch.zeros(2, P, dtype=torch.float64, device=d) if self.static_cache else None
        self.inc_ahead = torch.zeros(2, P, dtype=torch.float64, device=d)   # [t % 2]: incremental works of iteration t + 1, at the slots of launch t
        self.cache_valid = False
        self.cur = 0
        self.pending_gather = False  # anc has not been applied to (pos, vel, aux, label) yet
        self.iteration = 0
        self.record_history = record_history
        self.history = []
        self._side = torch.cuda.Stream(device=d)
        self._resampled = None   # event: the resampling launched last has finished (side stream)

    @property
    def anc(self):
        """Ancestors still to be applied to the current rows."""
        return self.anc_bufs[self.anc_cur]

    # ------------------------------------------------------------------ state access
    def set_positions(self, positions, velocities=None):
        rows = positions if torch.is_tensor(positions) else positions_to_rows(positions, self.device)
        assert tuple(rows.shape) == (self.P, self.A, 4), rows.shape
        self.pos[self.cur].copy_(rows, non_blocking=True)
        self.cache_valid = False
        if velocities is not None:
            vrows = velocities if torch.is_tensor(velocities) else positions_to_rows(velocities, self.device)
            self.vel[self.cur].copy_(vrows, non_blocking=True)
        else:
            self.vel[self.cur].zero_()
        self.pending_gather = False

    def join(self):
        """Make the current stream wait for the resampling that runs on the side stream."""
        if self._resampled is not None:
            torch.cuda.current_stream().wait_event(self._resampled)
            self._resampled = None

    def materialize(self):
        """Apply a pending resampling gather (K3b) so that pos/vel/aux/label hold the resampled particles."""
        self.join()
        if self.pending_gather:
            c, n = self.cur, 1 - self.cur
            _lib.check(self.lib.cdw_gather_particles(_lib.ptr(self.pos[c]), _lib.ptr(self.vel[c]), _lib.ptr(self.anc), _lib.ptr(self.pos[n]),
                                                    _lib.ptr(self.vel[n]), _lib.ptr(self.aux[c]), _lib.ptr(self.aux[n]), _lib.ptr(self.label[c]),
                                                    _lib.ptr(self.label[n]), self.P, self.A, _lib.current_stream()), "cdw_gather_particles")
            self.cur = n
            self.pending_gather = False
            self.cache_valid = False  # the rows that travel with the particles were not gathered

    @property
    def positions(self):
        self.materialize()
        return self.pos[self.cur][..., :3]

    @property
    def velocities(self):
        self.materialize()
        return self.vel[self.cur][..., :3]

    @property
    def auxiliary_work(self):
        self.materialize()
        return self.aux[self.cur]

    @property
    def labels(self):
        self.materialize()
        return self.label[self.cur]

    # ------------------------------------------------------------------ kernels
    def reduced_potential(self, lam_row, beta):
        """K1 on the current positions: u[p] = beta U(x_p; lambda)."""
        self.materialize()
        u = torch.empty(self.P, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.cdw_reduced_potential(self.system.handle, _lib.ptr(self.pos[self.cur]), self.P, _lib.ptr(lam_row), float(beta),
                                                 _lib.ptr(u), _lib.current_stream()), "cdw_reduced_potential")
        return u

    # ---- look-ahead loop
    def equip_lookahead(self, lam0, lam1, langevin):
        """aux = -u_0(x_0), cum = 0 (ProposalFactory.equip_initial_sample, distribution_factories.py:180-220), plus what iteration 1
        needs: F(x_0; lambda_1) and inc_1 = u_1(x_0) - u_0(x_0) (cdw_smc_lookahead with n_steps = 0)."""
        self.materialize()
        c = self.cur
        prm = langevin.as_struct(0, n_steps=0)
        _lib.check(self.lib.cdw_smc_lookahead(
            self.system.handle, _lib.ptr(self.pos[c]), _lib.ptr(self.vel[c]), None, None, None, _lib.ptr(self.fs[c]), self.P, None, None, _lib.ptr(lam0),
            _lib.ptr(lam1), C.byref(prm), C.c_uint64(0), C.c_uint64(0), C.c_int64(self.gid0), _lib.ptr(self.inc_ahead[0]) if lam1 is not None else None,
            _lib.ptr(self.aux[c]), None, None, _lib.current_stream()), "cdw_smc_lookahead")
        self.cum.zero_()
        self.iteration = 0
        self.anc_cur = 0
        self.wmin_key.fill_(-1)  # ~0ull: arm the running min; the statistics kernel re-arms it after every use
        self.cache_valid = False
        self._resampled = None

    def iterate_lookahead(self, t, lam_t, lam_next, langevin, seed, stats_row, kind, floor_threshold, resample_seed, randomize_velocities=False,
                          after_resample=None):
        """Iteration t of the reference loop (openmm/coddiwomple.py:314-328) as two concurrent pieces:
          side stream   W_t = cum + inc_t (the works the PREVIOUS launch left behind), nESS, decision, ancestors a_t, cum   (cdw_smc_weigh_resample)
          this stream   gather by a_{t-1}, (V R O R E V)^n at lambda_t, then E at lambda_{t+1}: forces and inc_{t+1}         (cdw_smc_lookahead)
        `after_resample(t)` (optional) is called with the side stream current, right behind the resampling."""
        self.iteration = t
        main = torch.cuda.current_stream()
        prev_anc = self.anc_bufs[(t - 1) % 2] if t > 1 else None
        launched = torch.cuda.Event()
        launched.record(main)
        self._side.wait_event(launched)           # the launch that wrote inc_ahead[(t - 1) % 2] (and read anc_bufs[t % 2]) is behind us
        with torch.cuda.stream(self._side):
            _lib.check(self.lib.cdw_smc_weigh_resample(
                int(kind), _lib.ptr(self.inc_ahead[(t - 1) % 2]), _lib.ptr(prev_anc), self.P, float(floor_threshold), _lib.ptr(self.wmin_key),
                C.c_uint64(resample_seed), C.c_uint64(t), _lib.ptr(stats_row), _lib.ptr(self.cum), _lib.ptr(self.inc), _lib.ptr(self.W),
                _lib.ptr(self.anc_bufs[t % 2]), _lib.ptr(self.cdf), _lib.ptr(self.uniforms), _lib.ptr(self.workspace), self.ws_bytes,
                _lib.current_stream()), "cdw_smc_weigh_resample")
            if after_resample is not None:
                after_resample(t)
            resampled = torch.cuda.Event()
            resampled.record(self._side)
        if self._resampled is not None:           # a_{t-1}: the resampling launched one iteration ago
            main.wait_event(self._resampled)
        c, n = self.cur, 1 - self.cur
        prm = langevin.as_struct(_lib.CDW_FLAG_RANDOMIZE_VELOCITIES if randomize_velocities else 0)
        _lib.check(self.lib.cdw_smc_lookahead(
            self.system.handle, _lib.ptr(self.pos[c]), _lib.ptr(self.vel[c]), _lib.ptr(self.fs[c]), _lib.ptr(self.pos[n]), _lib.ptr(self.vel[n]),
            _lib.ptr(self.fs[n]), self.P, _lib.ptr(prev_anc), _lib.ptr(self.label[c]), _lib.ptr(lam_t), _lib.ptr(lam_next), C.byref(prm), C.c_uint64(seed),
            C.c_uint64(t * max(langevin.n_steps, 1)), C.c_int64(self.gid0), _lib.ptr(self.inc_ahead[t % 2]) if lam_next is not None else None,
            _lib.ptr(self.aux[n]), _lib.ptr(self.label[n]), _lib.ptr(self.nan_flags), _lib.current_stream()), "cdw_smc_lookahead")
        self._resampled = resampled
        self.cur = n
        self.anc_cur = t % 2
        self.pending_gather = True
        self.cache_valid = False

    # ---- sequential loop
    def seed_static_cache(self, lam_row, kT, aux_out=None):
        """Evaluate the static terms at the current rows (an n_steps = 0 cdw_smc_propagate_cached call): fills fs/us[cur]
        and, optionally, aux_out = -u(x; lambda)."""
        self.materialize()
        c = self.cur
        prm = _lib.LangevinParams(0.0, 0.0, float(kT), 1e-6, 0, 0)
        cache = _lib.static_cache(None, None, self.fs[c], self.us[c])
        _lib.check(self.lib.cdw_smc_propagate_cached(
            self.system.handle, _lib.ptr(self.pos[c]), _lib.ptr(self.vel[c]), None, None, self.P, None, None, None, None, _lib.ptr(lam_row), C.byref(prm),
            C.c_uint64(0), C.c_uint64(0), C.c_int64(self.gid0), None, None, _lib.ptr(aux_out), None, None, None, None, None, None, None, C.byref(cache),
            _lib.current_stream()), "cdw_smc_propagate_cached")
        self.cache_valid = True

    def equip_initial(self, lam_row, beta):
        """aux = -u_0(x_0); cum = 0 (ProposalFactory.equip_initial_sample, distribution_factories.py:180-220)."""
        if self.static_cache:
            self.seed_static_cache(lam_row, 1.0 / beta, aux_out=self.aux[self.cur])
            u0 = -self.aux[self.cur]
        else:
            u0 = self.reduced_potential(lam_row, beta)
            self.aux[self.cur].copy_(-u0)
        self.cum.zero_()
        self.iteration = 0
        self.anc_cur = 0
        self._resampled = None
        self.wmin_key.fill_(-1)  # ~0ull: arm the running min; cdw_smc_resample re-arms it after every iteration
        return u0

    def propagate_and_weigh(self, lam_row, langevin, seed, randomize_velocities=False, track_works=False, outputs=None, ula=False, twist=None):
        """Fused K3b-on-load + K1 + K4 + first half of K2 (include/cdw.h: cdw_smc_propagate).  Advances `iteration`.
        ula=True swaps BAOAB for the Euler-Maruyama proposal and its generalised incremental work (CDW_FLAG_ULA); `twist` (a TwistSequence)
        twists that proposal with the row of this iteration."""
        self.join()
        self.iteration += 1
        c, n = self.cur, 1 - self.cur
        flags = (_lib.CDW_FLAG_RANDOMIZE_VELOCITIES if randomize_velocities else 0) | (_lib.CDW_FLAG_TRACK_WORKS if track_works else 0)
        if ula:
            flags = _lib.CDW_FLAG_ULA | (twist.flags if twist is not None else 0)
        elif twist is not None:
            raise ValueError("a twist applies to the Euler-Maruyama proposal only")
        prm = langevin.as_struct(flags, twist=None if twist is None else twist.row(self.iteration))
        o = outputs or {}
        use_cache = self.static_cache and not track_works and langevin.n_steps > 0 and not getattr(langevin, "splitting_word", 0)
        if use_cache and not self.cache_valid:
            self.seed_static_cache(lam_row, langevin.kT)
            c, n = self.cur, 1 - self.cur
        cache = _lib.static_cache(self.fs[c], self.us[c], self.fs[n], self.us[n]) if use_cache else None
        anc = self.anc if self.pending_gather else None
        _lib.check(self.lib.cdw_smc_propagate_cached(
            self.system.handle, _lib.ptr(self.pos[c]), _lib.ptr(self.vel[c]), _lib.ptr(self.pos[n]), _lib.ptr(self.vel[n]), self.P,
            _lib.ptr(anc), _lib.ptr(self.aux[c]), _lib.ptr(self.cum), _lib.ptr(self.label[c]), _lib.ptr(lam_row), C.byref(prm),
            C.c_uint64(seed), C.c_uint64(self.iteration * max(langevin.n_steps, 1)), C.c_int64(self.gid0), _lib.ptr(self.inc), _lib.ptr(self.W),
            _lib.ptr(self.aux[n]), _lib.ptr(self.label[n]), _lib.ptr(o.get("u_first")), _lib.ptr(o.get("u_last")), _lib.ptr(o.get("shadow")),
            _lib.ptr(o.get("proposal")), _lib.ptr(self.nan_flags), _lib.ptr(self.wmin_key), C.byref(cache) if cache is not None else None,
            _lib.current_stream()), "cdw_smc_propagate_cached")
        self.cache_valid = use_cache
        self.cur = n
        self.pending_gather = False

    def resample_fused(self, stats_row, kind, floor_threshold, seed, stream_id):
        """K2 second half + K3a (+ uniforms, + re-arming the running min) in one call: cdw_smc_resample."""
        out = self.anc_bufs[1 - self.anc_cur]
        _lib.check(self.lib.cdw_smc_resample(int(kind), _lib.ptr(self.W), self.P, float(floor_threshold), _lib.ptr(self.wmin_key), C.c_uint64(seed),
                                            C.c_uint64(stream_id), _lib.ptr(stats_row), _lib.ptr(self.cum), _lib.ptr(self.cum), _lib.ptr(out),
                                            _lib.ptr(self.cdf), _lib.ptr(self.uniforms), _lib.ptr(self.workspace), self.ws_bytes, _lib.current_stream()),
                   "cdw_smc_resample")
        self.anc_cur = 1 - self.anc_cur
        self.pending_gather = True

    def weight_stats(self, stats_row, floor_threshold, w=None, wmin_key=None, n=None):
        """Second half of K2: log-sum-exp, nESS, the resampling decision -> stats_row[8] (device)."""
        w = self.W if w is None else w
        _lib.check(self.lib.cdw_weight_stats(_lib.ptr(w), int(w.numel() if n is None else n), float(floor_threshold),
                                            _lib.ptr(self.wmin_key if wmin_key is None else wmin_key), _lib.ptr(stats_row),
                                            _lib.ptr(self.workspace), self.ws_bytes, _lib.current_stream()), "cdw_weight_stats")

    def draw_uniforms(self, seed, stream_id, n):
        _lib.check(self.lib.cdw_philox_uniforms(C.c_uint64(seed), C.c_uint64(stream_id), int(n), _lib.ptr(self.uniforms), _lib.current_stream()),
                   "cdw_philox_uniforms")

    def resample(self, stats_row, kind, uniforms=None):
        """K3a: ancestors from the integer cdf (identity + null update when the decision in stats_row says so)."""
        u = self.uniforms if uniforms is None else uniforms
        out = self.anc_bufs[1 - self.anc_cur]
        _lib.check(self.lib.cdw_resample_indices(int(kind), _lib.ptr(self.W), _lib.ptr(u), self.P, _lib.ptr(stats_row), _lib.ptr(self.cum),
                                                _lib.ptr(self.cum), _lib.ptr(out), _lib.ptr(self.cdf), _lib.ptr(self.workspace), self.ws_bytes,
                                                _lib.current_stream()), "cdw_resample_indices")
        self.anc_cur = 1 - self.anc_cur
        self.pending_gather = True


RESAMPLER_KINDS = {"multinomial": _lib.CDW_RESAMPLE_MULTINOMIAL, "systematic": _lib.CDW_RESAMPLE_SYSTEMATIC}


class SMCEngine:
    """The whole anneal on one GPU, no host synchronisation inside (the resampling decision is taken on the device).

    Order of operations = the reference loop (openmm/coddiwomple.py:312-331), see oracle/smc.py for the array form.
    MCMC moves (`proposal="baoab"`) run the look-ahead loop: ONE propagate launch per iteration on the main stream, the weight
    statistics / resampling of the same iteration beside it on a side stream.  The Euler-Maruyama proposal, whose weights need the
    move itself, and `lookahead=False` run propagate and resample back to back.
    """

    def __init__(self, system, n_particles, parameter_sequence, langevin=None, resampler="multinomial", floor_threshold=0.5,
                 always_resample=False, seed=0, resample_seed=1, gid0=0, record_history=False, device=None, record_works_every=None,
                 proposal="baoab", static_cache=True, record_lineage=False, lookahead=True, twist=None):
        if proposal not in ("baoab", "ula"):
            raise ValueError("proposal must be 'baoab' (MCMC move, openmm/integrators.py:19-465) or 'ula' (Euler-Maruyama, :666-888)")
        if twist is not None and proposal != "ula":
            raise ValueError("twist: controlled SMC twists the Euler-Maruyama proposal (proposal='ula', troubleshooting/csmc.py:625-747)")
        self.proposal = proposal
        self.twist = twist
        # the look-ahead launch is the fused "V R O R V" body; other splittings (Metropolised blocks, ...) run the programmable one
        self.lookahead = bool(lookahead) and proposal == "baoab" and not getattr(langevin, "splitting_word", 0)
        self.ensemble = ParticleEnsemble(system, n_particles, gid0=gid0, device=device, record_history=record_history, static_cache=static_cache)
        self.system = self.ensemble.system
        self.team = self.system.choose_team(n_particles)
        self.langevin = langevin or LangevinParameters()
        self.parameter_sequence = list(parameter_sequence)
        self.T = len(self.parameter_sequence)
        self.lam = self.system.lambda_table(self.parameter_sequence, self.ensemble.device)
        self.system.register_protocol(self.lam, self.langevin.kT)
        if twist is not None:
            if twist.n_iterations != self.T - 1:
                raise ValueError(f"the twist sequence holds {twist.n_iterations} iterations, the protocol has {self.T - 1}")
            twist.bind(self.system.spec.masses, self.langevin, self.ensemble.device, n_particles)
        self.kind = RESAMPLER_KINDS[resampler] if isinstance(resampler, str) else int(resampler)
        self.threshold = -1.0 if always_resample else float(floor_threshold)
        self.seed, self.resample_seed = int(seed), int(resample_seed)
        self.stats = torch.zeros(max(self.T, 1), _lib.STATS_LEN, dtype=torch.float64, device=self.ensemble.device)
        self.graph = None
        # optional cumulative-work trajectory (the AIS state-work record, openmm/propagators.py:311-337)
        self.record_works_every = int(record_works_every) if record_works_every else None
        if self.record_works_every:
            k = self.record_works_every
            self._work_rows = [t for t in range(1, self.T) if t % k == 0 or t == self.T - 1]
            self._works = torch.zeros(len(self._work_rows) + 1, self.ensemble.P, dtype=torch.float64, device=self.ensemble.device)
        # optional per-iteration spill of (cumulative work, root label) of every slot: what the reference keeps in
        # Particle._incremental_works / Particle._ancestry (particles.py:38-47); row t - 1 = state of the ensemble after iteration t
        self.record_lineage = bool(record_lineage)
        if self.record_lineage:
            self._cum_hist = torch.zeros(max(self.T - 1, 1), self.ensemble.P, dtype=torch.float64, device=self.ensemble.device)
            self._label_hist = torch.zeros(max(self.T - 1, 1), self.ensemble.P, dtype=torch.int32, device=self.ensemble.device)
            self._lineage_due = None

    @property
    def beta(self):
        return 1.0 / self.langevin.kT

    def load(self, positions, velocities=None):
        """Stage the initial particles (samples of the first target) into buffer 0."""
        e = self.ensemble
        e.cur, e.pending_gather, e.iteration, e.anc_cur, e._resampled = 0, False, 0, 0, None
        e.set_positions(positions, velocities)

    def _equip(self):
        if self.lookahead:
            self.ensemble.equip_lookahead(self.lam[0], self.lam[1] if self.T > 1 else None, self.langevin)
        else:
            self.ensemble.equip_initial(self.lam[0], self.beta)
        self.stats.zero_()
        if self.record_lineage:
            self._lineage_due = None

    def initialize(self, positions, velocities=None, initial_works=None):
        """Stage the initial particles and equip them (aux = -u_0(x_0), cumulative work 0).  `initial_works` (P,): the works -log w_0 of particles
        that were NOT drawn from the first target itself but from a proposal q_0 with weights w_0 = pi_0 / q_0 — the `initial_work` of
        ProposalFactory.equip_initial_sample with a generation_pdf (distribution_factories.py:203-212), the log w_0 of a twisted initial
        proposal (troubleshooting/csmc.py:564-591, `csmc.twisted_gmm_initial_sample`)."""
        self.load(positions, velocities)
        self._equip()
        if initial_works is not None:
            w = torch.as_tensor(np.asarray(initial_works, dtype=np.float64)).to(self.ensemble.device)
            if w.shape != (self.ensemble.P,):
                raise ValueError("initial_works must have one entry per particle")
            self.ensemble.cum.copy_(w)

    # ------------------------------------------------------------------ sequential loop (also the public two-call form)
    def propagate(self, t):
        """First kernel of iteration t (1-based): gather-on-load + E/F + incremental work + move + closing energy."""
        self.ensemble.propagate_and_weigh(self.lam[t], self.langevin, self.seed, randomize_velocities=(t == 1), ula=(self.proposal == "ula"), twist=self.twist)

    def finish_iteration(self, t):
        """Rest of iteration t: log-sum-exp / nESS / decision, uniforms, integer cdf + ancestors, cum update."""
        self.ensemble.resample_fused(self.stats[t], self.kind, self.threshold, self.resample_seed, t)
        self._after_resample(t)
        self._record_lineage(t)

    # ------------------------------------------------------------------ bookkeeping hooks
    def _after_resample(self, t):
        """Runs right behind the resampling of iteration t (on the stream it ran on)."""
        e = self.ensemble
        if self.record_works_every and t in self._work_rows:
            self._works[self._work_rows.index(t) + 1].copy_(e.cum)
        if e.record_history:
            n_u = e.P if self.kind == _lib.CDW_RESAMPLE_MULTINOMIAL else 1
            anc = e.anc_bufs[t % 2] if self.lookahead else e.anc
            e.history.append(dict(t=t, inc=e.inc.clone(), W=e.W.clone(), cum=e.cum.clone(), anc=anc.clone(), uniforms=e.uniforms[:n_u].clone()))

    def _record_lineage(self, t):
        """Spill (cum, root label) after iteration t; needs both the launch (labels) and the resampling (ancestors, cum) of t."""
        if not self.record_lineage:
            return
        e = self.ensemble
        _lib.check(e.lib.cdw_record_lineage(_lib.ptr(e.cum), _lib.ptr(e.label[e.cur]), _lib.ptr(e.anc if e.pending_gather else None),
                                           _lib.ptr(self._cum_hist[t - 1]), _lib.ptr(self._label_hist[t - 1]), e.P, _lib.current_stream()),
                   "cdw_record_lineage")

    def step(self, t):
        """Iteration t of the reference loop (openmm/coddiwomple.py:314-328)."""
        if not self.lookahead:
            self.propagate(t)
            self.finish_iteration(t)
            return
        e = self.ensemble
        if self.record_lineage and self._lineage_due is not None:
            # iteration t - 1 is complete once its resampling has finished; cum is rewritten by the resampling of iteration t, so the
            # spill has to sit between the two (it costs the overlap: histories are a debugging / analysis aid)
            e.join()
            self._record_lineage(self._lineage_due)
            self._lineage_due = None
        e.iterate_lookahead(t, self.lam[t], self.lam[t + 1] if t + 1 < self.T else None, self.langevin, self.seed, self.stats[t], self.kind,
                            self.threshold, self.resample_seed, randomize_velocities=(t == 1), after_resample=self._after_resample)
        if self.record_lineage:
            self._lineage_due = t

    def finish(self):
        """Join the side stream, spill the last iteration's history if histories are kept, and apply the last ancestors."""
        self.ensemble.join()
        if self.record_lineage and self.lookahead and self._lineage_due is not None:
            self._record_lineage(self._lineage_due)
            self._lineage_due = None
        self.ensemble.materialize()
        return self

    def step_with_host_hooks(self, t, decide=None, draw_ancestors=None):
        """Iteration t with the caller's own Python observable / threshold / resampler in the loop (the reference's pluggable
        `observable(particles, incremental_works)`, `threshold(value)` and `Resampler._resample(cumulative_works, num_resamples)`,
        resamplers.py:61-133, 202-216): the particles stay on the device and the propagate launch is the usual one; only the 8-byte
        works travel to the host and only the decision and the ancestors travel back (one host synchronisation per iteration).
          decide(cum, inc) -> bool                    cum, inc: host arrays of the cumulative works BEFORE this iteration and its increments
          draw_ancestors(W) -> indices                W = cum + inc; None: the engine's own resampler kind
        Bookkeeping as the reference: resampled -> cum[i] = -logsumexp(-W) + log P (resamplers.py:107,196-197), else cum[i] = W[i]."""
        if self.lookahead:
            raise ValueError("host hooks need the sequential loop (SMCEngine(..., lookahead=False))")
        e = self.ensemble
        self.propagate(t)
        cum_before = e.cum.cpu().numpy()
        inc = e.inc.cpu().numpy()
        wanted = True if decide is None else bool(decide(cum_before, inc))
        row = self.stats[t]
        if draw_ancestors is None:
            # the library's resampler with the decision forced: threshold < 0 always resamples, 0 never does (nESS > 0)
            e.resample_fused(row, self.kind, -1.0 if wanted else 0.0, self.resample_seed, t)
        else:
            e.weight_stats(row, -1.0 if wanted else 0.0)     # normaliser, nESS (and the forced decision) -> stats[t]
            e.wmin_key.fill_(-1)
            out = e.anc_bufs[1 - e.anc_cur]
            if wanted:
                idx = np.asarray(draw_ancestors(cum_before + inc), dtype=np.int64).reshape(-1)
                if idx.shape[0] != e.P or idx.min() < 0 or idx.max() >= e.P:
                    raise ValueError("the resampler must return one ancestor index in [0, P) per particle")
                out.copy_(torch.from_numpy(idx.astype(np.int32)))
                e.cum.copy_(row[_lib.STAT_MEAN].expand(e.P))
            else:
                out.copy_(torch.arange(e.P, dtype=torch.int32, device=e.device))
                e.cum.copy_(e.W)
            e.anc_cur = 1 - e.anc_cur
            e.pending_gather = True
        self._after_resample(t)
        self._record_lineage(t)
        return wanted

    def run(self):
        for t in range(1, self.T):
            self.step(t)
        return self.finish()

    def anneal(self, positions, velocities=None, initial_works=None):
        if self.graph is not None:
            if initial_works is not None:
                raise ValueError("initial_works need the eager loop (the captured graph equips the particles with zero works)")
            return self.anneal_graph(positions, velocities)
        self.initialize(positions, velocities, initial_works)
        return self.run()

    # ------------------------------------------------------------------ CUDA graph of the whole anneal
    def _body(self):
        self._equip()
        self.run()

    def capture(self):
        """Capture `equip initial sample + T-1 iterations + final gather` into one CUDA graph (look-ahead loop: a chain of T-1
        propagate launches with the resampling kernels forked beside it).  The loop has no host decision (the resampling
        test runs on the device), so the whole anneal replays as a single graph launch; positions are (re)staged with load()
        before each replay."""
        e = self.ensemble
        if e.record_history:
            raise ValueError("record_history needs the eager loop")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture (lazy module loading, cudaFuncSetAttribute, ...)
            e.cur, e.pending_gather, e.iteration, e.anc_cur, e._resampled = 0, False, 0, 0, None
            self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        e.cur, e.pending_gather, e.iteration, e.anc_cur, e._resampled = 0, False, 0, 0, None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body()
        self._graph_final = (e.cur, e.pending_gather, e.iteration, e.anc_cur)
        self.graph = g
        return self

    def anneal_graph(self, positions, velocities=None):
        self.load(positions, velocities)
        self.graph.replay()
        e = self.ensemble
        e.cur, e.pending_gather, e.iteration, e.anc_cur = self._graph_final
        return self

    # ------------------------------------------------------------------ results (host side; these synchronise)
    def cumulative_works(self):
        return self.ensemble.cum.cpu().numpy()

    def free_energy(self):
        """-log mean exp(-cum) in kT (reduced on the device)."""
        c = self.ensemble.cum
        return float(-torch.logsumexp(-c, 0) + np.log(c.numel()))

    def resampled_iterations(self):
        s = self.stats[:, _lib.STAT_RESAMPLE].cpu().numpy()
        return [t for t in range(1, self.T) if s[t] != 0.0]

    def work_history(self):
        """(n_records + 1, P) cumulative works in kT; row 0 is the zero every AIS record starts with."""
        return self._works.cpu().numpy()

    def lineage(self):
        """dict(cum=(T-1, P) cumulative works, labels=(T-1, P) root labels) after every iteration (host arrays)."""
        if not self.record_lineage:
            raise _lib.CdwError("the per-iteration histories were not recorded: construct the engine with record_lineage=True")
        return dict(cum=self._cum_hist.cpu().numpy(), labels=self._label_hist.cpu().numpy())

    def ness_trace(self):
        return self.stats[1:, _lib.STAT_NESS].cpu().numpy()
