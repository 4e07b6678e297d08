"""Multi-GPU SMC annealing: particles sharded over the GPUs of one box, one process per GPU.

The reference has no parallelism of any kind; this is new.  Design ("replicated weights, sharded
state"):

  * rank g owns the contiguous block of global particle ids [g * P/G, (g + 1) * P/G): positions, velocities, aux and
    labels of those slots live in ITS HBM only (double-buffered rows, exported to the other ranks with CUDA IPC so
    that every rank can read every rank's rows over NVLink);
  * the per-particle weights are tiny (8 B) and every rank needs the global normaliser, nESS, the resampling decision
    and the global cdf: each rank contributes its slice of incremental works to ONE NCCL all-gather per iteration
    (the (max, sum-exp) all-reduce of the north star is subsumed: every rank reduces the gathered vector itself with
    exact integer sums, so the statistics are bit-identical on all ranks and identical to a single-GPU run); in the
    look-ahead loop that all-gather and the resampling run on a side stream BESIDE the propagate launch;
  * ancestors are global ids; the next propagate kernel loads each slot's ancestor row straight from the owning
    rank's buffer through a peer-mapped pointer: the resampling gather, the NVLink all-to-all migration and the
    propagation are one kernel (cdw_smc_propagate_sharded);
  * cross-rank ordering (a rank's rows are only read by peers after the launch that wrote them has finished) is a
    one-block peer barrier over NVLink flags in front of every launch (cdw_peer_barrier); in the sequential loop the
    all-gather itself is the ordering point.

Noise is keyed by the global particle id and the weights are reduced identically everywhere, so an anneal on G GPUs
is bit-identical to the same anneal on one GPU (tests/multi_gpu_check.py asserts exactly that).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import RESAMPLER_KINDS, DeviceSystem, LangevinParameters, positions_to_rows


def shard_bounds(n_particles, world_size, rank):
    """[first, last) global ids of `rank`'s block; the particle count must divide evenly."""
    if n_particles % world_size:
        raise ValueError(f"{n_particles} particles do not divide over {world_size} ranks")
    per = n_particles // world_size
    return rank * per, (rank + 1) * per


def owner_of(global_ids, n_per_rank):
    """(owning rank, local slot) of global particle ids."""
    g = np.asarray(global_ids)
    return g // n_per_rank, g % n_per_rank


def migration_matrix(ancestors, world_size):
    """M[dst, src] = number of particle rows rank dst pulls from rank src for the global ancestor vector."""
    anc = np.asarray(ancestors)
    per = len(anc) // world_size
    dst = np.arange(len(anc)) // per
    src = anc // per
    m = np.zeros((world_size, world_size), dtype=np.int64)
    np.add.at(m, (dst, src), 1)
    return m


class _IpcBuffer:
    """A cudaMalloc'ed buffer owned by this rank, viewable as a torch tensor and exportable to peers."""

    def __init__(self, lib, shape, dtype):
        self.lib = lib
        self.shape, self.dtype = tuple(shape), dtype
        self.nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.ptr = C.c_void_p()
        _lib.check(lib.cdw_device_malloc(max(self.nbytes, 16), C.byref(self.ptr)), "cdw_device_malloc")
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}[dtype]
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.ptr.value, False), "version": 2}
        self.tensor = torch.as_tensor(self, device=f"cuda:{torch.cuda.current_device()}")
        self.tensor.zero_()

    def handle(self):
        h = (C.c_ubyte * 64)()
        _lib.check(self.lib.cdw_ipc_get_handle(self.ptr, h), "cdw_ipc_get_handle")
        return bytes(h)

    def free(self):
        if self.ptr:
            self.lib.cdw_device_free(self.ptr)
            self.ptr = C.c_void_p()


class ShardedSMCEngine:
    """Same loop as SMCEngine over `n_particles` GLOBAL particles, sharded over the ranks of `group` (NCCL).

    MCMC moves run the look-ahead loop (cdw_smc_lookahead_sharded): per iteration and rank
      main stream   peer barrier (every rank's rows of iteration t - 1 are in place), then ONE launch: gather the ancestors' rows
                    over NVLink, integrate, evaluate at lambda_{t+1} (forces for the next launch, incremental works inc_{t+1});
      side stream   all-gather of the 8-byte incremental works, then the global resampling of iteration t on the replicated weight
                    vector — beside the launch, so neither the collective's latency nor the resampling is on the critical path.
    The Euler-Maruyama proposal (weights need the move) keeps propagate -> all-gather -> resample back to back.
    """

    def __init__(self, system, n_particles, parameter_sequence, langevin=None, resampler="multinomial", floor_threshold=0.5, always_resample=False,
                 seed=0, resample_seed=1, group=None, proposal="baoab", twist=None):
        if proposal not in ("baoab", "ula"):
            raise ValueError("proposal must be 'baoab' or 'ula'")
        if twist is not None and proposal != "ula":
            raise ValueError("twist: controlled SMC twists the Euler-Maruyama proposal (proposal='ula')")
        self.proposal = proposal
        self.twist = twist
        self.lookahead = proposal == "baoab"
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (backend nccl, one process per GPU)")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.P = int(n_particles)
        self.first, self.last = shard_bounds(self.P, self.world, self.rank)
        self.P_loc = self.last - self.first
        self.system = system if isinstance(system, DeviceSystem) else DeviceSystem(system)
        self.lib = self.system.lib
        self.A = self.system.n_atoms
        self.team = self.system.choose_team(self.P)   # from the GLOBAL ensemble size: the same on every rank and in a single-GPU run
        self.langevin = langevin or LangevinParameters()
        self.parameter_sequence = list(parameter_sequence)
        self.T = len(self.parameter_sequence)
        self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.lam = self.system.lambda_table(self.parameter_sequence, self.device)
        self.system.register_protocol(self.lam, self.langevin.kT)
        if twist is not None:
            if twist.n_iterations != self.T - 1:
                raise ValueError(f"the twist sequence holds {twist.n_iterations} iterations, the protocol has {self.T - 1}")
            twist.bind(self.system.spec.masses, self.langevin, self.device, self.P_loc)
        self.kind = RESAMPLER_KINDS[resampler] if isinstance(resampler, str) else int(resampler)
        self.threshold = -1.0 if always_resample else float(floor_threshold)
        self.seed, self.resample_seed = int(seed), int(resample_seed)
        d, P, Pl, A = self.device, self.P, self.P_loc, self.A
        # ---- sharded state, double-buffered, IPC-exported ("fs": the force rows that travel with a particle — total forces at the
        # next lambda in the look-ahead loop, static-term forces in the sequential loop)
        self._opened = []
        self._bufs = {name: [_IpcBuffer(self.lib, shape, dt) for _ in range(2)]
                      for name, shape, dt in (("pos", (Pl, A, 4), torch.float32), ("vel", (Pl, A, 4), torch.float32), ("aux", (Pl,), torch.float64),
                                              ("label", (Pl,), torch.int32), ("fs", (Pl, A, 4), torch.float32), ("us", (Pl,), torch.float64))}
        self.peer = {name: [self._exchange(name, b) for b in range(2)] for name in self._bufs}
        # cross-rank ordering flags (cdw_peer_barrier): flags[r] = last epoch rank r has completed
        self._flags = _IpcBuffer(self.lib, (2 * self.world,), torch.float64)   # (uint64 words; the dtype only sizes the buffer)
        self._peer_flags = self._exchange_buffer(self._flags)
        self._barrier_status = torch.zeros(1, dtype=torch.int32, device=d)
        self._epoch_counter = torch.zeros(1, dtype=torch.int64, device=d)   # incremented by the barrier launches themselves (graph-replayable)
        # ---- replicated weights (global length on every rank)
        self.W = torch.zeros(P, dtype=torch.float64, device=d)
        self.inc = torch.zeros(P, dtype=torch.float64, device=d)
        self.W_loc = torch.zeros(Pl, dtype=torch.float64, device=d)
        self.inc_loc = torch.zeros(Pl, dtype=torch.float64, device=d)
        self.inc_ahead_loc = torch.zeros(2, Pl, dtype=torch.float64, device=d)   # [t % 2]: inc_{t+1} of this rank's slots after launch t
        self.inc_ahead = torch.zeros(P, dtype=torch.float64, device=d)           # all-gathered
        self.cum = torch.zeros(P, dtype=torch.float64, device=d)
        # ancestors are double-buffered: launch t gathers through anc_bufs[(t - 1) % 2] while the resampling of iteration t, which
        # runs beside it, writes anc_bufs[t % 2]
        self.anc_bufs = [torch.arange(P, dtype=torch.int32, device=d) for _ in range(2)]
        self.anc_cur = 0
        self._side = torch.cuda.Stream(device=d)
        self._resampled = None
        self.nan_flags = torch.zeros(Pl, dtype=torch.int32, device=d)
        self.wmin_key = torch.zeros(1, dtype=torch.int64, device=d)
        self.scratch_key = torch.zeros(1, dtype=torch.int64, device=d)
        self.uniforms = torch.zeros(P, dtype=torch.float64, device=d)
        self.cdf = torch.zeros(P + 2, dtype=torch.int64, device=d)[:P]
        self.W_copy = torch.zeros(P, dtype=torch.float64, device=d)
        self.ws_bytes = int(self.lib.cdw_weight_workspace_bytes(P))
        self.workspace = torch.zeros((self.ws_bytes + 7) // 8, dtype=torch.int64, device=d)
        self.stats = torch.zeros(max(self.T, 1), _lib.STATS_LEN, dtype=torch.float64, device=d)
        self.cur = 0
        self.iteration = 0
        self.pending_gather = False

    # ------------------------------------------------------------------ plumbing
    def _exchange_buffer(self, mine):
        """All-gather the IPC handle of `mine`; returns a device array of world-size base pointers."""
        handles = [None] * self.world
        dist.all_gather_object(handles, mine.handle(), group=self.group)
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(mine.ptr.value)
            else:
                out = C.c_void_p()
                _lib.check(self.lib.cdw_ipc_open_handle((C.c_ubyte * 64).from_buffer_copy(h), C.byref(out)), "cdw_ipc_open_handle")
                self._opened.append(out.value)
                ptrs.append(out.value)
        return torch.tensor(ptrs, dtype=torch.int64, device=self.device)

    def _exchange(self, name, b):
        return self._exchange_buffer(self._bufs[name][b])

    def local(self, name, b=None):
        return self._bufs[name][self.cur if b is None else b].tensor

    @property
    def beta(self):
        return 1.0 / self.langevin.kT

    def peer_barrier(self, t=None):
        """Every rank's launches so far (this stream) are complete before any rank continues (cdw_peer_barrier; bounded wait).
        In front of the launch of iteration t the wait is skipped on the device when neither iteration t - 1 nor t - 2 resampled
        (all ancestors involved are the identity: nobody touches a peer's rows)."""
        a = b = None
        if t is not None and t >= 2:
            a = self.stats[t - 1, _lib.STAT_RESAMPLE:]
            b = self.stats[t - 2, _lib.STAT_RESAMPLE:]   # (row 0 is never written: zero, the initial ancestors are the identity)
        if os.environ.get("CDW_DIAG_NO_BARRIER_WAIT"):   # timing diagnostic only (results are then unordered across ranks)
            a = b = self.stats[0, _lib.STAT_RESAMPLE:]
        _lib.check(self.lib.cdw_peer_barrier(_lib.ptr(self._peer_flags), C.c_void_p(self._flags.ptr.value), self.rank, self.world,
                                            _lib.ptr(self._epoch_counter), _lib.ptr(self._barrier_status), _lib.ptr(a), _lib.ptr(b),
                                            _lib.current_stream()), "cdw_peer_barrier")

    def check_barriers(self):
        if int(self._barrier_status.item()) != 0:
            raise _lib.CdwError("a peer did not reach a cdw_peer_barrier within its time limit: the results of this run are not valid")

    # ------------------------------------------------------------------ the loop
    def initialize(self, positions_local, velocities_local=None):
        """positions_local: this rank's block, (P/G, A, 3) host array or (P/G, A, 4) device rows."""
        rows = positions_local if torch.is_tensor(positions_local) else positions_to_rows(positions_local, self.device)
        self.cur = 0
        self.local("pos").copy_(rows, non_blocking=True)
        if velocities_local is None:
            self.local("vel").zero_()
        else:
            self.local("vel").copy_(velocities_local if torch.is_tensor(velocities_local) else positions_to_rows(velocities_local, self.device))
        if self.lookahead:
            # aux = -u_0(x_0), F(x_0; lambda_1) and inc_1 in one n_steps = 0 call (as SMCEngine does)
            prm = self.langevin.as_struct(0, n_steps=0)
            lam1 = self.lam[1] if self.T > 1 else None
            _lib.check(self.lib.cdw_smc_lookahead(
                self.system.handle, _lib.ptr(self.local("pos")), _lib.ptr(self.local("vel")), None, None, None, _lib.ptr(self.local("fs")), self.P_loc, None,
                None, _lib.ptr(self.lam[0]), _lib.ptr(lam1), C.byref(prm), C.c_uint64(0), C.c_uint64(0), C.c_int64(self.first),
                _lib.ptr(self.inc_ahead_loc[0]) if lam1 is not None else None, _lib.ptr(self.local("aux")), None, None, _lib.current_stream()),
                "cdw_smc_lookahead")
        else:
            # aux = -u_0(x_0) and the static-term cache of the initial rows, in one n_steps = 0 call
            prm = _lib.LangevinParams(0.0, 0.0, float(self.langevin.kT), 1e-6, 0, 0)
            cache = _lib.static_cache(None, None, self.local("fs"), self.local("us"))
            _lib.check(self.lib.cdw_smc_propagate_cached(
                self.system.handle, _lib.ptr(self.local("pos")), _lib.ptr(self.local("vel")), None, None, self.P_loc, None, None, None, None,
                _lib.ptr(self.lam[0]), C.byref(prm), C.c_uint64(0), C.c_uint64(0), C.c_int64(self.first), None, None, _lib.ptr(self.local("aux")), None, None,
                None, None, None, None, None, C.byref(cache), _lib.current_stream()), "cdw_smc_propagate_cached")
        self.local("label").copy_(torch.arange(self.first, self.last, dtype=torch.int32, device=self.device))
        self.cum.zero_()
        self.anc_cur = 0
        self.anc.copy_(torch.arange(self.P, dtype=torch.int32, device=self.device))
        self.wmin_key.fill_(-1)
        self.scratch_key.fill_(-1)
        self.stats.zero_()
        self.iteration = 0
        self.pending_gather = False
        self._resampled = None
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)  # every rank's initial rows are in place before any peer read

    @property
    def anc(self):
        """Ancestors still to be applied to the current rows (global ids)."""
        return self.anc_bufs[self.anc_cur]

    # ---- look-ahead loop
    def step_lookahead(self, t):
        self.iteration = t
        main = torch.cuda.current_stream()
        prev_anc = self.anc_bufs[(t - 1) % 2]
        launched = torch.cuda.Event()
        launched.record(main)
        self._side.wait_event(launched)
        with torch.cuda.stream(self._side):
            # C1: one all-gather of the 8-byte incremental works this rank's last launch left behind
            dist.all_gather_into_tensor(self.inc_ahead, self.inc_ahead_loc[(t - 1) % 2], group=self.group)
            # every rank forms the SAME weight vector and reduces it with integer sums: statistics, decision and ancestors are
            # bit-identical on all ranks and to a single-GPU run
            _lib.check(self.lib.cdw_smc_weigh_resample(
                self.kind, _lib.ptr(self.inc_ahead), _lib.ptr(prev_anc if t > 1 else None), self.P, self.threshold, _lib.ptr(self.wmin_key),
                C.c_uint64(self.resample_seed), C.c_uint64(t), _lib.ptr(self.stats[t]), _lib.ptr(self.cum), _lib.ptr(self.inc), _lib.ptr(self.W),
                _lib.ptr(self.anc_bufs[t % 2]), _lib.ptr(self.cdf), _lib.ptr(self.uniforms), _lib.ptr(self.workspace), self.ws_bytes, _lib.current_stream()),
                "cdw_smc_weigh_resample")
            resampled = torch.cuda.Event()
            resampled.record(self._side)
        if self._resampled is not None:
            main.wait_event(self._resampled)
        # the rows this launch gathers were written by the peers' launches of iteration t - 1, and the buffer it writes is the one the
        # peers' launches of iteration t - 1 read: both need every rank to be through iteration t - 1
        self.peer_barrier(t)
        c, n = self.cur, 1 - self.cur
        prm = self.langevin.as_struct(_lib.CDW_FLAG_RANDOMIZE_VELOCITIES if t == 1 else 0)
        lam_next = self.lam[t + 1] if t + 1 < self.T else None
        _lib.check(self.lib.cdw_smc_lookahead_sharded(
            self.system.handle, _lib.ptr(self.peer["pos"][c]), _lib.ptr(self.peer["vel"][c]), _lib.ptr(self.peer["fs"][c]), _lib.ptr(self.peer["label"][c]),
            self.world, self.P_loc, _lib.ptr(self.local("pos", n)), _lib.ptr(self.local("vel", n)), _lib.ptr(self.local("fs", n)), self.P_loc,
            _lib.ptr(prev_anc[self.first:self.last]), _lib.ptr(self.lam[t]), _lib.ptr(lam_next), C.byref(prm), C.c_uint64(self.seed),
            C.c_uint64(t * max(self.langevin.n_steps, 1)), C.c_int64(self.first), _lib.ptr(self.inc_ahead_loc[t % 2]) if lam_next is not None else None,
            _lib.ptr(self.local("aux", n)), _lib.ptr(self.local("label", n)), _lib.ptr(self.nan_flags), _lib.current_stream()), "cdw_smc_lookahead_sharded")
        self._resampled = resampled
        self.cur = n
        self.anc_cur = t % 2
        self.pending_gather = True

    def join(self):
        if self._resampled is not None:
            torch.cuda.current_stream().wait_event(self._resampled)
            self._resampled = None

    # ---- sequential loop (Euler-Maruyama proposal)
    def propagate(self, t):
        """The whole propagate call of iteration t (gather over peer memory, weights, move) as one launch."""
        self.iteration = t
        c, n = self.cur, 1 - self.cur
        flags = _lib.CDW_FLAG_ULA if self.proposal == "ula" else (_lib.CDW_FLAG_RANDOMIZE_VELOCITIES if t == 1 else 0)
        if self.twist is not None:
            flags |= self.twist.flags
        prm = self.langevin.as_struct(flags, twist=None if self.twist is None else self.twist.row(t))
        cache = _lib.static_cache(None, None, self.local("fs", n), self.local("us", n), self.peer["fs"][c], self.peer["us"][c])
        _lib.check(self.lib.cdw_smc_propagate_sharded_cached(
            self.system.handle, _lib.ptr(self.peer["pos"][c]), _lib.ptr(self.peer["vel"][c]), _lib.ptr(self.peer["aux"][c]), _lib.ptr(self.peer["label"][c]),
            self.world, self.P_loc, _lib.ptr(self.local("pos", n)), _lib.ptr(self.local("vel", n)), self.P_loc, _lib.ptr(self.anc[self.first:self.last]),
            _lib.ptr(self.cum[self.first:self.last]), _lib.ptr(self.lam[t]), C.byref(prm), C.c_uint64(self.seed),
            C.c_uint64(t * max(self.langevin.n_steps, 1)), C.c_int64(self.first), _lib.ptr(self.inc_loc), _lib.ptr(self.W_loc), _lib.ptr(self.local("aux", n)),
            _lib.ptr(self.local("label", n)), _lib.ptr(self.nan_flags), None, C.byref(cache), _lib.current_stream()), "cdw_smc_propagate_sharded_cached")
        self.cur = n

    def finish_iteration(self, t):
        """Weight exchange + global resampling after a propagate call; the all-gather is also the ordering point for the
        peer reads of the next iteration."""
        anc_out = self.anc_bufs[1 - self.anc_cur]
        dist.all_gather_into_tensor(self.W, self.W_loc, group=self.group)
        # every rank reduces the SAME vector: global min, log-sum-exp, nESS, decision (scratch_key is armed once in initialize();
        # the statistics kernel re-arms it after every use)
        _lib.check(self.lib.cdw_weight_update(_lib.ptr(self.W), None, None, None, self.P, None, _lib.ptr(self.W_copy), _lib.ptr(self.scratch_key),
                                             _lib.current_stream()), "cdw_weight_update")
        _lib.check(self.lib.cdw_smc_resample(self.kind, _lib.ptr(self.W), self.P, self.threshold, _lib.ptr(self.scratch_key), C.c_uint64(self.resample_seed),
                                            C.c_uint64(t), _lib.ptr(self.stats[t]), _lib.ptr(self.cum), _lib.ptr(self.cum), _lib.ptr(anc_out), _lib.ptr(self.cdf),
                                            _lib.ptr(self.uniforms), _lib.ptr(self.workspace), self.ws_bytes, _lib.current_stream()), "cdw_smc_resample")
        self.anc_cur = 1 - self.anc_cur
        self.pending_gather = True

    def step(self, t):
        if self.lookahead:
            return self.step_lookahead(t)
        self.propagate(t)
        self.finish_iteration(t)

    def materialize(self):
        """Apply the last resampling to this rank's block (peer gather over NVLink)."""
        self.join()
        if self.pending_gather:
            if self.lookahead:
                self.peer_barrier()   # every rank's last launch has written its rows
            c, n = self.cur, 1 - self.cur
            _lib.check(self.lib.cdw_gather_particles_peer(
                _lib.ptr(self.peer["pos"][c]), _lib.ptr(self.peer["vel"][c]), _lib.ptr(self.peer["aux"][c]), _lib.ptr(self.peer["label"][c]), self.world,
                self.P_loc, _lib.ptr(self.anc[self.first:self.last]), _lib.ptr(self.local("pos", n)), _lib.ptr(self.local("vel", n)),
                _lib.ptr(self.local("aux", n)), _lib.ptr(self.local("label", n)), self.P_loc, self.A, _lib.current_stream()), "cdw_gather_particles_peer")
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)  # nobody reuses the source buffers before every rank has pulled its rows
            self.cur = n
            self.pending_gather = False
        self.check_barriers()

    def run(self):
        for t in range(1, self.T):
            self.step(t)
        self.materialize()
        return self

    graph = None

    def capture(self, positions_local, velocities_local=None):
        """Capture iterations 1..T-1 (per iteration: peer barrier + the fused launch over peer memory on the main branch, NCCL
        all-gather + global resample on a side branch) into one CUDA graph per rank; every rank must call this collectively.
        `anneal` then replays it."""
        self.graph = None
        self.initialize(positions_local, velocities_local)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture (lazy loading, NCCL channel setup, cudaFuncSetAttribute)
            for t in range(1, self.T):
                self.step(t)
            self.join()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.initialize(positions_local, velocities_local)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for t in range(1, self.T):
                self.step(t)
            self.join()
        self._graph_final = (self.cur, self.pending_gather, self.iteration, self.anc_cur)
        self.graph = g
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        return self

    def anneal(self, positions_local, velocities_local=None):
        self.initialize(positions_local, velocities_local)
        if self.graph is not None:
            self.graph.replay()
            self.cur, self.pending_gather, self.iteration, self.anc_cur = self._graph_final
            self.materialize()
            return self
        return self.run()

    # ------------------------------------------------------------------ results
    def cumulative_works(self):
        return self.cum.cpu().numpy()

    def free_energy(self):
        c = self.cum
        return float(-torch.logsumexp(-c, 0) + np.log(c.numel()))

    def resampled_iterations(self):
        s = self.stats[:, _lib.STAT_RESAMPLE].cpu().numpy()
        return [t for t in range(1, self.T) if s[t] != 0.0]

    def close(self):
        """Drop the captured graph (NCCL refuses to tear a communicator down while a graph that holds its collectives is alive:
        destroy_process_group then blocks for ever), unmap the peers' buffers, then (after every rank has done so) free this rank's own."""
        import gc
        self.graph = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for ptr in self._opened:
            self.lib.cdw_ipc_close_handle(C.c_void_p(ptr))
        self._opened = []
        dist.barrier(group=self.group)
        for bufs in self._bufs.values():
            for b in bufs:
                b.free()
        self._flags.free()
