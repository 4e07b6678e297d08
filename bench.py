#!/usr/bin/env python
"""bench.py — particle-Langevin-steps/s and SMC anneals/s of the B200 SMC annealing hot path.

Workload (BASELINE.json configs[2], the largest single-GPU configuration): perses-style vacuum ligand mutation — the
benzene -> fluorobenzene hybrid of coddiwomple/data/perses_data (13 atoms), 10^5 particles per GPU drawn from the reference's
lambda = 0 cache, the reference's alchemical protocol (coddiwomple/tests/legacy_tests.py:58-76) on 100 lambda states, 1 BAOAB step
per lambda, multinomial resampling at EVERY iteration (the reference's observable=None mode: all 99 launches gather random
ancestor rows, over NVLink on several GPUs).  One "step" = one full anneal (99 SMC iterations over all particles).  Multi-GPU runs keep
10^5 particles PER GPU (weak scaling).  BASELINE configs[1] (10^4 x 22-atom AlanineDipeptideVacuum shape) is timed beside it and
reported under "config2"; on 8 GPUs configs[3] (10^6 particles, BAOAB and Euler-Maruyama proposals) under "config4".

  python bench.py --gpus N --steps K --warmup W          # our arm (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W  # the CPU arm: the same loop compiled for the host, all cores

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for how every field is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PARTICLES = 100_000          # per GPU
N_LAMBDA = 100
FLOOR_THRESHOLD = -1.0         # < 0: resample at EVERY iteration — the reference's observable=None / threshold=None mode (resamplers.py:90-92).
#                                With vanilla_floor_threshold's 0.5 (utils.py:67-72) this workload never resamples (min nESS over the anneal is
#                                0.99995): the hard case is timed, the 0.5 case is reported beside it ("threshold_0.5")
CPU_SAMPLE_PARTICLES = 10_000  # the CPU arm times a bounded sample of the same workload: this many of the 10^5 particles, all 100 lambda
WORKLOAD = ("perses-style vacuum ligand mutation (benzene->fluorobenzene hybrid, 13 atoms, reference fixture), 1e5 particles/GPU, 100 lambda "
            "(legacy_tests.py:58-76 protocol), 1 BAOAB step/lambda, multinomial resampling at every iteration (observable=None, resamplers.py:90-92)")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def workload_inputs(P, seed):
    """(xml, x0 host (P, 13, 3) float32, parameter sequence) of configs[2]; particles drawn with replacement from the 151 lambda = 0 frames."""
    from coddiwomple_b200 import testsystems as ts
    from coddiwomple_b200.system import load_xml
    xml = load_xml(os.path.join(GOLDEN, "systems", "fluorobenzene.vacuum.xml.gz"))
    frames = np.load(os.path.join(GOLDEN, "fluorobenzene_cache.npy"))
    x0 = frames[np.random.RandomState(3 + seed).randint(0, len(frames), P)]
    return xml, x0, ts.relative_alchemical_sequence(N_LAMBDA)


def algorithmic_flops(spec, n_steps=1):
    """SURVEY.md §8(d) flop convention, frozen: per energy+force evaluation
    31 N_bond + 75 N_angle + 130 N_torsion + 50 N_pair + 30 N_pair_softcore; integrator 60 A and SHAKE/RATTLE 90 N_c per step.
    One SMC iteration = 2 evaluations + n_steps integrator steps (+ 1 evaluation per extra step)."""
    with open(os.path.join(ROOT, "bench", "roofline.json")) as fh:   # the frozen constants
        k = json.load(fh)
    t, st = k["flop_per_term"], k["flop_per_step"]
    c = spec.counts()
    n_sc = int((spec.pair_sc_class > 0).sum())
    f_eval = t["bond"] * c["bonds"] + t["angle"] * c["angles"] + t["torsion"] * c["torsions"] + t["pair"] * c["pairs"] + t["pair_softcore_extra"] * n_sc
    per_step = st["per_atom"] * c["atoms"] + st["per_constraint"] * c["constraints"]
    return dict(f_eval=f_eval, per_iteration=(1 + n_steps) * f_eval + n_steps * per_step)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic():
    """DRAM bytes per launch of the propagate kernel from the committed `ncu --set full` capture (tools/ncu_summary.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "propagate_traffic.json")) as fh:
            return json.load(fh)
    except OSError:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ CPU arm
def run_reference(args):
    """The reference arm: the hot loop compiled for the host (oracle/cpu_port: C++ / OpenMP over particles, every host core), on a
    bounded sample of the same workload.  The reference's own engine (OpenMM's CPU platform) is not installable offline, so `kind` is
    "port": the per-particle arithmetic is this library's per-replica code built by g++, the loop is the reference's."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from coddiwomple_b200.system import spec_from_xml
    from oracle import cpu_port
    cores = cpu_port.use_all_cores()   # (torchrun exports OMP_NUM_THREADS=1 to its ranks)
    P = CPU_SAMPLE_PARTICLES
    xml, x0, seq = workload_inputs(P, 0)
    port = cpu_port.CpuPort(spec_from_xml(xml))
    for _ in range(max(args.warmup, 1)):
        port.anneal(x0[:256], seq[:10], floor_threshold=FLOOR_THRESHOLD)
    times, fe = [], None
    for k in range(args.steps):
        t0 = time.perf_counter()
        out = port.anneal(x0, seq, floor_threshold=FLOOR_THRESHOLD, seed=k, resample_seed=1000 + k)   # (negative threshold: resample always)
        times.append(time.perf_counter() - t0)
        fe = out["free_energy"]
    value = P * (N_LAMBDA - 1) / float(np.mean(times))
    sample = (f"oracle/cpu_port (the loop of openmm/coddiwomple.py:312-331 in C++/OpenMP around the library's per-replica code built by g++ -O3 -march=native), "
              f"{cores} threads, per step one anneal of {P} of the 1e5 particles x {N_LAMBDA} lambda; the reference itself cannot run: OpenMM is not installable offline")
    line = {"impl": "reference", "metric": "particle-Langevin-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": f"{P} of the 1e5 particles x {N_LAMBDA} lambda per step", "same_config": True},
            "anneals_per_s": value / (N_PARTICLES * (N_LAMBDA - 1)), "free_energy_kT": fe,
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess():
    """B2 (compiled port, all cores) on a bounded sample, in a child process; B1 (the reference's own Python classes driving the numpy
    oracle, one particle at a time) is quoted from the committed measurement made where /root/reference exists."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"], env=env,
                         capture_output=True, text=True, timeout=900)
    base = None
    for ln in reversed(out.stdout.splitlines()):
        if ln.startswith("{"):
            base = json.loads(ln)["cpu_baseline"]
            break
    if base is None:
        raise RuntimeError("cpu baseline failed: " + out.stderr[-2000:])
    try:
        with open(os.path.join(ROOT, "profiles", "r2_cpu_baseline_b1.json")) as fh:
            b1 = json.load(fh)
        base["b1_reference_python_loop"] = {k: b1[k] for k in ("value", "unit", "cores", "kind", "particles", "n_lambda", "where")}
    except (OSError, KeyError):
        pass
    base["b0_reference_notebook"] = {"value": 3.0e2, "unit": "particle-steps/s", "source": "troubleshooting/troubleshoot.ipynb JSON 137: 10 x 500 OpenMM-CPU steps in 16 s (quoted, not re-measured)"}
    return base


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from coddiwomple_b200 import _lib, kernels, testsystems as ts
    from coddiwomple_b200.engine import LangevinParameters, SMCEngine, positions_to_rows
    from coddiwomple_b200.system import spec_from_xml

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    P = N_PARTICLES
    xml, x0_host, seq = workload_inputs(P, rank)
    spec = spec_from_xml(xml)
    langevin = LangevinParameters(temperature=300.0, collision_rate=1.0, timestep=0.002, n_steps=1)
    thr = float(args.threshold)
    always = thr < 0.0
    if world > 1:
        from coddiwomple_b200.parallel import ShardedSMCEngine
        eng = ShardedSMCEngine(xml, P * world, seq, langevin=langevin, seed=0, resample_seed=1, floor_threshold=thr, always_resample=always)
    else:
        eng = SMCEngine(xml, P, seq, langevin=langevin, seed=0, resample_seed=1, floor_threshold=thr, always_resample=always)
    x0_pinned = positions_to_rows(x0_host, device=None, pinned=True)
    x0_dev = x0_pinned.cuda()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    cum_dev = eng.cum if world > 1 else eng.ensemble.cum          # sharded runs keep the (8-byte) weights replicated
    cum_host = torch.empty(cum_dev.numel(), dtype=torch.float64, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph
    if use_graph and world == 1:
        eng.load(x0_dev)
        eng.capture()   # the whole anneal (no host decisions inside) as one CUDA graph
    elif use_graph:
        eng.capture(x0_dev)   # per rank: T-1 x (peer barrier + launch over peer memory | all-gather + global resample on a side branch)

    def anneal_resident():
        eng.anneal(x0_dev)

    def anneal_e2e():
        # the call a user makes: host positions in, cumulative works (the free-energy estimate) out
        dev = x0_pinned.to("cuda", non_blocking=True)
        eng.anneal(dev)
        cum_host.copy_(cum_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps):
        ms = []
        for _ in range(steps):
            flush.fill_(1.0)  # L2 flush between anneals (one iteration's rows are 62 MB: L2-sized)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t.item()))
        return ms

    for _ in range(max(args.warmup, 3)):
        anneal_resident()
    barrier()
    for _ in range(max(args.warmup, 3)):   # the end-to-end call has its own first-use costs (staging buffer of the device copy, pinned read-back)
        anneal_e2e()
    barrier()
    with ClockSampler(local) as clocks:
        ms = timed(anneal_resident, args.steps)
        ms_e2e = timed(anneal_e2e, args.steps)
    units_per_step = P * world * (N_LAMBDA - 1) * langevin.n_steps
    value = units_per_step / (np.mean(ms) * 1e-3)
    e2e_value = units_per_step / (np.mean(ms_e2e) * 1e-3)
    fe = eng.free_energy()
    n_resampled = len(eng.resampled_iterations())

    # ---- the same anneal with the reference's default rule (nESS <= 0.5: this workload then never resamples), for the record
    thr_half = None
    if always:
        eng.threshold = 0.5
        if use_graph:
            if world == 1:
                eng.load(x0_dev); eng.capture()
            else:
                eng.capture(x0_dev)
        for _ in range(2):
            anneal_resident()
        ms_half = timed(anneal_resident, max(args.steps // 4, 3))
        thr_half = {"ms_per_step": float(np.mean(ms_half)), "value": units_per_step / (np.mean(ms_half) * 1e-3), "resampled_iterations": len(eng.resampled_iterations())}

    # ---- parity probe (outside the timed region): the multi-GPU anneal against the single-GPU one, or on one GPU the look-ahead loop
    # against the sequential loop, bit for bit
    probe = parity_probe(torch, dist, world, rank, xml)

    # ---- extra workloads, rank-collective ones first
    config4 = config4_runs(torch, dist, world, rank, xml, langevin) if world == 8 else None
    sharded_sweep = sharded_resampling(torch, dist, _lib, world, rank, spec.n_atoms) if world > 1 else None

    line = None
    if rank == 0:
        flops = algorithmic_flops(spec, langevin.n_steps)
        fp64_peak, _ = kernels.fma_peak(True, 8192)
        fp32_peak, _ = kernels.fma_peak(False, 8192)
        # the dominant kernel runs once per iteration and nothing else is on the critical path: the whole graph-timed anneal divided
        # by its T - 1 launches is an UPPER bound of the launch duration (it also holds the equip launch, the final gather and, on
        # several GPUs, the peer barriers)
        k_ms = float(np.mean(ms)) / (N_LAMBDA - 1)
        achieved = flops["per_iteration"] * P / (k_ms * 1e-3) / 1e12
        peaks, peaks_src = measured_peaks()
        traffic = measured_traffic()
        eager = eager_launch_ms(torch, SMCEngine, xml, P, seq, langevin, x0_dev, thr, always) if world == 1 else None
        config2 = config2_run(torch, ts, SMCEngine, LangevinParameters, positions_to_rows, spec_from_xml, kernels, fp64_peak, args) if world == 1 else None
        sweep = resampling_sweep(torch, kernels, _lib, 1 << 24, 13, peaks.get("hbm_gbs")) if world == 1 else None
        cpu_base = cpu_baseline_subprocess() if world == 1 else None
        per_iter = 1 + (5 if eng.kind == _lib.CDW_RESAMPLE_MULTINOMIAL else 4) + (1 if world > 1 else 0)
        line = {"metric": "particle-Langevin-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "particles_per_gpu": P, "n_lambda": N_LAMBDA, "atoms": spec.n_atoms, "floor_threshold": thr,
                           "parallelism": (f"particles sharded over {world} GPUs; per iteration: peer barrier + ONE launch gathering ancestors over NVLink; all-gather of "
                                           f"incremental works + global resampling on a side stream" if world > 1 else "single GPU"),
                           "step": "one full anneal = 99 SMC iterations (one propagate launch each; the resampling of iteration t runs beside launch t on a side stream)",
                           "l2": "flushed between anneals by a 256 MiB write; one iteration's rows (62 MB) are L2-sized", "terms": spec.counts()},
                "anneals_per_s": 1e3 / float(np.mean(ms)), "free_energy_kT": fe, "resampled_iterations": n_resampled,
                "e2e": {"value": e2e_value, "unit": "particle-steps/s", "ms_per_step": float(np.mean(ms_e2e)), "h2d_bytes_per_step": int(x0_pinned.numel() * 4),
                        "d2h_bytes_per_step": int(cum_host.numel() * 8), "api": "SMCEngine.anneal(pinned host rows) + cumulative works to host", "cuda_graph": use_graph,
                        "ms_each": [round(v, 3) for v in ms_e2e]},
                "ms_each": [round(v, 3) for v in ms],
                # both timed loops; per anneal: equip + (T - 1) x (look-ahead launch + weight compose + statistics + tile sums + tile scan [+ lookup]
                # [+ peer barrier]) + final gather [+ its barrier]
                "gpu_launches": int(args.steps * 2 * (1 + per_iter * (N_LAMBDA - 1) + (2 if world > 1 else 1))),
                "threshold_0.5": thr_half,
                "parity_probe": probe,
                "roofline": {"bound": "fp64", "kernel": "smc_lookahead_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": achieved / fp64_peak, "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                             "avg_launch_ms": k_ms, "avg_launch_ms_is": "ms_per_step / (n_lambda - 1): upper bound, the graph-timed anneal holds nothing else on its critical path",
                             "avg_launch_ms_events": eager, "avg_launch_ms_events_is": "CUDA events on the launching stream around each look-ahead launch of an eager "
                             "(un-graphed) anneal, iterations 3 .. T-1, mean; the launch also waits for the previous iteration's resampling on the side stream",
                             "share_of_step": None if eager is None else min(1.0, eager * (N_LAMBDA - 1) / float(np.mean(ms))),
                             "algorithmic_hbm_bytes_per_launch": (traffic or {}).get("algorithmic_hbm_bytes_per_launch"),
                             "algorithmic_flop_per_particle_iteration": flops["per_iteration"],
                             "peak_source": "cdw_fma_peak(fp64) micro-benchmark run in this process (MEASURED_PEAKS.json has no FP pipe figure)",
                             "fp32_fma_peak_tflops": fp32_peak, "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peaks_src},
                "config2": config2, "config4": config4,
                "resampling_sweep": sweep, "resampling_sweep_sharded": sharded_sweep,
                "cpu_baseline": cpu_base,
                "clocks": clocks.summary()}
        print(json.dumps(line))
    if world > 1:
        eng.close()
        dist.destroy_process_group()


def eager_launch_ms(torch, SMCEngine, xml, P, seq, langevin, x0_dev, thr, always):
    """Mean duration of the dominant kernel's launches, measured with CUDA events on the stream it is launched on (eager loop)."""
    eng = SMCEngine(xml, P, seq, langevin=langevin, seed=0, resample_seed=1, floor_threshold=thr, always_resample=always)
    eng.initialize(x0_dev)
    evs = []
    for t in range(1, eng.T):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.step(t); e1.record()
        evs.append((e0, e1))
    eng.ensemble.join()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs[2:]]))


def parity_probe(torch, dist, world, rank, xml):
    """True iff a small anneal run two ways gives bitwise the same cumulative works, statistics and particles."""
    from coddiwomple_b200 import testsystems as ts
    from coddiwomple_b200.engine import SMCEngine
    frames = np.load(os.path.join(GOLDEN, "fluorobenzene_cache.npy"))
    seq = ts.relative_alchemical_sequence(8)
    if world == 1:
        P = 4096
        x0 = frames[np.random.RandomState(0).randint(0, len(frames), P)]
        outs = []
        for lookahead in (True, False):
            e = SMCEngine(xml, P, seq, seed=3, resample_seed=4, floor_threshold=0.9999, lookahead=lookahead)
            e.anneal(x0)
            outs.append((e.ensemble.cum.clone(), e.stats.clone(), e.ensemble.positions.clone(), e.ensemble.labels.clone()))
        same = all(torch.equal(a, b) for a, b in zip(*outs))
        return {"ok": bool(same), "what": "look-ahead loop == sequential loop, bitwise (4096 particles, 8 lambda, 7 resamplings)"}
    from coddiwomple_b200.parallel import ShardedSMCEngine
    P = 2048 * world
    x0 = frames[np.random.RandomState(0).randint(0, len(frames), P)]
    sh = ShardedSMCEngine(xml, P, seq, seed=3, resample_seed=4, floor_threshold=0.9999)
    sh.anneal(x0[sh.first:sh.last])
    parts = [torch.empty_like(sh.local("pos")) for _ in range(world)]
    dist.all_gather(parts, sh.local("pos").contiguous())
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    if rank == 0:
        single = SMCEngine(xml, P, seq, seed=3, resample_seed=4, floor_threshold=0.9999)
        single.anneal(x0)
        e = single.ensemble
        same = torch.equal(torch.cat(parts), e.pos[e.cur]) and torch.equal(sh.cum, e.cum) and torch.equal(sh.stats, single.stats)
        ok.fill_(1 if same else 0)
    dist.broadcast(ok, 0)
    sh.close()
    return {"ok": bool(int(ok.item())), "what": f"{world}-GPU sharded anneal == single-GPU anneal, bitwise ({P} particles, 8 lambda, 7 resamplings)"}


def config2_run(torch, ts, SMCEngine, LangevinParameters, positions_to_rows, spec_from_xml, kernels, fp64_peak, args):
    """BASELINE configs[1]: AlanineDipeptideVacuum-shaped annealing, 10^4 particles, 100 lambda (last round's headline workload)."""
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    P = 10_000
    rows = positions_to_rows(ts.alanine_dipeptide_shaped_ensemble(P, seed=2))
    eng = SMCEngine(xml, P, ts.relative_alchemical_sequence(N_LAMBDA), langevin=LangevinParameters(), seed=0, resample_seed=1)
    eng.load(rows)
    eng.capture()
    for _ in range(3):
        eng.anneal(rows)
    torch.cuda.synchronize()
    ms = []
    for _ in range(max(args.steps // 2, 5)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.anneal(rows); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    flops = algorithmic_flops(spec_from_xml(xml))
    k_ms = float(np.mean(ms)) / (N_LAMBDA - 1)
    achieved = flops["per_iteration"] * P / (k_ms * 1e-3) / 1e12
    return {"workload": "AlanineDipeptideVacuum-shaped alchemical annealing (synthetic 22-atom perses-form hybrid), 1e4 particles, 100 lambda, 1 BAOAB step/lambda",
            "value": P * (N_LAMBDA - 1) / (np.mean(ms) * 1e-3), "unit": "particle-steps/s", "ms_per_anneal": float(np.mean(ms)), "anneals_per_s": 1e3 / float(np.mean(ms)),
            "free_energy_kT": eng.free_energy(), "resampled_iterations": len(eng.resampled_iterations()),
            "roofline_frac_fp64": achieved / fp64_peak, "algorithmic_flop_per_particle_iteration": flops["per_iteration"]}


def config4_runs(torch, dist, world, rank, xml, langevin):
    """BASELINE configs[3]: 10^6 particles of the small-molecule vacuum system sharded over 8 GPUs (125 000 per GPU), with the MCMC
    (BAOAB) move and with the Euler-Maruyama proposal + generalised weight of the controlled-SMC scripts (troubleshooting/csmc.py:96-191)."""
    from coddiwomple_b200 import testsystems as ts
    from coddiwomple_b200.engine import LangevinParameters, positions_to_rows
    from coddiwomple_b200.parallel import ShardedSMCEngine
    frames = np.load(os.path.join(GOLDEN, "fluorobenzene_cache.npy"))
    P = 1_000_000
    seq = ts.relative_alchemical_sequence(N_LAMBDA)
    out = {}
    from coddiwomple_b200.engine import TwistSequence
    r = np.random.RandomState(9)
    twist = TwistSequence(5.0 * r.rand(N_LAMBDA - 1, 13, 3), 0.5 * r.randn(N_LAMBDA - 1, 13, 3), c=np.zeros(N_LAMBDA - 1))   # controlled SMC (csmc.py:460-747)
    G = r.randn(N_LAMBDA - 1, 39, 39) / np.sqrt(39.0)
    dense = TwistSequence.dense(2.0 * G @ np.swapaxes(G, 1, 2), 0.5 * r.randn(N_LAMBDA - 1, 39))                              # the same with full matrices A_t
    twists = {"ula_twisted": twist, "ula_twisted_dense": dense}
    for name, lang in (("baoab", langevin), ("ula", LangevinParameters(timestep=0.0005)), ("ula_twisted", LangevinParameters(timestep=0.0005)),
                       ("ula_twisted_dense", LangevinParameters(timestep=0.0005))):
        proposal = "baoab" if name == "baoab" else "ula"
        eng = ShardedSMCEngine(xml, P, seq, langevin=lang, seed=0, resample_seed=1, proposal=proposal, twist=twists.get(name))
        rows = positions_to_rows(frames[np.random.RandomState(11 + rank).randint(0, len(frames), eng.P_loc)])
        eng.capture(rows)
        for _ in range(2):
            eng.anneal(rows)
        ms = []
        for _ in range(5):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.anneal(rows); e1.record(); e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t.item()))
        out[name] = {"particles": P, "particles_per_gpu": eng.P_loc, "n_lambda": N_LAMBDA, "ms_per_anneal": float(np.mean(ms)),
                         "value": P * (N_LAMBDA - 1) / (np.mean(ms) * 1e-3), "unit": "particle-steps/s", "free_energy_kT": eng.free_energy(),
                         "resampled_iterations": len(eng.resampled_iterations())}
        eng.close()
    return out


def sharded_resampling(torch, dist, _lib, world, rank, A):
    """BASELINE configs[4] on several GPUs: one resampling step of the sharded loop at 2^22 particles per GPU — all-gather of the
    incremental works (NCCL), global weigh + resample on every rank (replicated, bit-identical), this rank's share of the peer gather."""
    import ctypes as C
    lib = _lib.load()
    n_loc = 1 << 22
    n = n_loc * world
    g = torch.Generator("cuda").manual_seed(5 + rank)
    inc_loc = torch.randn(n_loc, dtype=torch.float64, device="cuda", generator=g)
    inc = torch.empty(n, dtype=torch.float64, device="cuda")
    cum, W = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
    anc = torch.empty(n, dtype=torch.int32, device="cuda")
    cdf = torch.empty(n + 2, dtype=torch.int64, device="cuda")[:n]
    u = torch.empty(n, dtype=torch.float64, device="cuda")
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")
    key = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    ws_bytes = int(lib.cdw_weight_workspace_bytes(n))
    ws = torch.zeros((ws_bytes + 7) // 8, dtype=torch.int64, device="cuda")
    st = _lib.current_stream()
    out = {"particles": n, "particles_per_gpu": n_loc}
    for kind, name in ((0, "multinomial"), (1, "systematic")):
        def step():
            dist.all_gather_into_tensor(inc, inc_loc)
            cum.zero_()
            _lib.check(lib.cdw_smc_weigh_resample(kind, _lib.ptr(inc), None, n, -1.0, _lib.ptr(key), C.c_uint64(6), C.c_uint64(1), _lib.ptr(stats), _lib.ptr(cum), None,
                                                  _lib.ptr(W), _lib.ptr(anc), _lib.ptr(cdf), _lib.ptr(u), _lib.ptr(ws), ws_bytes, st))
        step(); torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t.item()))
        out[name] = {"ms": float(np.min(ms)), "particles_per_s": n / (np.min(ms) * 1e-3)}
        out[name].update(_peer_gather(torch, dist, _lib, lib, world, rank, n_loc, A, anc, st))
    out["gather_bytes_model"] = "per particle: pos + vel rows read (NVLink when the ancestor lives on a peer) and written, 2 x (2 x A x 16) + 28 B (SURVEY.md 8(d))"
    return out


def _peer_gather(torch, dist, _lib, lib, world, rank, n_loc, A, anc, st):
    """K3b over peer memory (cdw_gather_particles_peer): this rank's n_loc slots pull the rows of their GLOBAL ancestors from the owning ranks'
    buffers through CUDA-IPC peer-mapped pointers.  GB/s per GPU = algorithmic bytes of the rank's share / max-over-ranks time."""
    import ctypes as C
    from coddiwomple_b200.parallel import _IpcBuffer
    bufs = {k: _IpcBuffer(lib, shape, dt) for k, shape, dt in (("pos", (n_loc, A, 4), torch.float32), ("vel", (n_loc, A, 4), torch.float32),
                                                                ("aux", (n_loc,), torch.float64), ("label", (n_loc,), torch.int32))}
    opened, peers = [], {}
    for k, b in bufs.items():
        b.tensor.copy_(torch.randn(b.shape, device="cuda").to(b.dtype) if b.dtype != torch.int32 else torch.arange(n_loc, dtype=torch.int32, device="cuda"))
        handles = [None] * world
        dist.all_gather_object(handles, b.handle())
        ptrs = []
        for r, h in enumerate(handles):
            if r == rank:
                ptrs.append(b.ptr.value)
            else:
                o = C.c_void_p()
                _lib.check(lib.cdw_ipc_open_handle((C.c_ubyte * 64).from_buffer_copy(h), C.byref(o)), "cdw_ipc_open_handle")
                opened.append(o.value)
                ptrs.append(o.value)
        peers[k] = torch.tensor(ptrs, dtype=torch.int64, device="cuda")
    pos_o, vel_o = torch.empty(n_loc, A, 4, device="cuda"), torch.empty(n_loc, A, 4, device="cuda")
    aux_o, lab_o = torch.empty(n_loc, dtype=torch.float64, device="cuda"), torch.empty(n_loc, dtype=torch.int32, device="cuda")
    mine = anc[rank * n_loc:(rank + 1) * n_loc].contiguous()
    remote = float(((mine // n_loc) != rank).double().mean().item())

    def go():
        _lib.check(lib.cdw_gather_particles_peer(_lib.ptr(peers["pos"]), _lib.ptr(peers["vel"]), _lib.ptr(peers["aux"]), _lib.ptr(peers["label"]), world, n_loc,
                                                 _lib.ptr(mine), _lib.ptr(pos_o), _lib.ptr(vel_o), _lib.ptr(aux_o), _lib.ptr(lab_o), n_loc, A, st), "cdw_gather_particles_peer")
    go(); torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
    dist.barrier(); torch.cuda.synchronize()
    for o in opened:
        lib.cdw_ipc_close_handle(C.c_void_p(o))
    dist.barrier()
    for b in bufs.values():
        b.free()
    best = float(np.min(ms))
    per_gpu = n_loc * (2 * 2 * A * 16.0 + 28) / (best * 1e-3) / 1e9
    return {"gather_ms": best, "gather_gbs_per_gpu": per_gpu, "gather_gbs_total": per_gpu * world, "gather_remote_fraction": remote}


def resampling_sweep(torch, kernels, _lib, n, A, hbm_peak):
    """GB/s of the HBM-bound kernels at n particles (BASELINE configs[4] on one GPU; the full 1e3 .. 1e8 sweep is
    tools/resampling_sweep.py -> profiles/): K2 update + statistics, K3a tile sums + scan (+ emission / lookup), K3b gather of
    pos + vel rows.  Bytes per particle are SURVEY.md §8(d)'s algorithmic figures."""
    lib = _lib.load()
    g = torch.Generator("cuda").manual_seed(5)
    W = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    u = kernels.philox_uniforms(6, 0, n)
    ws = kernels.WeightWorkspace(n)
    anc = torch.empty(n, dtype=torch.int32, device="cuda")
    Wc = torch.empty_like(W)
    nP = min(n, 1 << 22)
    pos = torch.randn(nP, A, 4, device="cuda"); vel = torch.randn(nP, A, 4, device="cuda")
    pos_o, vel_o = torch.empty_like(pos), torch.empty_like(vel)
    aux = torch.randn(nP, dtype=torch.float64, device="cuda"); aux_o = torch.empty_like(aux)
    lab = torch.arange(nP, dtype=torch.int32, device="cuda"); lab_o = torch.empty_like(lab)
    anc_g = torch.randint(0, nP, (nP,), dtype=torch.int32, device="cuda")
    st = _lib.current_stream()

    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def stats():
        ws.wmin_key.fill_(-1)
        _lib.check(lib.cdw_weight_update(_lib.ptr(W), None, None, None, n, None, _lib.ptr(Wc), _lib.ptr(ws.wmin_key), st))
        _lib.check(lib.cdw_weight_stats(_lib.ptr(W), n, -1.0, _lib.ptr(ws.wmin_key), _lib.ptr(ws.stats), _lib.ptr(ws.buf), ws.bytes, st))

    def resample(kind):
        _lib.check(lib.cdw_resample_indices(kind, _lib.ptr(W), _lib.ptr(u), n, _lib.ptr(ws.stats), None, None, _lib.ptr(anc), _lib.ptr(ws.cdf), _lib.ptr(ws.buf),
                                            ws.bytes, st))

    def gather():
        _lib.check(lib.cdw_gather_particles(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(anc_g), _lib.ptr(pos_o), _lib.ptr(vel_o), _lib.ptr(aux), _lib.ptr(aux_o),
                                            _lib.ptr(lab), _lib.ptr(lab_o), nP, A, st))

    ms_stats = t(stats)
    ms_sys = t(lambda: resample(_lib.CDW_RESAMPLE_SYSTEMATIC))
    ms_mul = t(lambda: resample(_lib.CDW_RESAMPLE_MULTINOMIAL))
    ms_g = t(gather)
    gb = lambda b, ms: b / (ms * 1e-3) / 1e9
    rates = {"ess_update_stats_gbs": gb(n * 24.0, ms_stats), "scan_search_systematic_gbs": gb(n * (16.0 + 12.0), ms_sys),
             "scan_search_multinomial_gbs": gb(n * (16.0 + 20.0), ms_mul), "gather_gbs": gb(nP * (2 * 2 * A * 16.0 + 4 + 16 + 8), ms_g)}
    return dict(n=n, gather_n=nP, **rates, frac_of_hbm_peak={k[:-4]: v / hbm_peak for k, v in rates.items()} if hbm_peak else None,
                ms={"stats": ms_stats, "systematic": ms_sys, "multinomial": ms_mul, "gather": ms_g},
                bytes_model="SURVEY.md 8(d), algorithmic: update+stats 24 B, scan 16 B, search 12 (systematic) / 20 (multinomial) B, gather 2*(2*A*16)+28 B per particle")


def main():
    # a hang (a rank waiting for a peer that is gone) must not hold a GPU node until the caller's time limit: dump where every thread is
    # and leave.  CDW_WATCHDOG_S overrides the limit (0 disables it).
    import faulthandler
    limit = int(os.environ.get("CDW_WATCHDOG_S", "1500"))
    if limit > 0:
        faulthandler.dump_traceback_later(limit, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--threshold", type=float, default=FLOOR_THRESHOLD, help="resample iff nESS <= threshold (negative: at every iteration)")
    ap.add_argument("--no-graph", action="store_true", help="run the anneal as individual launches instead of one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
