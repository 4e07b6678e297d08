#!/usr/bin/env python
"""bench.py — particle-Langevin-steps/s and SMC anneals/s of the B200 SMC annealing hot path.

Workload (BASELINE.json configs[1]): AlanineDipeptideVacuum-shaped alchemical annealing, 10^4 particles per GPU,
100 lambda states, 1 BAOAB step per lambda, multinomial resampling at nESS <= 0.5, synthetic replicas
(coddiwomple_b200/testsystems.py).  One "step" = one full anneal (99 SMC iterations over all particles).

  python bench.py --gpus N --steps K --warmup W          # our arm (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W  # the CPU oracle port of the reference path on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for how every field is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:  # one numpy worker per core: keep BLAS/OpenMP from oversubscribing the host
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(_k, "1")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PARTICLES = 10_000
N_LAMBDA = 100
WORKLOAD = "AlanineDipeptideVacuum-shaped alchemical annealing (synthetic 22-atom perses-form hybrid), 1e4 particles/GPU, 100 lambda, 1 BAOAB step/lambda"


def algorithmic_flops(spec, n_steps=1):
    """SURVEY.md §8(d) flop convention, frozen: per energy+force evaluation
    31 N_bond + 75 N_angle + 130 N_torsion + 50 N_pair + 30 N_pair_softcore; integrator 60 A and SHAKE/RATTLE 90 N_c per step.
    One SMC iteration = 2 evaluations + n_steps integrator steps (+ 1 evaluation per extra step)."""
    c = spec.counts()
    n_sc = int((spec.pair_sc_class > 0).sum())
    f_eval = 31 * c["bonds"] + 75 * c["angles"] + 130 * c["torsions"] + 50 * c["pairs"] + 30 * n_sc
    per_step = 60 * c["atoms"] + 90 * c["constraints"]
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


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(n_particles, n_lambda, seed=0):
    """Time the oracle port (fp64 numpy restatement of the reference's loop) on a bounded sample of the workload."""
    from coddiwomple_b200 import testsystems as ts
    from oracle import smc
    from oracle.potential import XmlSystemOracle
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    o = XmlSystemOracle(xml)
    x0 = ts.alanine_dipeptide_shaped_ensemble(n_particles)
    seq = ts.relative_alchemical_sequence(n_lambda)
    t0 = time.perf_counter()
    out = smc.anneal(o, x0, seq, seed=seed, resample_seed=seed + 1)
    dt = time.perf_counter() - t0
    return n_particles * (n_lambda - 1) / dt, dt, out["free_energy"]


CPU_SAMPLE_PARTICLES = 128  # per worker and step: ~3 s of the numpy port per step on one core


def _cpu_worker(job):
    P, T, seed = job
    _, dt, fe = cpu_port_rate(P, T, seed=seed)
    return dt, fe


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference arm: the CPU restatement of the reference's loop on ALL host cores.  The reference's own engine
    (OpenMM's CPU platform) is not installable offline; the numpy port is single-threaded per anneal, so the host is
    filled with one independent anneal of CPU_SAMPLE_PARTICLES particles per core and the rates are summed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    cores = host_cores()
    P, T = CPU_SAMPLE_PARTICLES, N_LAMBDA
    rates, times = [], []
    with ProcessPoolExecutor(cores, mp_context=mp.get_context("fork")) as pool:
        for _ in range(max(args.warmup, 1)):
            list(pool.map(_cpu_worker, [(8, 10, 0)] * cores))
        for k in range(args.steps):
            t0 = time.perf_counter()
            list(pool.map(_cpu_worker, [(P, T, 1000 * k + c) for c in range(cores)]))
            dt = time.perf_counter() - t0
            times.append(dt); rates.append(cores * P * (T - 1) / dt)
    value = float(np.mean(rates))
    sample = (f"oracle/ (fp64 numpy restatement of openmm/coddiwomple.py:312-331 + OpenMM's arithmetic), per step {cores} independent anneals "
              f"(one per core) of {P} particles x {T} lambda, vectorised over particles; the reference itself cannot run: OpenMM is not installable offline")
    line = {"impl": "reference", "metric": "particle-Langevin-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": f"{cores} x {P} particles x {T} lambda per step"},
            "anneals_per_s": value / (N_PARTICLES * (N_LAMBDA - 1)),
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess():
    """Run the reference arm in a child process (a fork pool after CUDA initialisation is not safe) on a bounded sample."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"], env=env,
                         capture_output=True, text=True, timeout=600)
    for ln in reversed(out.stdout.splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError("cpu baseline failed: " + out.stderr[-2000:])


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

    xml, _ = ts.alanine_dipeptide_shaped_xml()
    spec = spec_from_xml(xml)
    seq = ts.relative_alchemical_sequence(N_LAMBDA)
    P = N_PARTICLES
    x0_host = ts.alanine_dipeptide_shaped_ensemble(P, seed=2 + rank)
    langevin = LangevinParameters(temperature=300.0, collision_rate=1.0, timestep=0.002, n_steps=1)
    if world > 1:
        from coddiwomple_b200.parallel import ShardedSMCEngine
        eng = ShardedSMCEngine(xml, P * world, seq, langevin=langevin, seed=0, resample_seed=1)
    else:
        eng = SMCEngine(xml, P, seq, langevin=langevin, seed=0, resample_seed=1)
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
        eng.capture(x0_dev)   # per rank: T-1 x (fused propagate over peer memory, NCCL all-gather of W, global resample)

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
            flush.fill_(1.0)  # L2 flush between anneals (the working set of one anneal is L2-resident by design)
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
    with ClockSampler(local) as clocks:
        ms = timed(anneal_resident, args.steps)
        ms_e2e = timed(anneal_e2e, args.steps)
    units_per_step = P * world * (N_LAMBDA - 1) * langevin.n_steps
    value = units_per_step / (np.mean(ms) * 1e-3)
    e2e_value = units_per_step / (np.mean(ms_e2e) * 1e-3)
    fe = eng.free_energy()

    # ---- roofline of the dominant kernel: the fused propagate kernel, timed with CUDA events on its stream
    flops = algorithmic_flops(spec, langevin.n_steps)
    eng.graph = None  # per-launch timing uses the eager loop
    eng.initialize(x0_dev)
    evs = []
    loop0, loop1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loop0.record()
    for t in range(1, eng.T):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world == 1:
            eng.step(t)          # ONE launch: the propagate kernel with the in-launch resample block (cdw_smc_iteration)
            e1.record()
        else:
            eng.propagate(t)
            e1.record()
            eng.finish_iteration(t)
        evs.append((e0, e1))
    loop1.record()
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    loop_ms = float(loop0.elapsed_time(loop1))  # the same eager loop end to end: the kernel's share is taken against it
    line = None
    if rank == 0:
        fp64_peak, _ = kernels.fma_peak(True, 8192)
        fp32_peak, _ = kernels.fma_peak(False, 8192)
        achieved = flops["per_iteration"] * P / (k_ms * 1e-3) / 1e12
        peaks, peaks_src = measured_peaks()
        share = k_ms * (eng.T - 1) / loop_ms
        traffic = measured_traffic()
        # ---- HBM-bound resampling kernels at sweep size (config 5), for the record
        sweep = resampling_sweep(torch, kernels, _lib, 1 << 24, spec.n_atoms if spec.n_atoms < 14 else 13) if world == 1 else None
        cpu_base = cpu_baseline_subprocess() if world == 1 else None
        line = {"metric": "particle-Langevin-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "particles_per_gpu": P, "parallelism": (f"particles sharded over {world} GPUs; all-gather of W + peer-mapped gather-on-load" if world > 1 else "single GPU"), "n_lambda": N_LAMBDA, "atoms": spec.n_atoms, "resampler": "multinomial@nESS<=0.5",
                           "step": "one full anneal = 99 SMC iterations (single GPU: one kernel launch each)", "l2": "flushed between anneals by a 256 MiB write", "terms": spec.counts()},
                "anneals_per_s": 1e3 / float(np.mean(ms)), "free_energy_kT": fe, "resampled_iterations": len(eng.resampled_iterations()),
                "e2e": {"value": e2e_value, "unit": "particle-steps/s", "ms_per_step": float(np.mean(ms_e2e)), "h2d_bytes_per_step": int(x0_pinned.numel() * 4),
                        "d2h_bytes_per_step": int(cum_host.numel() * 8), "api": "SMCEngine.anneal(pinned host rows) + cumulative works to host", "cuda_graph": use_graph},
                # both timed loops; single GPU: seed + ONE launch per iteration (propagate with the in-launch resample block) + final
                # gather; sharded: seed + (propagate, weight_update, fused resample or its 7-kernel form above 24576 particles) per
                # iteration + final peer gather
                "gpu_launches": int(args.steps * 2 * (1 + (1 if world == 1 else (3 if P * world <= 24576 else 9)) * (N_LAMBDA - 1) + 1)),
                "roofline": {"bound": "fp64", "kernel": "smc_propagate_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": achieved / fp64_peak, "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                             "avg_launch_ms": k_ms, "share_of_step": share,
                             "algorithmic_flop_per_particle_iteration": flops["per_iteration"],
                             "peak_source": "cdw_fma_peak(fp64) micro-benchmark run in this process (MEASURED_PEAKS.json has no FP pipe figure)",
                             "fp32_fma_peak_tflops": fp32_peak, "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peaks_src},
                "resampling_sweep": sweep,
                "cpu_baseline": cpu_base,
                "clocks": clocks.summary()}
        print(json.dumps(line))
    if world > 1:
        eng.close()
        dist.destroy_process_group()


def resampling_sweep(torch, kernels, _lib, n, A):
    """GB/s of the HBM-bound kernels at n particles: K2 stats, K3a scan (3 kernels) + search, K3b gather (pos+vel rows)."""
    lib = _lib.load()
    g = torch.Generator("cuda").manual_seed(5)
    W = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    u = kernels.philox_uniforms(6, 0, n)
    ws = kernels.WeightWorkspace(n)
    anc = torch.empty(n, dtype=torch.int32, device="cuda")
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
        _lib.check(lib.cdw_weight_update(_lib.ptr(W), None, None, None, n, None, _lib.ptr(ws.cdf.view(torch.float64)), _lib.ptr(ws.wmin_key), st))
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
    return {"n": n, "weights_pass12_gbs": gb(n * 40.0, ms_stats), "scan_search_systematic_gbs": gb(n * (16.0 + 12.0), ms_sys),
            "scan_search_multinomial_gbs": gb(n * (16.0 + 20.0), ms_mul), "gather_n": nP,
            "gather_gbs": gb(nP * (2 * 2 * A * 16.0 + 4 + 16 + 8), ms_g), "ms": {"stats": ms_stats, "systematic": ms_sys, "multinomial": ms_mul, "gather": ms_g},
            "bytes_model": "SURVEY.md 8(d): update+stats 40 B, scan 16 B, search 12-20 B, gather 2*(2*A*16)+28 B per particle"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="run the anneal as individual launches instead of one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
