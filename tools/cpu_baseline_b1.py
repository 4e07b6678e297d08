"""CPU baseline "B1" (SURVEY.md §8d): the reference's OWN importable classes — Particle, MultinomialResampler, TargetFactory,
ProposalFactory, nESS, vanilla_floor_threshold (coddiwomple/{particles,resamplers,distribution_factories,utils}.py) — executing the
hot loop of mcmc_smc_resampler (coddiwomple/openmm/coddiwomple.py:312-331) verbatim, one particle at a time, with OpenMM's
arithmetic (which cannot be installed offline) replaced by the fp64 numpy oracle behind the reference's duck-typed PDFState /
Propagator interfaces.  This is the reference's control flow and per-particle Python overhead; it runs only where /root/reference
exists (the build container), single-threaded like the reference.  Result -> profiles/r2_cpu_baseline_b1.json.

    PYTHONPATH=/root/reference python tools/cpu_baseline_b1.py [P] [T]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402

from coddiwomple.distribution_factories import ProposalFactory, TargetFactory  # noqa: E402  (the reference)
from coddiwomple.particles import Particle  # noqa: E402
from coddiwomple.resamplers import MultinomialResampler  # noqa: E402
from coddiwomple.utils import nESS, vanilla_floor_threshold  # noqa: E402

from coddiwomple_b200 import testsystems as ts  # noqa: E402
from coddiwomple_b200.system import load_xml  # noqa: E402
from oracle import langevin  # noqa: E402
from oracle.potential import XmlSystemOracle, kT_kj_per_mol  # noqa: E402


class State:
    def __init__(self, x, v=None):
        self.positions, self.velocities = x, (np.zeros_like(x) if v is None else v)


class OraclePDFState:
    """PDFState (coddiwomple/states.py:17-53) over the numpy oracle: what OpenMMPDFState (openmm/states.py:45-136) does with a Context."""

    def __init__(self, oracle):
        self.o, self.lam, self.kT = oracle, {}, kT_kj_per_mol()

    def set_parameters(self, p):
        self.lam = dict(p)

    def get_parameters(self):
        return dict(self.lam)

    def reduced_potential(self, state):
        return float(self.o.reduced_potential(state.positions[None], self.lam)[0])


class OraclePropagator:
    """Propagator (coddiwomple/propagators.py:14-39) with OMMBIP.apply's signature (openmm/propagators.py:89-96): one BAOAB step."""

    def __init__(self, pdf_state, seed=0):
        self.pdf_state, self.seed, self.calls = pdf_state, seed, 0

    def apply(self, state, n_steps=1, randomize_velocities=False, apply_pdf_to_context=False, **kw):
        o, lam = self.pdf_state.o, dict(self.pdf_state.lam)
        self.calls += 1
        out = langevin.baoab(lambda xx: o.energy_forces(xx, lam), state.positions[None], state.velocities[None], o.masses, o.constraints, 0.002, 1.0,
                             self.pdf_state.kT, n_steps, seed=self.seed, gids=np.array([self.calls]), step0=self.calls, randomize_velocities=randomize_velocities)
        return State(out["x"][0], out["v"][0]), 0.


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    golden = os.path.join(ROOT, "tests", "golden")
    oracle = XmlSystemOracle(load_xml(os.path.join(golden, "systems", "fluorobenzene.vacuum.xml.gz")))
    frames = np.load(os.path.join(golden, "fluorobenzene_cache.npy")).astype(np.float64)
    parameter_sequence = ts.relative_alchemical_sequence(T)
    np.random.seed(0)
    proposal_pdf_state = target_pdf_state = OraclePDFState(oracle)   # the reference's propagator and factories share the pdf_state (test_openmm.py:260)
    propagator = OraclePropagator(proposal_pdf_state)
    proposal_factory = ProposalFactory(parameter_sequence, propagator)
    target_factory = TargetFactory(target_pdf_state, parameter_sequence, termination_parameters=None)
    resampler = MultinomialResampler()
    particles = [Particle(index=idx) for idx in range(P)]
    for idx in range(P):
        proposal_factory.equip_initial_sample(particle=particles[idx], initial_particle_state=State(frames[np.random.choice(range(len(frames)))]), generation_pdf=None)
    t0 = time.perf_counter()
    while True:  # coddiwomple/openmm/coddiwomple.py:312-331, verbatim (reporter calls dropped)
        [particle.update_iteration() for particle in particles]
        incremental_works = np.array([target_factory.compute_incremental_work(particle, neglect_proposal_work=True) for particle in particles])
        randomize_velocities = True if particles[0].iteration == 1 else False
        [proposal_factory.propagate(particle, randomize_velocities=randomize_velocities, apply_pdf_to_context=True) for particle in particles]
        resampler.resample(particles, incremental_works, observable=nESS, threshold=vanilla_floor_threshold, update_particle_indices=True)
        if all(target_factory.terminate(particle) for particle in particles):
            break
        [particle.zero_auxiliary_work() for particle in particles]
        [particle.update_auxiliary_work(-target_factory.pdf_state.reduced_potential(particle.state)) for particle in particles]
    dt = time.perf_counter() - t0
    cum = np.array([p.cumulative_work for p in particles])
    fe = float(-np.log(np.mean(np.exp(-(cum - cum.min())))) + cum.min())
    out = {"kind": "reference control flow (coddiwomple Particle / MultinomialResampler / TargetFactory / ProposalFactory, openmm/coddiwomple.py:312-331 verbatim) "
                   "over the fp64 numpy oracle PDFState / Propagator, one particle at a time",
           "value": P * (T - 1) / dt, "unit": "particle-steps/s", "cores": 1, "particles": P, "n_lambda": T, "seconds": dt, "free_energy_kT": fe,
           "system": "fluorobenzene vacuum hybrid (13 atoms)", "where": "build container (the reference is not present on the GPU box)",
           "cpu": open("/proc/cpuinfo").read().split("model name")[1].split(":")[1].split("\n")[0].strip()}
    print(json.dumps(out))
    with open(os.path.join(ROOT, "profiles", "r2_cpu_baseline_b1.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
