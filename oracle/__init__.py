"""CPU oracle for the SMC annealing hot path — TEST INFRASTRUCTURE ONLY.

fp64 numpy restatement of the reference's algorithm for the path named in BASELINE.json:
reduced potential -> incremental work / log-weight -> ESS -> resample -> BAOAB Langevin propagate
(reference loop: coddiwomple/openmm/coddiwomple.py:312-331).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and there only as the checker (or the timed CPU baseline), never as the product path.  The product
package `coddiwomple_b200` never imports it and fails loudly when its CUDA library is missing.

Pinning status (see tests/test_oracle_golden.py):
  * potential.py    pinned by the four deterministic numbers the reference's own OpenMM run logged
                    (troubleshooting/openmm_smc.ipynb JSON lines 260/264/268/339) at <= 5e-6 relative
                    (the reference ran OpenMM's single-precision CPU platform; the fp64 restatement sits 1-2e-6 away).
  * weights.py / resampling.py   pinned bit-for-bit against the reference's importable generic layer
                    (coddiwomple/{utils,resamplers,particles}.py) via tests/golden/generic_layer.json.
  * langevin.py     parity unpinned against the reference's OpenMM trajectories (OpenMM absent, different RNG);
                    pinned statistically by the harmonic-oscillator analytics the reference tests use
                    (coddiwomple/tests/test_openmm.py:192-221) and by the reference's free-energy band
                    (coddiwomple/tests/legacy_tests.py:90-94).
  * systematic resampling and the synthetic alanine-dipeptide-shaped system have no artefact in the reference:
                    parity unpinned (textbook definition / fp64 oracle only).
"""
