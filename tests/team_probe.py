"""Helper of tests/test_gpu_smc.py::test_other_team_sizes_run_both_loops (run in a subprocess because CDW_TEAM is read once per
process): a short anneal through SMCEngine.step with the look-ahead or the sequential loop, every output dumped to an .npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out, mode):
    from coddiwomple_b200 import testsystems as ts
    from coddiwomple_b200.engine import SMCEngine
    xml, _ = ts.alanine_dipeptide_shaped_xml()
    P = 3000
    x0 = ts.alanine_dipeptide_shaped_ensemble(P, seed=4)
    seq = ts.relative_alchemical_sequence(7)
    eng = SMCEngine(xml, P, seq, seed=8, resample_seed=9, floor_threshold=0.9999, lookahead=(mode == "lookahead"))
    eng.initialize(x0)
    for t in range(1, eng.T):
        eng.step(t)
    eng.finish()
    e = eng.ensemble
    np.savez(out, cum=e.cum.cpu().numpy(), stats=eng.stats.cpu().numpy(), anc=e.anc.cpu().numpy(), pos=e.pos[e.cur].cpu().numpy(),
             aux=e.aux[e.cur].cpu().numpy(), label=e.label[e.cur].cpu().numpy(), W=e.W.cpu().numpy())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
