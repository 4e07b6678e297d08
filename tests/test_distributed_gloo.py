"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: sharding arithmetic, the replicated-weights protocol
(all-gather of the per-rank weight slices -> identical statistics / ancestors on every rank and identical to one
process), and the particle-migration plan.  The kernels themselves are exercised on GPUs by tests/multi_gpu_check.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coddiwomple_b200 import parallel
from oracle import resampling, rng, weights


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, P, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, last = parallel.shard_bounds(P, world, rank)
        r = np.random.RandomState(7)
        cum_all, inc_all = r.normal(0, 1.0, P), r.normal(0.3, 1.5, P)   # what a single process would hold
        w_loc = torch.from_numpy((cum_all + inc_all)[first:last].copy())
        W = torch.empty(P, dtype=torch.float64)
        dist.all_gather_into_tensor(W, w_loc)                           # C1 of the sharded engine
        W = W.numpy()
        u = rng.uniforms53(3, P, stream=5)                              # same counter-based stream on every rank
        ness = weights.nESS(W)
        anc = resampling.ancestors_fixed(W, u)
        anc_sys = resampling.ancestors(W, None, kind="systematic", u0=float(rng.uniforms53(3, 1, stream=5)[0]))
        my_rows = anc[first:last]
        src_rank, src_local = parallel.owner_of(my_rows, P // world)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), W=W, ness=ness, anc=anc, anc_sys=anc_sys, src_rank=src_rank, src_local=src_local,
                 m=parallel.migration_matrix(anc, world), m_sys=parallel.migration_matrix(anc_sys, world))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharding_arithmetic():
    assert parallel.shard_bounds(1000, 4, 0) == (0, 250) and parallel.shard_bounds(1000, 4, 3) == (750, 1000)
    with pytest.raises(ValueError):
        parallel.shard_bounds(1001, 4, 0)
    rk, loc = parallel.owner_of([0, 249, 250, 999], 250)
    assert list(rk) == [0, 0, 1, 3] and list(loc) == [0, 249, 0, 249]
    m = parallel.migration_matrix(np.array([0, 0, 0, 3]), 2)
    assert m.tolist() == [[2, 0], [1, 1]] and m.sum() == 4


def test_replicated_weights_protocol_world2(tmp_path):
    P, world = 4096, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, P, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{k}.npz") for k in range(world))
    rs = np.random.RandomState(7)
    W_single = rs.normal(0, 1.0, P) + rs.normal(0.3, 1.5, P)
    u = rng.uniforms53(3, P, stream=5)
    for r in (r0, r1):
        np.testing.assert_array_equal(r["W"], W_single)                  # the all-gather reassembles the global vector
        assert float(r["ness"]) == weights.nESS(W_single)                # identical statistics on every rank
        np.testing.assert_array_equal(r["anc"], resampling.ancestors_fixed(W_single, u))   # == the single-process ancestors
    np.testing.assert_array_equal(r0["m"], r1["m"])
    # multinomial ancestors are unsorted: about half of the rows cross the link; systematic ones are sorted: only the
    # weight imbalance between the shards does
    m, ms = r0["m"], r0["m_sys"]
    assert m.sum() == P and ms.sum() == P
    assert (m[0, 1] + m[1, 0]) > 0.3 * P
    assert (ms[0, 1] + ms[1, 0]) < (m[0, 1] + m[1, 0])
    # each rank's pull plan addresses valid rows of the owning rank
    for k, r in enumerate((r0, r1)):
        assert r["src_rank"].min() >= 0 and r["src_rank"].max() < world and r["src_local"].max() < P // world
        np.testing.assert_array_equal(r["src_rank"] * (P // world) + r["src_local"], r["anc"][k * P // world:(k + 1) * P // world])
