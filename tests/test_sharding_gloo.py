"""N>1 path on CPU: two gloo ranks each solve their contiguous shard (with the oracle standing in for the
GPU, which this box lacks) and all_gather the per-instance statistics; the result must equal the
single-process solve instance for instance -- results do not depend on the shard count."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from altro_mpc_icra2021_b200 import sharding
    from altro_mpc_icra2021_b200.problems import quadruped
    from tests.helpers import OracleSolver

    r, w = sharding.init_distributed("gloo")
    assert (r, w) == (rank, world)
    prob, _ = quadruped.mpc_problem(11, linearized_friction=True, seed=5)  # every rank builds the global batch
    i0, i1 = sharding.shard_range(prob.B, rank, world)
    shard = prob.slice(i0, i1)
    s = OracleSolver(shard, quadruped.mpc_options(), nthreads=1).solve()
    st = s.stats
    g = sharding.gather_stats({"iterations": st.iterations, "status": st.status, "cost": st.cost, "c_max": st.c_max})
    tmax = sharding.max_over_ranks(float(rank + 1))
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), tmax=tmax, **g)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = np.load(tmp_path / "gathered.npz")
    from altro_mpc_icra2021_b200.problems import quadruped
    from tests.helpers import OracleSolver

    prob, _ = quadruped.mpc_problem(11, linearized_friction=True, seed=5)
    s = OracleSolver(prob, quadruped.mpc_options(), nthreads=2).solve()
    assert g["tmax"] == 2.0
    assert np.array_equal(g["iterations"], s.stats.iterations) and g["iterations"].dtype == s.stats.iterations.dtype
    assert np.array_equal(g["status"], s.stats.status)
    assert np.array_equal(g["cost"], s.stats.cost) and np.array_equal(g["c_max"], s.stats.c_max)
