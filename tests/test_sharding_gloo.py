"""Sample sharding on CPU: world_size-2 gloo process group.

The product's compute kernels need a GPU, but its sharding plumbing (admm_lstm_b200/comm.py: balanced
contiguous shard ranges, in-place all-reduce of the packed accumulators, global-N bookkeeping) is host code.
It is exercised here for real over gloo, with the oracle standing in for the per-shard arithmetic: every
cross-sample sum of the algorithm goes through Comm.allreduce_sum_, and the sharded run must reproduce the
unsharded one (weights bit-identical across ranks, equal to the single-process result to reduction-order
tolerance).  The same property is checked on the GPU in test_gpu_parity.py::test_sharded_fake_world_equals_single
and with NCCL by `bench.py --gpus 2`."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, variant, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from admm_lstm_b200.comm import Comm
        from helpers import GOOGLE, synthetic_problem
        from oracle.admm_oracle import OracleADMM
        comm = Comm()
        assert comm.active and comm.world_size == world and comm.rank == rank
        n, t, d, h, o = 203, 4, 3, 7, 2            # odd N: shards of 102 and 101 samples
        x, y, w = synthetic_problem(n, t, d, h, o, seed=21)
        lo, hi = comm.shard_range(n)
        assert comm.sum_int(hi - lo, "cpu") == n

        def allreduce(arr):
            tns = torch.from_numpy(np.ascontiguousarray(arr).copy())
            comm.allreduce_sum_(tns)
            return tns.numpy().reshape(np.shape(arr)).astype(arr.dtype)

        ora = OracleADMM(w, x[lo:hi], y[lo:hi], GOOGLE, variant=variant, allreduce=allreduce, n_global=n)
        for _ in range(3):
            prev = ora.snapshot_primal()
            ora.step()
        m = ora.metrics(prev)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, h=ora.gates["h"], a=ora.gates["a"],
                 objective=m["objective"], primal=m["primal_residual"], **{"w_" + k: v for k, v in ora.w.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant", ["admm", "no_dual_y"])
def test_two_rank_sharding_matches_single_process(variant, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import GOOGLE, WKEYS, rel_err, synthetic_problem
    from oracle.admm_oracle import OracleADMM
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), variant, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world))
    n, t, d, h, o = 203, 4, 3, 7, 2
    x, y, w = synthetic_problem(n, t, d, h, o, seed=21)
    ref = OracleADMM(w, x, y, GOOGLE, variant=variant)
    for _ in range(3):
        prev = ref.snapshot_primal()
        ref.step()
    m = ref.metrics(prev)
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 102, 102, 203)
    for k in WKEYS:
        assert np.array_equal(r0["w_" + k], r1["w_" + k]), k          # replicas never drift apart
        assert rel_err(r0["w_" + k], ref.w[k]) < 1e-5, k
    assert rel_err(r0["h"], ref.gates["h"][:102]) < 1e-5
    assert rel_err(r1["h"], ref.gates["h"][102:]) < 1e-5
    assert rel_err(r1["a"], ref.gates["a"][102:]) < 1e-5
    for r in (r0, r1):      # metrics are global quantities: identical on both ranks and equal to the unsharded ones
        assert abs(float(r["objective"]) - m["objective"]) < 1e-5 * abs(m["objective"])
        assert abs(float(r["primal"]) - m["primal_residual"]) < 1e-5 * m["primal_residual"]


def test_shard_ranges_are_balanced_and_cover():
    from admm_lstm_b200.comm import Comm

    class Fake(Comm):
        def __init__(self, rank, world):
            self.active, self.world_size, self.rank, self.group = True, world, rank, None

    for n in (1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            if n < world:
                continue
            edges = [Fake(r, world).shard_range(n) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
