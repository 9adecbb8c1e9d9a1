"""CPU tests of the multi-GPU scheduler's host logic: sharding, the gather (gloo, world_size 2 and 3), global-id
keyed initial states, and the credible-effect calls."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from sccoda_b200 import scheduler as sch


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [sch.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_initial_states_do_not_depend_on_the_sharding():
    full = sch.initial_states(10, 4, 1, 8, seed=3)
    part = sch.initial_states(4, 4, 1, 8, seed=3, dataset0=6)
    assert np.array_equal(full[6:], part)
    assert np.all(full[..., 0] == 1.0) and np.all(full[..., 8:15] == 0.0)        # sigma_d = 1, ind_raw = 0
    assert abs(full[..., 1:8].std() - 1.0) < 0.15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    comm = sch.Comm.from_env(backend="gloo")

    def compute(lo, hi):
        idx = torch.arange(lo, hi, dtype=torch.float64)
        return {"a": idx[:, None] * torch.ones(1, 3, dtype=torch.float64), "b": (idx ** 2)[:, None, None].repeat(1, 2, 2)}

    res = sch.run_sharded(n_items, compute, comm)
    mx = comm.max_float(float(rank + 1))
    sm = comm.sum_float(float(rank + 1))
    comm.barrier()
    torch.save({"a": res["a"], "b": res["b"], "mx": mx, "sm": sm}, os.path.join(out_dir, f"r{rank}.pt"))
    comm.close()


@pytest.mark.parametrize("world,n_items", [(2, 10), (2, 7), (3, 8)])
def test_gather_over_gloo(tmp_path, world, n_items):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True)
    idx = torch.arange(n_items, dtype=torch.float64)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(got["a"], idx[:, None] * torch.ones(1, 3, dtype=torch.float64))
        assert torch.equal(got["b"], (idx ** 2)[:, None, None].repeat(1, 2, 2))
        assert got["mx"] == world and got["sm"] == world * (world + 1) / 2


def test_run_sweep_refuses_more_ranks_than_datasets():
    """A rank without a dataset would have nothing to pack and no shapes for the gather: a clear error on every rank,
    before anything touches the device."""
    x, y = np.zeros((2, 6, 1)), np.ones((2, 6, 4))
    for rank in range(4):
        with pytest.raises(ValueError, match="at least one dataset per rank"):
            sch.run_sweep(x, y, 3, sch.SweepConfig(), comm=sch.Comm(rank=rank, world=4))


def test_credible_effect_calls_follow_the_reference_threshold_rule():
    inc = np.array([[1.0, 0.4, 0.0, 0.97], [0.2, 0.1, 0.0, 0.0]])
    nz = np.array([[0.5, 0.2, 0.0, -0.3], [0.1, 0.1, 0.0, 0.0]])
    thr, final = sch.credible_effect_calls(inc, nz, 0.05)
    # dataset 0: c=0.4 -> fdr=mean(0,0.6,0.03)=0.21; c=0.97 -> mean(0,0.03)=0.015 < 0.05 -> 0.97
    assert thr[0] == 0.97 and np.allclose(final[0], [0.5, 0.0, 0.0, -0.3])
    assert thr[1] == 1.0 and np.all(final[1] == 0.0)
