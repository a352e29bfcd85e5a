"""Host-side logic of the multi-GPU path, exercised with world_size 2 on CPU (gloo): env sharding and the gradient
all-reduce identity the data-parallel learner relies on (equal shards: mean of shard means == global mean)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import td3_oracle as to


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import rtd3_b200 as rt
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, L, B = 32, 2, 64
        rs = np.random.RandomState(0)                      # same weights and data on every rank
        w = to.kaiming_uniform_params(rs, 4, H, L, 1)
        x = rs.uniform(-1, 1, (B, 4)).astype(np.float32)
        y = rs.uniform(-1, 1, (B, 1)).astype(np.float32)
        lo, hi = rt.shard_range(B, rank, world)
        q_, acts = to.mlp_forward(w, x[lo:hi], 4, H, L, 1)
        g_local, _ = to.mlp_backward(w, acts, 2.0 * (q_ - y[lo:hi]) / (hi - lo), 4, H, L, 1)   # gradient of the SHARD mean
        flat = torch.from_numpy(g_local.copy())
        world_seen = rt.trainer.allreduce_grads_(flat)
        g_dp = flat.numpy() / world_seen                   # what the optimiser kernel's grad_scale applies
        qf, af = to.mlp_forward(w, x, 4, H, L, 1)
        g_full, _ = to.mlp_backward(w, af, 2.0 * (qf - y) / B, 4, H, L, 1)
        q.put((rank, world_seen, float(np.abs(g_dp - g_full).max()), float(np.abs(g_full).max()), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    import rtd3_b200 as rt
    assert [rt.shard_range(65536, r, 8) for r in range(8)] == [(r * 8192, (r + 1) * 8192) for r in range(8)]
    with pytest.raises(ValueError):
        rt.shard_range(10, 0, 4)


def test_allreduce_without_process_group_is_identity():
    import rtd3_b200 as rt
    t = torch.ones(5)
    assert rt.trainer.allreduce_grads_(t) == 1 and torch.equal(t, torch.ones(5))


@pytest.mark.timeout(120)
def test_gradient_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    ranges = sorted(r[4] for r in res)
    assert ranges == [(0, 32), (32, 64)]
    for rank, seen, err, scale, _ in res:
        assert seen == 2
        assert err <= 1e-6 * max(1.0, scale)


def _due_worker(rank, world, port, q):
    import rtd3_b200 as rt
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = []
        # (local counts per rank, episodes_per_update): ranks disagree locally, the collective decision must be the same
        for counts, thr in (((5, 0), 4), ((5, 4), 4), ((0, 0), 1), ((3, 5), 4), ((8, 0), 4)):
            out.append(rt.trainer.update_due(torch.tensor([counts[rank]], dtype=torch.int32), thr))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_update_decision_is_collective_world2_gloo():
    """The full loop's `td3_update` all-reduces gradients, so every rank must take the same update / no-update decision in the
    same tick (a rank-local decision deadlocked the 2-GPU run): SUM of the finished-episode counters against threshold x world."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_due_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res[0] == res[1] == [False, True, False, True, True]


def test_update_decision_single_rank():
    import rtd3_b200 as rt
    assert rt.trainer.update_due(torch.tensor([3]), 3) and not rt.trainer.update_due(torch.tensor([2]), 3)
