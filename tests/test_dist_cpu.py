"""Host-side logic of the multi-GPU path on CPU: world_size-2 `gloo` process group.  Checks the shard
partition and that the all-gather of per-rank window partials is ordered by rank (what
gmsm_ctx_finalize_device expects: partials[r][j]).  The CUDA composition itself is covered on one GPU by
tests/test_gpu_msm.py::test_window_sums_and_finalize_compose and on N GPUs by tests/test_gpu_dist.py."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    distmod = importlib.import_module("gnark-crypto_b200.dist")
    W, words = 16, 16  # 16 windows x 128-byte partials (bn254 G1)
    local = torch.full((W * words,), rank + 1, dtype=torch.int64)
    local[0] = 1000 + rank
    allp = distmod.gather_partials(local, world)
    lo, hi = distmod.shard_range(1000003, rank, world)
    q.put((rank, allp.tolist(), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_partials_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 16 * 16
    for rank, allp, lo, hi in res:
        assert len(allp) == world * n
        for r in range(world):
            assert allp[r * n] == 1000 + r and allp[r * n + 1] == r + 1  # rank-major order
    assert res[0][2:] == (0, 500001) and res[1][2:] == (500001, 1000003)


def test_shard_range_partitions():
    distmod = importlib.import_module("gnark-crypto_b200.dist")
    for n in (0, 1, 7, 1 << 24, (1 << 26) + 5):
        for world in (1, 2, 4, 8):
            cuts = [distmod.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
    with pytest.raises(ValueError):
        distmod.shard_range(10, 2, 2)


class _FakeEngine:
    def __init__(self, c, nwin):
        self.c, self.nwin, self.device = c, nwin, 0


def _plan_worker(rank, world, port, q, plans):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    distmod = importlib.import_module("gnark-crypto_b200.dist")
    try:
        distmod.ShardedMultiExp(_FakeEngine(*plans[rank]))
        q.put((rank, "ok"))
    except ValueError as e:
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("plans,ok", [([(17, 15), (17, 15)], True), ([(17, 15), (15, 17)], False)])
def test_sharded_ranks_must_share_one_window_plan(plans, ok):
    """ranks whose engines picked different window widths (uneven shards, c = 0) used to add partials of
    different plans silently; ShardedMultiExp now all-gathers (c, W) and refuses a mismatch"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, q, plans)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, msg in res:
        assert (msg == "ok") if ok else ("different window plans" in msg)


def test_window_plan_of_a_sharded_call_is_that_of_the_largest_shard():
    """dist.window_bits_for_total: one width for every rank, from (total, world) alone -- the model's choice for ceil(n / world)
    points, whatever the rank's own (possibly smaller) shard is"""
    import importlib

    distmod = importlib.import_module("gnark-crypto_b200.dist")
    L = importlib.import_module("gnark-crypto_b200._native").lib()
    for n_total, world in [(1 << 26, 4), (1 << 26, 8), ((1 << 24) + 5, 3), (1000, 2), (7, 8)]:
        want = L.gmsm_choose_window_bits(0, (n_total + world - 1) // world)
        assert distmod.window_bits_for_total("bn254_g1", n_total, world) == want
        lo, hi = distmod.shard_range(n_total, world - 1, world)
        assert hi - lo <= (n_total + world - 1) // world
    assert distmod.window_bits_for_total("bn254_g1", 1 << 26, 4) == L.gmsm_choose_window_bits(0, 1 << 24)
    assert distmod.window_bits_for_total("bn254_g1", 1 << 20) == L.gmsm_choose_window_bits(0, 1 << 20)
