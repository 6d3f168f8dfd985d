"""N-GPU sharded MultiExp (one process per GPU, NCCL all-gather of window partials) against the
oracle.  Launched by this test through torch.distributed.run when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["GMSM_ROOT"])
from oracle import cref
from oracle import oracle as O
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pkg = importlib.import_module("gnark_crypto_b200")
distmod = importlib.import_module("gnark-crypto_b200.dist")
g, n = "bn254_g1", 100003
G = O.GROUPS[g]
base = G.encode_affine([G.gen])[0]
pts = cref.generate_multiples(g, base, 1, n, nthreads=4)
s = cref.random_scalars(g, n, 77)
pts[11, :] = 0; s[12, :] = 0
want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=8)
lo, hi = distmod.shard_range(n, rank, world)
eng = pkg.Engine(g, hi - lo, c=13, device=lr)
sh = distmod.ShardedMultiExp(eng)
jac = sh.msm(eng.to_device(pts[lo:hi]), eng.to_device(s[lo:hi]), hi - lo).cpu().numpy().view(np.uint64)
assert np.array_equal(jac[:8], want), "rank %d: sharded result differs from oracle" % rank
# host-buffer variant through the C ABI's pipelined shard entry point (gmsm_multiexp_window_sums)
eng2 = pkg.Engine(g, hi - lo, c=0, device=lr)
sh2 = distmod.ShardedMultiExp(eng2)
jac2 = sh2.msm_from_host_pipelined(np.ascontiguousarray(pts[lo:hi]), np.ascontiguousarray(s[lo:hi])).cpu().numpy().view(np.uint64)
assert np.array_equal(jac2[:8], want), "rank %d: pipelined host-shard result differs from oracle" % rank
jac3 = sh.msm_from_host(torch.from_numpy(pts[lo:hi].view(np.int64).reshape(-1).copy()).pin_memory(),
                        torch.from_numpy(s[lo:hi].view(np.int64).reshape(-1).copy()).pin_memory(), hi - lo,
                        torch.empty((hi - lo) * 8, dtype=torch.int64, device="cuda"),
                        torch.empty((hi - lo) * 4, dtype=torch.int64, device="cuda"), chunks=3).cpu().numpy().view(np.uint64)
assert np.array_equal(jac3[:8], want), "rank %d: chunked torch-copy result differs from oracle" % rank
dist.barrier()
if rank == 0:
    print("DIST_OK world=%d" % world)
dist.destroy_process_group()
'''


def test_sharded_multiexp_nccl(tmp_path):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GMSM_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK world=%d" % world in r.stdout


def test_in_process_multi_device_one_shot(monkeypatch):
    """GMSM_DEVICES: the one-shot C-ABI call shards over several GPUs of ONE process (host thread per device,
    partials joined on the first device) -- what a single-process Go caller would use"""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    sys.path.insert(0, ROOT)
    import gnark_crypto_b200 as pkg
    from oracle import cref
    from tests.gpu_common import make_inputs

    ndev = min(torch.cuda.device_count(), 4)
    for g, n in (("bn254_g1", (1 << 17) + 5), ("bn254_g2", (1 << 16) + 3)):
        pts, s = make_inputs(g, n, 11)
        want, _, _, _ = cref.msm(g, pts, s, c=0, nthreads=8)
        A1, J1, A2, J2 = pkg.curve_package("bn254")
        Aff = A1 if g.endswith("g1") else A2
        monkeypatch.setenv("GMSM_DEVICES", ",".join(str(d) for d in range(ndev)))
        got = Aff().MultiExp(pts, s, pkg.MultiExpConfig())
        assert np.array_equal(got.limbs, want), g
        # resident bases sharded over the same devices (device = -1), full range and a sub-range that
        # straddles shard boundaries
        from importlib import import_module

        mx = import_module("gnark-crypto_b200.multiexp")
        rb = mx.ResidentBases(g, pts, device=-1)
        w = pts.shape[1]
        assert np.array_equal(rb.MultiExp(s)[:w], want), g
        lo, m = n // 5, n // 2
        want_sub, _, _, _ = cref.msm(g, pts[lo : lo + m], s[:m], c=0, nthreads=8)
        assert np.array_equal(rb.MultiExp(s[:m], offset=lo)[:w], want_sub), g
        rb.close()
        monkeypatch.delenv("GMSM_DEVICES")
        got1 = Aff().MultiExp(pts, s, pkg.MultiExpConfig())
        assert np.array_equal(got1.limbs, want), g
