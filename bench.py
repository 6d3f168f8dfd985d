#!/usr/bin/env python
"""bench.py -- MultiExp (MSM) throughput on B200, per the driver contract.

  python bench.py --gpus N --steps K --warmup W            # ours (CUDA engine through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU path (C port)

A "step" is one bn254 G1 MultiExp over one batch of synthetic inputs (BASELINE.json configs[1]:
n = 2^24 per GPU, window width from the engine's model).  Metric: scalar-muls/s = n_total / time.
  value : inputs already resident in HBM, CUDA events, max over ranks
  e2e   : the reference-facing call with HOST buffers (pinned), H2D of points+scalars and D2H of
          the result inside the timed region (N = 1: gmsm_multiexp one-shot drop-in)
  roofline : the accumulate kernel (bucket pass) against the measured HBM peak
  cpu_baseline : the oracle's C port of the reference algorithm on a bounded sample, all host cores
N > 1 (torchrun): points/scalars sharded contiguously, per-rank window partials, one NCCL all-gather
of W partial points, finalize on every rank ("weak": n per GPU fixed).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CURVE_BITS = {"bn254_g1": 254, "bn254_g2": 254, "bls12381_g1": 255, "bls12381_g2": 255, "bls12377_g1": 253, "bls12377_g2": 253}
AFF_BYTES = {"bn254_g1": 64, "bn254_g2": 128, "bls12381_g1": 96, "bls12381_g2": 192, "bls12377_g1": 96, "bls12377_g2": 192}
FR_MOD = {
    254: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    255: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    253: 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
}


def ncores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def synth_scalars(n, bits, seed):
    """n uniform values < r as 4 x u64 limbs (mask top limb to fr.Bits, rejection-sample), read as
    the Montgomery representation -- the analogue of fr.SetRandom (fr/element.go:302-343)."""
    rng = np.random.default_rng(seed)
    q = FR_MOD[bits]
    ql = [np.uint64((q >> (64 * i)) & (2**64 - 1)) for i in range(4)]
    out = np.empty((n, 4), dtype=np.uint64)
    todo = np.arange(n)
    top = np.uint64((1 << (bits - 192)) - 1)
    while todo.size:
        v = rng.integers(0, 2**64, size=(todo.size, 4), dtype=np.uint64)
        v[:, 3] &= top
        lt = np.zeros(todo.size, dtype=bool)
        eq = np.ones(todo.size, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (v[:, k] < ql[k])
            eq &= v[:, k] == ql[k]
        out[todo[lt]] = v[lt]
        todo = todo[~lt]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def int_pipe_fraction(g, mixed_adds_per_s, sm_mhz):
    """The binding unit of the bucket pass is the INT32 multiplier: one mixed addition is 8 M + 2 S in
    the coordinate field = 10 Fp multiplications for G1, 28 for G2 (Fp2: M = 3, S = 2 Fp multiplications), each 2N^2 + N
    IMAD.WIDE (N 32-bit limbs); an SM sub-partition issues one IMAD.WIDE warp-instruction every 4 cycles: 148 SMs x 32
    wide MADs per clock.  Returns {"wide_mads_per_s", "peak", "frac"}, or None if the clock is unknown."""
    try:
        limbs = AFF_BYTES[g] // (16 if g.endswith("g2") else 8)        # 32-bit limbs of one Fp element
        fp_muls = 28 if g.endswith("g2") else 10
        per_add = fp_muls * (2 * limbs * limbs + limbs)
        peak = 148 * 32 * float(sm_mhz) * 1e6
        ach = float(mixed_adds_per_s) * per_add
        return {"wide_mads_per_mixed_add": per_add, "wide_mads_per_s": ach, "peak": peak, "frac": ach / peak,
                "peak_source": "148 SMs x 32 IMAD.WIDE/clk at the sampled SM clock"}
    except Exception:
        return None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (C port in oracle/, the Go original cannot be
    built here) on the host cores; each step a bounded sample of the arm's workload."""
    if rank != 0:
        return
    from oracle import cref
    from oracle import oracle as O

    g = args.curve
    G = O.GROUPS[g]
    cores = ncores()
    logs = args.sample_logn if args.sample_logn else (20 if cores >= 16 else 18)
    n = 1 << logs
    base = G.encode_affine([G.scalar_mul(G.gen, 0xC0FFEE)])[0]
    pts = cref.generate_multiples(g, base, 1, n, nthreads=cores)
    s = synth_scalars(n, CURVE_BITS[g], 0x5EED0000 + 2)
    for _ in range(args.warmup):
        cref.msm(g, pts, s, c=0, nthreads=cores)
    t0 = time.perf_counter()
    used_c = 0
    for _ in range(args.steps):
        _, _, used_c, leaves = cref.msm(g, pts, s, c=0, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    line = {
        "impl": "reference", "metric": "bn254 G1 MultiExp scalar-muls/s" if g == "bn254_g1" else g + " MultiExp scalar-muls/s",
        "value": val, "unit": "scalar-muls/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": "%s MultiExp n=2^%d per GPU (configs[1]); each step a 2^%d sample" % (g, args.logn, logs),
                                        "c": used_c, "threads": cores},
        "cpu_baseline": {"value": val, "unit": "scalar-muls/s", "cores": cores, "kind": "port",
                         "sample": "n=2^%d of the same workload, C port of the reference algorithm (bestC c=%d, batch-affine buckets as in getChunkProcessorG1)" % (logs, used_c)},
        "e2e": {"value": val, "unit": "scalar-muls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curve", default="bn254_g1", choices=list(CURVE_BITS))
    ap.add_argument("--logn", type=int, default=24, help="log2 of points PER GPU")
    ap.add_argument("--total-logn", type=int, default=0,
                    help="strong scaling: log2 of the TOTAL number of points, sharded over the N ranks (e.g. 26 for BASELINE "
                         "configs[4]); default 0 = weak scaling with 2^logn points per GPU")
    ap.add_argument("--c", type=int, default=0, help="window width (0 = engine model)")
    ap.add_argument("--sample-logn", type=int, default=0, help="cpu baseline sample size (log2)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tables", action="store_true", help="skip the window-table (precomputed resident bases) measurement")
    ap.add_argument("--table-c", type=int, default=0, help="window width of the table mode (0 = engine model)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: keep NCCL's version banner off it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("gnark_crypto_b200")
    distmod = importlib.import_module("gnark-crypto_b200.dist")
    mx = importlib.import_module("gnark-crypto_b200.multiexp")
    native = importlib.import_module("gnark-crypto_b200._native")

    g = args.curve
    bits = CURVE_BITS[g]
    scaling = "weak"
    if args.total_logn:
        scaling = "strong"
        if (1 << args.total_logn) % world:
            raise SystemExit("--total-logn: 2^%d is not divisible by %d ranks" % (args.total_logn, world))
        args.logn = args.total_logn - (world.bit_length() - 1)
    n = 1 << args.logn                      # per GPU
    n_total = n * world
    lo, hi = distmod.shard_range(n_total, rank, world)
    assert hi - lo == n
    eng = pkg.Engine(g, n, c=args.c, device=local_rank)
    W, c = eng.nwin, eng.c

    # ---- synthetic inputs: on-curve bases [1+i]B generated on the device, B = [0xC0FFEE]G ----
    wds = AFF_BYTES[g] // 8
    gen_limbs = {
        "bn254_g1": None,
    }
    # generator in Go layout: take it from the engine itself: [1]G needs G; encode from constants below
    base_G = _generator_limbs(g)
    d_B = eng.generate_multiples(base_G, 0xC0FFEE, 1)
    base = d_B.cpu().numpy().view(np.uint64).copy()
    d_points = eng.generate_multiples(base, 1 + lo, n)
    h_scalars_np = synth_scalars(n, bits, 0x5EED0000 + 2 + rank)
    d_scalars = eng.to_device(h_scalars_np)
    torch.cuda.synchronize()

    sharded = distmod.ShardedMultiExp(eng) if world > 1 else None

    def step():
        if sharded is not None:
            return sharded.msm(d_points, d_scalars, n)
        return eng.msm(d_points, d_scalars, n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    eng.set_profiling(True)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    launches_per_step = eng.last_launches + (1 if world > 1 else 0)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_total / (ms_step * 1e-3)
    result_jac = out.cpu().numpy().view(np.uint64).copy()

    # stage times of the dominant kernel, averaged over a few more steps (events on the launch stream)
    acc_ms, stages = [], None
    for _ in range(3):
        step()
        torch.cuda.synchronize()
        st = eng.last_stage_ms()
        acc_ms.append(st[3])
        stages = st
    acc = float(np.mean(acc_ms))
    peak, peak_src = measured_hbm_peak()
    A = AFF_BYTES[g]
    alg_bytes = n * W * (A + 2)             # per (point, window): one affine point + one u16 digit
    achieved = alg_bytes / (acc * 1e-3) / 1e9
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("%s:%d:%d" % (g, args.logn, c))
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": "k_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": acc,
                "alg_bytes_per_launch": alg_bytes,
                "int_pipe": {"mixed_adds_per_s": n * W / (acc * 1e-3), "note": "INT32-pipe bound: ~10 modmul (~1.4k IMAD.WIDE) per 66 B"}}
    stage_names = ["digits_hist", "scan", "scatter", "accumulate", "carries", "bucket_reduce", "finalize", "total"]

    if clocks and clocks.get("sm_mhz"):
        roofline["int_pipe"]["multiplier_pipe"] = int_pipe_fraction(g, n * W / (acc * 1e-3), clocks["sm_mhz"])
    line = {
        "metric": "bn254 G1 MultiExp scalar-muls/s" if g == "bn254_g1" else g + " MultiExp scalar-muls/s",
        "value": value, "unit": "scalar-muls/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": "%s MultiExp n=2^%d per GPU%s, random scalars, on-curve bases [i]B" % (
            g, args.logn, " (BASELINE configs[1])" if (g == "bn254_g1" and args.logn == 24) else ""),
            "n_total": n_total, "c": c, "windows": W, "parallelism": "shard%d" % world,
            "l2": "inputs %.2f GiB per GPU > 126 MB L2 (no flush needed)" % ((n * (A + 32)) / 2**30)},
        "roofline": roofline,
        "stages_ms": dict(zip(stage_names, [float(x) for x in stages])),
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
    }

    # ---- window tables for resident bases (gmsm_bases_precompute): the static-SRS flow.  NOT the headline: the
    # headline `value` is the plain MultiExp whose bases may change on every call, like the reference's. ----
    if world == 1 and not args.no_tables:
        teng = pkg.Engine(g, n, c=args.table_c, device=local_rank, tables=True)
        t0 = time.perf_counter()
        d_table = teng.build_tables(d_points, n)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        teng.set_profiling(True)
        for _ in range(max(args.warmup, 3)):
            tout = teng.msm_tables(d_table, n, d_scalars, n)
        torch.cuda.synchronize()
        tev0, tev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tev0.record()
        for _ in range(args.steps):
            tout = teng.msm_tables(d_table, n, d_scalars, n)
        tev1.record()
        torch.cuda.synchronize()
        t_ms = tev0.elapsed_time(tev1) / args.steps
        if not np.array_equal(tout.cpu().numpy().view(np.uint64), result_jac):
            raise SystemExit("bench.py: window-table result differs from the plain result")
        tst = teng.last_stage_ms()
        line["resident_tables"] = {
            "value": n / (t_ms * 1e-3), "unit": "scalar-muls/s", "ms_per_step": t_ms, "c": teng.c, "windows": teng.nwin,
            "table_bytes": int(d_table.numel()) * 8, "table_build_s": build_s,
            "stages_ms": dict(zip(stage_names, [float(x) for x in tst])),
            "accumulate_alg_GBps": n * teng.nwin * (A + 2) / (tst[3] * 1e-3) / 1e9,
            "note": "bases precomputed once as W rows 2^(c*j)*P (one shared bucket set, no Horner); result bit-identical to `value`'s",
        }
        line["gpu_launches"] += teng.last_launches * args.steps
        teng.close()
        del d_table

    # ---- end to end through the reference-facing call, host buffers ----
    if not args.no_e2e:
        h_points = torch.empty(n * wds, dtype=torch.int64).pin_memory()
        h_points.copy_(d_points.cpu())
        h_scal = torch.empty(n * 4, dtype=torch.int64).pin_memory()
        h_scal.copy_(torch.from_numpy(h_scalars_np.view(np.int64).reshape(-1)))
        hp, hs = h_points.numpy().view(np.uint64), h_scal.numpy().view(np.uint64)
        jac_words = 3 * wds // 2
        if world == 1:
            L = native.lib()
            outj = np.zeros(jac_words, dtype=np.uint64)

            def e2e_step():
                rc = L.gmsm_multiexp(mx.CURVES[g], hp.ctypes.data, hs.ctypes.data, n, 0, outj.ctypes.data)
                if rc != 0:
                    raise RuntimeError(native.last_error())
                return outj
        else:
            dp2 = torch.empty_like(d_points)
            ds2 = torch.empty(n * 4, dtype=torch.int64, device=d_points.device)

            def e2e_step():
                return sharded.msm_from_host_pipelined(hp, hs).cpu().numpy().view(np.uint64)
        for _ in range(2):
            r = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if not np.array_equal(np.asarray(r), result_jac):
            raise SystemExit("bench.py: e2e result differs from the device-resident result")
        line["e2e"] = {"value": n_total / dt, "unit": "scalar-muls/s", "ms_per_step": dt * 1e3,
                       "h2d_bytes_per_step": n_total * (A + 32), "d2h_bytes_per_step": jac_words * 8 * world,
                       "path": "gmsm_multiexp one-shot (points+scalars H2D every call)" if world == 1 else
                               "per rank: gmsm_multiexp_window_sums (pinned host shard, H2D pipelined under the bucket pass) -> NCCL all-gather of W partials -> finalize -> D2H"}
        if world == 1:
            launches_e2e = native.lib().gmsm_last_oneshot_launches()
            line["gpu_launches"] += launches_e2e * args.steps
            # resident bases (prover flow: SRS static, scalars per call)
            rb = mx.ResidentBases(g, hp.reshape(n, wds), device=local_rank)
            rb.MultiExp(hs.reshape(n, 4))
            t0 = time.perf_counter()
            for _ in range(args.steps):
                rr = rb.MultiExp(hs.reshape(n, 4))
            dtr = (time.perf_counter() - t0) / args.steps
            if not np.array_equal(rr, result_jac):
                raise SystemExit("bench.py: resident-bases result differs")
            line["e2e_resident_bases"] = {"value": n / dtr, "unit": "scalar-muls/s", "ms_per_step": dtr * 1e3,
                                          "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": jac_words * 8}
            if not args.no_tables:
                tc = rb.Precompute(args.table_c)
                rb.MultiExp(hs.reshape(n, 4))
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    rr = rb.MultiExp(hs.reshape(n, 4))
                dtt = (time.perf_counter() - t0) / args.steps
                if not np.array_equal(rr, result_jac):
                    raise SystemExit("bench.py: resident-bases (window tables) result differs")
                line["e2e_resident_tables"] = {"value": n / dtt, "unit": "scalar-muls/s", "ms_per_step": dtt * 1e3, "c": tc,
                                               "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": jac_words * 8}
            rb.close()
        del h_points, h_scal

    # ---- CPU baseline: the oracle's C port on a bounded sample (rank 0, N = 1 only) + parity on it ----
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cref

        cores = ncores()
        logs = args.sample_logn if args.sample_logn else min(args.logn, 20 if cores >= 16 else 18)
        ns = 1 << logs
        sp = d_points[: ns * wds].cpu().numpy().view(np.uint64).reshape(ns, wds)
        ss = h_scalars_np[:ns]
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            want, _, used_c, leaves = cref.msm(g, sp, ss, c=0, nthreads=cores)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        eng_s = pkg.Engine(g, ns, c=0, device=local_rank)
        got = eng_s.msm_host_result(eng_s.to_device(sp), eng_s.to_device(ss), ns)
        eng_s.close()
        if not np.array_equal(got[:wds], want):
            raise SystemExit("bench.py: GPU result differs from the CPU oracle on the baseline sample")
        line["cpu_baseline"] = {"value": ns / best, "unit": "scalar-muls/s", "cores": cores, "kind": "port",
                                "sample": "first 2^%d points/scalars of the same workload; C port of the reference algorithm "
                                          "(bestC c=%d, %d sub-MSMs, batch-affine buckets as in getChunkProcessorG1), best of 2, GPU result on the sample "
                                          "bit-exact" % (logs, used_c, leaves)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def _generator_limbs(g):
    """group generators in Go memory layout (Montgomery limbs), precomputed from
    ecc/bn254/bn254.go:111-119 and ecc/bls12-381/bls12-381.go:107-116 by tools/gen_generators.py"""
    tab = json.load(open(os.path.join(ROOT, "gnark-crypto_b200", "generators.json")))
    return np.array([int(x, 16) for x in tab[g]], dtype=np.uint64)


if __name__ == "__main__":
    main()
