#!/usr/bin/env python3
"""bench.py -- DAPOL+ tree build throughput (leaves/s) on B200, BASELINE.json config[1]:
2^20 users, height-32 tree (leaf derivation + commit + hash + merge + padding) on 1 GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                         CPU arm: the oracle port on all host threads

One JSON line on stdout (rank 0).  `value` = whole-job leaves/s with inputs resident in HBM,
`e2e` = the same through Dapol.new() on pinned host buffers (H2D of ids/values + D2H of the root
inside the timed region).  `roofline` is for the dominant kernel (padding-node pass, k_pad) against
the integer-multiply pipe peak measured live (IMAD.WIDE.U32 microbenchmark); `cpu_baseline` is the
C oracle (a port: the Rust reference cannot be built here) on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AUDIT_SEED = b"dapol-b200-bench"
PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
# algorithmic work per unit, MAC32 = one 32x32->64 multiply-accumulate (SURVEY.md 8(d) / BASELINE.md / DESIGN.md)
MAC32_LEAF, MAC32_PAD, MAC32_MERGE = 53.6e3, 45.5e3, 13.9e3


def executed_mac32(comb_window, node_batch):
    """Field-arithmetic MAC32 the kernels here execute per unit: fixed-base comb on half points
    (253/W + 1 windows for a 253-bit scalar, 64/WV + 1 for a 64-bit value; the first window initialises the accumulator
    with one product, every other one is a mixed addition of 7 FM), batched double-and-compress (prepare 4 FS + 5 FM, 3 FM of
    Montgomery's trick, finish 11 FM, one inversion of 254 FS + 11 FM per node_batch nodes).
    roofline.frac uses these; survey_unit_frac uses SURVEY 8(d)'s figures for the reference's algorithm."""
    FM, FS, SCMUL = 72, 44, 128
    madd, full_add = 7 * FM, 9 * FM
    compress = (4 * FS + 5 * FM) + 3 * FM + 11 * FM + (254 * FS + 11 * FM) / node_batch
    value_window = comb_window if comb_window <= 16 else 22      # comb_value_window in ge25519.cuh
    nwr, nwv = 253 // comb_window + 1, 64 // value_window + 1
    pad = (nwr - 1) * madd + FM + 2 * SCMUL + compress          # ChaCha draw -> wide reduce with the halving folded in; comb; compress
    leaf = (nwr + nwv - 1) * madd + FM + SCMUL + compress       # halve r; two combs; compress
    merge = full_add + 2 * SCMUL + compress                     # point add; r_L + r_R mod l; compress
    return leaf, pad, merge
NODE_BYTES = 104  # com 32 + hash 32 + v 8 + r 32
# ncu --set full capture of the k_pad launch of THIS kernel version on THIS workload (roofline.traffic); re-captured whenever k_pad changes
PAD_NCU_CAPTURE = "r02a_k_pad_ncu_full.json"


def splitmix64(n, seed=0xDA901):
    """Deterministic value stream: values = splitmix64_i & 0xffffffff."""
    x = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed))
    z = x.copy()
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synth_liabilities(n, first=0):
    """internal_id = le64(i), external_id = le64(i ^ 0x9E3779B97F4A7C15), value = u32."""
    i = np.arange(first, first + n, dtype=np.uint64)
    iid = i.view(np.uint8).copy()
    eid = (i ^ np.uint64(0x9E3779B97F4A7C15)).view(np.uint8).copy()
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(8)
    vals = splitmix64(first + n)[first:] & np.uint64(0xFFFFFFFF)
    return iid, off, eid, off.copy(), vals


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        rows = [r for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if rows:
            sm = [float(r[0]) for r in rows]
            busy = [x for x in sm if x > 0.5 * max(sm)] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [nm for k, nm in enumerate(names) if any(r[2 + k].strip().lower() == "active" for r in rows)]
            out["samples"] = len(rows)
        return out


def run_reference(args):
    """CPU arm: the oracle port (oracle/c, OpenMP) on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    cref.build()
    cref.lib()
    cores = os.cpu_count() or 1
    sample_log2 = min(args.users_log2, args.cpu_sample_log2)
    H = args.height - (args.users_log2 - sample_log2)  # same sparsity 2^H / N as the full workload
    n = 1 << sample_log2
    iid, io, eid, eo, vals = synth_liabilities(n)
    times = []
    for step in range(args.warmup_ref + args.steps):
        t0 = time.perf_counter()
        rc, idx, bl, _ = cref.derive_leaves(0, iid, io, eid, eo, AUDIT_SEED, H)
        assert rc == 0
        order = np.argsort(idx, kind="stable")
        t = cref.Tree(0, H, idx[order], vals[order], bl[order], PAD_SEED, 0, cores)
        root = t.root()
        dt = time.perf_counter() - t0
        del t
        if step >= args.warmup_ref:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    sample = f"2^{sample_log2} users at height {H} (same 2^{args.height - args.users_log2} sparsity as the full workload), {len(times)} builds"
    line = {
        "impl": "reference", "metric": "leaves/sec tree build", "value": value, "unit": "leaves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup_ref, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs / u64 (integer)", "data": "synthetic",
        "config": {"workload": f"DAPOL+ tree build, 2^{args.users_log2} users, height {args.height}, D=blake3 (timed on a bounded sample)",
                   "note": "Rust reference unbuildable here (no cargo, un-vendored crates); this is the C oracle port, OpenMP"},
        "cpu_baseline": {"value": value, "unit": "leaves/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "leaves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "root": root["comc"].hex()[:16],
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(args):
    from oracle import cref
    cref.build()
    cref.lib()
    cores = os.cpu_count() or 1
    sample_log2 = min(args.users_log2, args.cpu_sample_log2)
    H = args.height - (args.users_log2 - sample_log2)
    n = 1 << sample_log2
    iid, io, eid, eo, vals = synth_liabilities(n)
    reps, dt = 2, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        rc, idx, bl, _ = cref.derive_leaves(0, iid, io, eid, eo, AUDIT_SEED, H)
        order = np.argsort(idx, kind="stable")
        t = cref.Tree(0, H, idx[order], vals[order], bl[order], PAD_SEED, 0, cores)
        dt += time.perf_counter() - t0
        root = t.root()["comc"]
        del t
    return {"value": reps * n / dt, "unit": "leaves/s", "cores": cores, "kind": "port",
            "sample": f"2^{sample_log2} users at height {H} (same sparsity as the workload), {reps} builds, {dt:.1f} s, OpenMP over {cores} threads"}, (n, H, root)


def rangeproof_leg(ctx, L, dev, world, dist, args, imad_peak):
    """Range proofs/s (BASELINE.json's second metric): K independent 64-bit Bulletproofs per GPU, prove then verify,
    device-resident (CUDA events) and end to end through the host-buffer C ABI.  Shapes: m = 1 (single proofs, C5) and
    m = 32 (the aggregated proof of one height-32 inclusion proof under the Padding policy, C3)."""
    import ctypes as C
    import torch
    out = {}
    seed = (C.c_uint8 * 32).from_buffer_copy(PAD_SEED)
    for m, K in ((1, args.rp_singles), (32, args.rp_aggregates)):
        if K <= 0:
            continue
        rng = np.random.default_rng(1234 + m)
        vals = rng.integers(0, 1 << 63, size=(K, m), dtype=np.uint64)
        bl = rng.integers(0, 256, size=(K, m, 32), dtype=np.uint8); bl[:, :, 31] &= 0x0F
        streams = np.arange(K, dtype=np.uint64); bases = np.zeros(K, np.uint64)
        size = L.dapol_rangeproof_size(64, m)
        coms = np.stack([ctx.commit_batch(vals[:, j], bl[:, j]) for j in range(m)], axis=1)
        tv, tb, ts, tbs, tc = (torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev) for a in (vals, bl, streams, bases, coms))
        d_proofs = torch.empty(K * size, dtype=torch.uint8, device=dev)
        d_ok = torch.empty(K, dtype=torch.uint8, device=dev)

        def prove_dev():
            rc = L.dapol_rangeproof_prove_batch_dev(ctx._h, 64, m, K, tv.data_ptr(), tb.data_ptr(), seed, ts.data_ptr(), tbs.data_ptr(), d_proofs.data_ptr())
            assert rc == 0, rc

        def verify_dev():
            rc = L.dapol_rangeproof_verify_batch_dev(ctx._h, 64, m, K, d_proofs.data_ptr(), size, tc.data_ptr(), d_ok.data_ptr())
            assert rc == 0, rc
        prove_dev(); verify_dev()  # builds the generator tables, warms up
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev[0].record(); prove_dev(); ev[1].record()
        kt_prove = ctx.rangeproof_last_kernel_times()   # prove_dev returned after its own stream sync
        verify_dev(); ev[2].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], device=dev)
        all_ok = bool(d_ok.all().item())
        # head-to-head of the verifier's variable-base MSM: per-proof Straus (above) vs groups of G proofs checked by one random
        # linear combination with the bucket method (Pippenger; dapol_ctx_set_verify_mode).  Same proofs, all valid.
        batched = []
        for G in [g for g in (16, 256, 4096, K) if g <= K]:
            ctx.set_verify_mode(G)
            f0 = ctx.verify_fallbacks
            verify_dev()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); verify_dev(); e1.record()
            torch.cuda.synchronize()
            batched.append({"group": G, "verify_per_s": world * K / e0.elapsed_time(e1) * 1e3, "all_verified": bool(d_ok.all().item()),
                            "fallbacks": int(ctx.verify_fallbacks - f0), "kernel_class_ms": ctx.rangeproof_last_kernel_times()})
        ctx.set_verify_mode(0)
        # end to end: host buffers in, proofs / verdicts out
        h_proofs = np.zeros((K, size), np.uint8); h_ok = np.zeros(K, np.uint8)
        t0 = time.perf_counter()
        rc = L.dapol_rangeproof_prove_batch(ctx._h, 64, m, K, vals.ctypes.data, bl.ctypes.data, seed, streams.ctypes.data, bases.ctypes.data,
                                            h_proofs.ctypes.data)
        t1 = time.perf_counter()
        rc |= L.dapol_rangeproof_verify_batch(ctx._h, 64, m, K, h_proofs.ctypes.data, size, coms.ctypes.data, h_ok.ctypes.data)
        t2 = time.perf_counter()
        assert rc == 0
        te = torch.tensor([(t1 - t0) * 1e3, (t2 - t1) * 1e3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        pm, vm = t.tolist(); pe, ve = te.tolist()
        # roofline of the dominant kernel k_rp_p10 (L / R of the inner-product rounds over the generator tables): executed
        # field + scalar MAC32 / its device time (CUDA events on the launching stream, this rank)
        N, lg = 64 * m, (64 * m).bit_length() - 1
        rpw = ctx.params()["rangeproof_window"]
        w = 253 // rpw + 1
        tab_rounds = lg if N < 1024 else lg - 5       # rp_switch_round: large aggregates fold the generators from round lg - 4 on
        mac_p10 = K * tab_rounds * (2 * N * w * 7 * 72 + 3 * N * 128)
        kt = kt_prove
        achieved = mac_p10 / (kt["p10"] * 1e-3) / 1e9 if kt["p10"] > 0 else None
        out[f"n64_m{m}"] = {"proofs_per_gpu": K, "prove_per_s": world * K / pm * 1e3, "verify_per_s": world * K / vm * 1e3,
                            "prove_plus_verify_per_s": world * K / (pm + vm) * 1e3,
                            "e2e_prove_per_s": world * K / pe * 1e3, "e2e_verify_per_s": world * K / ve * 1e3,
                            "e2e_prove_plus_verify_per_s": world * K / (pe + ve) * 1e3,
                            "all_verified": all_ok and bool(h_ok.all()), "proof_bytes": int(size),
                            "verify_batched_bucket_method": batched,
                            "roofline": {"bound": "imad", "kernel": "k_rp_p10 (L/R MSMs of the inner-product rounds over the generator tables)",
                                         "achieved": achieved, "peak": imad_peak, "unit": "GMAC32/s",
                                         "frac": achieved / imad_peak if achieved and imad_peak else None,
                                         "launch_ms_total": kt["p10"], "launches": tab_rounds, "rangeproof_window": rpw,
                                         "mac32_per_proof_p10": mac_p10 / K, "share_of_prove": kt["p10"] / kt["total"] if kt["total"] else None,
                                         "kernel_class_ms": kt,
                                         "whole_prover_mac32_per_proof": prover_mac32(m, rpw),
                                         "whole_prover_frac": (prover_mac32(m, rpw) * K / (pm * 1e-3 / 1) / 1e9 / imad_peak) if imad_peak else None}}
    return out


def prover_mac32(m, rpw):
    """Executed field + scalar MAC32 of one 64-bit, m-party proof: A and S (3N table multiplications worth of
    mixed additions: A adds one table entry per bit = N/w multiplications' worth), the table rounds (2N multiplications each), the
    hybrid late rounds of large aggregates (materialise: 2 * 32 * N/32 = 2N; then ~184 variable-base multiplications at ~19
    table ones each), T1/T2 and the vector passes (~40 N scalar products)."""
    N, lg = 64 * m, (64 * m).bit_length() - 1
    w = 253 // rpw + 1
    tmul = w * 7 * 72
    tab_rounds = lg if N < 1024 else lg - 5
    hybrid = (2 * N + 184 * 19) * tmul if N >= 1024 else 0
    return (2 * N + N / w) * tmul + tab_rounds * (2 * N * tmul + 3 * N * 128) + hybrid + 40 * N * 128


def rangeproof_cpu_baseline(budget_s=6.0):
    """CPU port (oracle/c) of the prover and the verifier of BASELINE's second metric on all host threads, one proof per
    thread at a time (the reference is single-threaded per proof): m = 1 and m = 32, a bounded sample each."""
    import concurrent.futures as cf
    from oracle import cref
    cref.build(); cref.lib()
    cores = os.cpu_count() or 1
    out = {}
    for m, per_thread in ((1, 192), (32, 6)):  # ~4 s of proving per shape on every core
        k = cores * per_thread
        rng = np.random.default_rng(99 + m)
        vals = rng.integers(0, 1 << 63, size=(k, m), dtype=np.uint64)
        bl = rng.integers(0, 256, size=(k, m, 32), dtype=np.uint8); bl[:, :, 31] &= 0x0F
        coms = [[cref.commit(int(vals[i, j]), bl[i, j].tobytes()) for j in range(m)] for i in range(k)]

        def prove(i):
            return cref.rp_prove([int(x) for x in vals[i]], [b.tobytes() for b in bl[i]], PAD_SEED, i, 0, 64)
        with cf.ThreadPoolExecutor(cores) as ex:
            t0 = time.perf_counter(); proofs = list(ex.map(prove, range(k))); tp = time.perf_counter() - t0
            t0 = time.perf_counter(); oks = list(ex.map(lambda i: cref.rp_verify(proofs[i], coms[i], 64), range(k))); tv = time.perf_counter() - t0
        assert all(oks)
        out[f"n64_m{m}"] = {"prove_per_s": k / tp, "verify_per_s": k / tv, "prove_plus_verify_per_s": k / (tp + tv), "unit": "proofs/s",
                            "cores": cores, "kind": "port",
                            "sample": f"{k} proofs ({per_thread} per thread), prove {tp:.1f} s, verify {tv:.1f} s; generators precomputed once (the reference "
                                      f"re-derives BulletproofGens::new(64, m) per call, src/range/mod.rs:50,66,85,104)"}
    return out


def c1_leg(ctx, L):
    """BASELINE config 1 (the only configuration the reference itself benchmarks, benches/dapol.rs:24-57,59-141,149-175):
    1024 leaves at stride 2^16 / 1024 of a height-16 tree (new_blank + build from ready nodes), aggregation_factor = height,
    ONE inclusion proof generated and verified, both policies.  Latencies are wall clock through the host-buffer C ABI
    (median of repeats); the CPU port runs the same calls on one thread (the reference is single-threaded)."""
    from dapol_b200 import Dapol, DapolProof, DapolProofNode
    from oracle import cref
    cref.build(); cref.lib()
    n, H = 1024, 16
    idx = (np.arange(n, dtype=np.uint64) * np.uint64((1 << H) // n))
    vals = splitmix64(n) & np.uint64(0xFFFFFFFF)
    bl = np.frombuffer(b"".join(cref.rng_scalar(PAD_SEED, i, 7) for i in range(n)), np.uint8).reshape(n, 32).copy()

    def med(f, reps):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
        return statistics.median(ts) * 1e3, r
    out = {"workload": "1024 leaves, height 16, stride layout, aggregation_factor 16 (benches/dapol.rs:149-175)"}
    tree = Dapol.new_blank(ctx, 0, H, H)
    launches0 = ctx.kernel_launches
    tree.build(idx, vals, bl, PAD_SEED)
    out["build_launches"] = int(ctx.kernel_launches - launches0)
    ms, _ = med(lambda: tree.build(idx, vals, bl, PAD_SEED), 20)
    out["gpu_build_ms"] = ms
    t0 = time.perf_counter(); ora = cref.Tree(0, H, idx, vals, bl, PAD_SEED, 0, 1); out["cpu_build_ms_1thread"] = (time.perf_counter() - t0) * 1e3
    ro, rg = ora.root(), tree.root_raw()
    out["gpu_root_matches"] = bool((rg.value, rg.com, rg.hash) == (ro["v"], ro["comc"], ro["hash"]))
    leaf = int(idx[n // 2])
    paths = tree.paths([leaf])
    ln = DapolProofNode(paths["leaf_comc"][0].tobytes(), paths["leaf_hash"][0].tobytes())
    for pol, name in ((0, "padding"), (1, "splitting")):
        tree.policy = pol
        tree.generate_proof(leaf, PAD_SEED)  # generator tables + warm-up
        launches0 = ctx.kernel_launches
        ms, pr = med(lambda: tree.generate_proof(leaf, PAD_SEED), 10)
        out[f"gpu_prove_ms_{name}"] = ms
        out[f"prove_launches_{name}"] = int((ctx.kernel_launches - launches0) // 10)
        ms, ok = med(lambda: pr.verify(ctx, tree.root(), ln), 10)
        out[f"gpu_verify_ms_{name}"] = ms
        t0 = time.perf_counter(); want = ora.prove_inclusion(leaf, H, pol, PAD_SEED); out[f"cpu_prove_ms_{name}"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); okc = cref.verify_inclusion(0, pol, want, ro["comc"], ro["hash"], ln.com, ln.hash)
        out[f"cpu_verify_ms_{name}"] = (time.perf_counter() - t0) * 1e3
        out[f"proof_bytes_match_{name}"] = bool(pr.serialize() == want) and bool(ok) and bool(okc)
    tree.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--users-log2", type=int, default=20, help="users per GPU (weak scaling)")
    ap.add_argument("--height", type=int, default=32, help="tree height at 1 GPU; N GPUs build ONE tree of height + log2(N)")
    ap.add_argument("--comb-window", type=int, default=26, help="window of the tree's fixed-base tables: 26 bits = 32 GB of HBM tables, 10 windows per "
                    "blinding (measured 25.0 vs 26.0 ms per build against the library default 24 = 8.9 GB, profiles/r02_variants.txt 11); 0 = library default")
    ap.add_argument("--cpu-sample-log2", type=int, default=17)
    ap.add_argument("--warmup-ref", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c1", action="store_true", help="skip the BASELINE config-1 latency leg (1024 leaves / height 16)")
    ap.add_argument("--staged-sharding", action="store_true", help="N > 1: the staged C-ABI calls orchestrated from Python (round-1 path) "
                    "instead of the one-call dapol_sharded_build")
    ap.add_argument("--rp-singles", type=int, default=16384, help="single range proofs per GPU in the range-proof leg (0 = skip)")
    ap.add_argument("--rp-aggregates", type=int, default=2048, help="m = 32 aggregated range proofs per GPU (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dapol_b200 import Comm, Context, CudaEngine, Dapol, NativeComm, ShardedDapol, _ffi
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    ctx = Context(local, args.comb_window)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    L = _ffi.lib()
    comm = Comm()
    engine = CudaEngine(ctx)
    k = (world - 1).bit_length()
    assert (1 << k) == world, "--gpus must be a power of two"
    # N > 1: the library's own NCCL communicator; the whole sharded build is ONE C-ABI call per rank (dapol_sharded_build)
    native = NativeComm(ctx, comm, "nccl") if world > 1 and not args.staged_sharding else None
    exch_ms = []

    # Weak scaling: every rank holds 2^users_log2 users.  N GPUs build ONE tree over all N * 2^users_log2 users, of height
    # height + log2 N (same sparsity, so per-GPU work is fixed): per-user hashing on the owner of the slice, all-to-all of
    # 96 B per local user (duplicate check + index claims; collisions resolved by the owner of the index prefix), independent
    # subtrees per leaf-index prefix, one all-gather of the N subtree roots.
    n, H = 1 << args.users_log2, args.height + k
    iid, io, eid, eo, vals = synth_liabilities(n, first=rank * n)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_iid, h_io, h_eid, h_eo, h_vals = map(pin, (iid, io.view(np.int64), eid, eo.view(np.int64), vals.view(np.int64)))
    d_iid, d_io, d_eid, d_eo, d_vals = (t.to(dev) for t in (h_iid, h_io, h_eid, h_eo, h_vals))
    seed = (C.c_uint8 * 32).from_buffer_copy(PAD_SEED)
    aseed = (C.c_uint8 * len(AUDIT_SEED)).from_buffer_copy(AUDIT_SEED)
    phase_keys = ("structure", "leaves", "padding", "merges", "total")

    def step_dev():
        """one build with the liabilities already resident in HBM; returns (nodes, pads, phase times of the rank's tree build)"""
        if world == 1:
            h = C.c_void_p(); err = C.c_uint64()
            rc = L.dapol_tree_build_from_liabilities_dev(ctx._h, 0, H, n, d_iid.data_ptr(), d_io.data_ptr(), d_eid.data_ptr(), d_eo.data_ptr(),
                                                         d_vals.data_ptr(), aseed, len(AUDIT_SEED), seed, 0, C.byref(h), C.byref(err))
            assert rc == 0, (rc, L.dapol_last_cuda_error())
            res = (L.dapol_tree_num_nodes(h), L.dapol_tree_num_padding(h), ctx.last_build_times())
            L.dapol_tree_destroy(h)
            return res
        t = ShardedDapol.new(engine, comm, 0, (d_iid, d_io, d_eid, d_eo, d_vals), AUDIT_SEED, H, H, PAD_SEED, native=native)
        if t.phase_ms:
            exch_ms.append(t.phase_ms)
        res = (L.dapol_tree_num_nodes(t.subtree) if t.subtree else 0, L.dapol_tree_num_padding(t.subtree) if t.subtree else 0,
               engine.last_shard_times or dict.fromkeys(phase_keys, 0.0))
        t.close()
        return res

    def step_host():
        """the same through the public call on pinned HOST buffers: H2D of the liabilities + D2H of the root inside"""
        if world == 1:
            h = C.c_void_p(); err = C.c_uint64()
            rc = L.dapol_tree_build_from_liabilities(ctx._h, 0, H, n, h_iid.data_ptr(), h_io.data_ptr(), h_eid.data_ptr(), h_eo.data_ptr(),
                                                     h_vals.data_ptr(), aseed, len(AUDIT_SEED), seed, 0, C.byref(h), C.byref(err))
            assert rc == 0, (rc, L.dapol_last_cuda_error())
            com = np.zeros(32, np.uint8); hs = np.zeros(32, np.uint8); bl = np.zeros(32, np.uint8); v = C.c_uint64()
            L.dapol_tree_root(h, com.ctypes.data, hs.ctypes.data, C.byref(v), bl.ctypes.data)  # D2H of the step's result
            L.dapol_tree_destroy(h)
            return com.tobytes(), v.value
        t = ShardedDapol.new(engine, comm, 0, (h_iid, h_io, h_eid, h_eo, h_vals), AUDIT_SEED, H, H, PAD_SEED, native=native)
        r = t.root_raw()
        t.close()
        return r.com, r.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    imad_peak = ctx.imad_peak(1)  # GMAC32/s: IMAD.WIDE.U32 with a 64-bit addend, measured live
    fe_rate = (ctx.fe_bench(0), ctx.fe_bench(1))

    # ---- device-resident arm
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step_dev()
    phase_ms = {kk: [] for kk in phase_keys}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = ctx.kernel_launches
    e0.record()
    for _ in range(args.steps):
        nodes, pads, times = step_dev()
        for kk in phase_keys:
            phase_ms[kk].append(times[kk])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ctx.kernel_launches - launches1
    t_ms = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * n * args.steps / (ms_max * 1e-3)

    # ---- end-to-end arm: public call on pinned host buffers, H2D + D2H inside the timed region
    for _ in range(2):
        step_host()
    barrier()
    e0.record()
    for _ in range(args.steps):
        root_com, root_v = step_host()
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / (float(t_ms.item()) * 1e-3)
    h2d = int(iid.nbytes + io.nbytes + eid.nbytes + eo.nbytes + vals.nbytes)
    d2h = 104 + 65 * 8 + 32 * 4  # root record + level histogram + collision-round counters (approx. 4 rounds)

    # The tree context (up to 33 GB of comb tables) is released before the range-proof leg, which runs on a context of its own with the
    # L2-resident comb tables, so that the generator tables get their full HBM budget (110 GB at m = 32).
    params = ctx.params()
    spot = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cb, (sn, sH, sroot) = cpu_baseline_leg(args)
        # the same sample through the GPU path must give the oracle's root (parity spot check, untimed)
        t = Dapol.new(ctx, 0, synth_liabilities(sn), AUDIT_SEED, sH, sH, PAD_SEED)
        cb["gpu_root_matches"] = bool(t.root_raw().com == sroot)
        t.close()
        spot = cb
    used_native = native is not None
    if native is not None:
        native.close()
        native = None
    ctx.close()
    ctx = Context(local, 15)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    rp = rangeproof_leg(ctx, L, dev, world, dist, args, imad_peak) if (args.rp_singles or args.rp_aggregates) else None

    if rank == 0:
        internal = (nodes - 1) // 2  # every internal node has exactly two children
        leaves_here = nodes - pads - internal
        med = {kk: statistics.median(v) for kk, v in phase_ms.items()}
        MAC32_LEAF_EXEC, MAC32_PAD_EXEC, MAC32_MERGE_EXEC = executed_mac32(params["comb_window"], params["node_batch"])
        achieved = pads * MAC32_PAD_EXEC / (med["padding"] * 1e-3) / 1e9  # GMAC32/s
        build_exec = leaves_here * MAC32_LEAF_EXEC + pads * MAC32_PAD_EXEC + internal * MAC32_MERGE_EXEC
        build_survey = leaves_here * MAC32_LEAF + pads * MAC32_PAD + internal * MAC32_MERGE
        traffic = None  # DRAM bytes of one k_pad launch from the committed ncu --set full capture, if it is of this workload
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", PAD_NCU_CAPTURE)))["launches"][0]
            if int(cap["units_in_launch"]) == int(pads) and params["comb_window"] == 24:
                traffic = cap["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "leaves/sec tree build", "value": value, "unit": "leaves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 limbs (8x32-bit GF(2^255-19), integer)", "data": "synthetic",
            "config": {"workload": f"DAPOL+ tree build from liabilities, 2^{args.users_log2} users/GPU, one tree of height {H} over "
                                   f"{world * n} users, D=blake3 (leaf derivation + commit + hash + merge + padding)",
                       "users_per_gpu": n, "height": H, "nodes_rank0": nodes, "padding_nodes_rank0": pads,
                       "comb_window": params["comb_window"], "nodes_per_inversion": params["node_batch"],
                       "parallelism": ((f"one tree sharded by {k}-bit leaf-index prefix over {world} GPUs, one dapol_sharded_build call per rank "
                                        f"(library-owned NCCL communicator): all-to-all of 96 B per local user (duplicate check + index claims), "
                                        f"losers-only re-claim rounds, all-gather of padding counts and of {world} subtree roots (232 B)")
                                       if used_native else (f"one tree sharded by {k}-bit leaf-index prefix over {world} GPUs (staged C-ABI calls, "
                                                       f"torch.distributed): all-gather of user records (112 B/user) + all-gather of {world} subtree roots"))
                       if world > 1 else "single GPU",
                       "l2": "per-step working set (node store + half points ~%.1f GB) >> 126 MB L2; no reuse across steps" % (nodes * 232 / 1e9)},
            "phase_ms": dict(med, **({"sharded_" + kk: statistics.median(x[kk] for x in exch_ms[-args.steps:]) for kk in exch_ms[-1]} if exch_ms else {})),
            "roofline": {"bound": "imad", "kernel": "k_pad (padding-node pass)", "achieved": achieved, "peak": imad_peak,
                         "unit": "GMAC32/s", "frac": achieved / imad_peak,
                         "peak_source": "measured live: IMAD.WIDE.U32 (64-bit addend) microbenchmark, dapol_imad_peak variant 1; "
                                        "MEASURED_PEAKS.json has no integer peak",
                         "algorithmic_mac32_per_launch": pads * MAC32_PAD_EXEC, "mac32_per_pad_executed": MAC32_PAD_EXEC,
                         "launch_ms": med["padding"],
                         "survey_unit_frac": pads * MAC32_PAD / (med["padding"] * 1e-3) / 1e9 / imad_peak,
                         "traffic": traffic,
                         "whole_build_frac": build_exec / (med["total"] * 1e-3) / 1e9 / imad_peak,
                         "whole_build_survey_unit_frac": build_survey / (med["total"] * 1e-3) / 1e9 / imad_peak,
                         "fe_mul_Gop_s": fe_rate[0], "fe_sq_Gop_s": fe_rate[1],
                         "fe_mul_frac_of_peak": fe_rate[0] * 72 / imad_peak,
                         "hbm_GBs_algorithmic": (NODE_BYTES * nodes + 2 * NODE_BYTES * internal) / (med["total"] * 1e-3) / 1e9},
            "e2e": {"value": e2e_value, "unit": "leaves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "root": root_com.hex()[:16],
        }
        if rp is not None:
            line["range_proofs"] = rp
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = spot
            # the timed full-size build against the oracle's root of the same workload (tests/golden/full_size_golden.json,
            # produced once on the CPU by tests/golden/gen_golden_full.py: the oracle needs minutes for this tree)
            try:
                gold = json.load(open(os.path.join(ROOT, "tests", "golden", "full_size_golden.json")))["c2_2p20_h32"]
                if (gold["users_log2"], gold["height"]) == (args.users_log2, H):
                    line["cpu_baseline"]["full_size_root_matches_oracle_golden"] = bool(root_com.hex() == gold["root"]["com"])
            except (OSError, KeyError):
                pass
            if rp is not None:
                rp["cpu_baseline"] = rangeproof_cpu_baseline()
                for key, cb_rp in rp["cpu_baseline"].items():
                    if key in rp:
                        rp[key]["e2e_speedup_vs_cpu_port"] = rp[key]["e2e_prove_plus_verify_per_s"] / cb_rp["prove_plus_verify_per_s"]
            if not args.no_c1:
                line["c1"] = c1_leg(ctx, L)
        elif world > 1:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if native is not None:
        native.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
