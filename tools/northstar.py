"""The north-star target of BASELINE.json, for real: ONE DAPOL+ tree of 2^24 users / height 40 built over the GPUs of one box
(one dapol_sharded_build call per rank, library-owned NCCL communicator), then the inclusion proof of EVERY user
(aggregation_factor = height => one m = 64 aggregated Bulletproof under the Padding policy, 3,655 bytes) generated, streamed
to a per-rank file and verified against the gathered root -- each rank handles the leaves of its own prefix; there is no
communication after the tree build.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/northstar.py
      [users_log2=24] [height=40] [policy=0] [chunk=4096] [limit_per_rank=0] [oracle_sample_per_rank=8]
VERIFY_GROUP=G in the environment: the verifier checks G proofs per random linear combination with the bucket method
(dapol_ctx_set_verify_mode; 0 / unset = every proof on its own).
Checks: rank 0's subtree root equals the ORACLE's root of shard 0 (tests/golden/full_size_golden.json, when the config is the
golden's); every proof verifies on the GPU; a tampered proof is rejected; a sample of proofs per rank is verified by the CPU
oracle's DapolProof::verify restatement.  Timing: wall clock between barrier + synchronize pairs, max over ranks."""
import concurrent.futures as cf
import ctypes as C
import hashlib
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bench import AUDIT_SEED, PAD_SEED, synth_liabilities
from dapol_b200 import Comm, Context, CudaEngine, NativeComm, ShardedDapol, _ffi

PROVE_SEED = hashlib.sha256(b"dapol-b200-prove").digest()


def main():
    a = sys.argv[1:]
    ul = int(a[0]) if len(a) > 0 else 24
    H = int(a[1]) if len(a) > 1 else 40
    policy = int(a[2]) if len(a) > 2 else 0
    chunk = int(a[3]) if len(a) > 3 else 4096
    limit = int(a[4]) if len(a) > 4 else 0
    n_oracle = int(a[5]) if len(a) > 5 else 8
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    verify_group = int(os.environ.get("VERIFY_GROUP", "0"))
    ctx.set_verify_mode(verify_group)
    L = _ffi.lib()
    comm, engine = Comm(), CudaEngine(ctx)
    native = NativeComm(ctx, comm, "nccl")
    n = (1 << ul) // world
    p = lambda x: x.ctypes.data_as(C.c_void_p)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    liab = synth_liabilities(n, first=rank * n)
    warm = ShardedDapol.new(engine, comm, 0, liab, AUDIT_SEED, H, H, PAD_SEED, policy=policy, native=native)  # warm-up: pools, NCCL channels
    warm.close()
    barrier(); t0 = time.perf_counter()
    tree = ShardedDapol.new(engine, comm, 0, liab, AUDIT_SEED, H, H, PAD_SEED, policy=policy, native=native)
    barrier(); build_s = time.perf_counter() - t0
    root = tree.root()
    sub = tree.subtree if world > 1 else tree.subtree
    Hs = L.dapol_tree_height(sub)
    # the leaves this rank owns: real nodes of its subtree's leaf level, as whole-tree indexes
    nl = L.dapol_tree_level_size(sub, Hs)
    lidx = np.zeros(nl, np.uint64); lpad = np.zeros(nl, np.uint8)
    assert L.dapol_tree_level_copy(sub, Hs, p(lidx), None, None, None, None, p(lpad)) == 0
    mine = lidx[lpad == 0] | (np.uint64(rank) << np.uint64(Hs) if world > 1 else np.uint64(0))
    n_owned = len(mine)
    shard_root_ok = None
    if rank == 0:
        try:
            g = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "full_size_golden.json")))["c4_2p24_h40_shard0"]
            if (g["users_log2"], g["height"], 1 << g["prefix_bits"]) == (ul, H, world):
                rec = np.zeros(232, np.uint8)
                assert L.dapol_tree_root_record(sub, p(rec)) == 0
                shard_root_ok = bool(rec[128:160].tobytes().hex() == g["root"]["com"] and rec[160:192].tobytes().hex() == g["root"]["hash"] and
                                     n_owned == g["shard_leaves"])
        except (OSError, KeyError):
            pass
    if limit:
        mine = mine[:limit]
    size = L.dapol_inclusion_proof_size(H, H, policy)
    # where the proofs go: a per-rank file if the scratch disk can hold all of them, else the null device (bytes still counted)
    need = len(mine) * size
    out_dir = os.environ.get("NORTHSTAR_OUT", "/tmp")
    free = shutil.disk_usage(out_dir).free
    path = os.path.join(out_dir, f"dapol_proofs_rank{rank}.bin") if free > world * need * 1.1 + (8 << 30) else os.devnull
    sd = (C.c_uint8 * 32).from_buffer_copy(PROVE_SEED)
    rcom = np.frombuffer(root.com, np.uint8).copy(); rhash = np.frombuffer(root.hash, np.uint8).copy()
    out = np.zeros(chunk * size, np.uint8)
    off = np.arange(chunk + 1, dtype=np.uint64) * np.uint64(size)
    ok = np.zeros(chunk, np.uint8)
    Ht = max(H, 1)
    pv = np.zeros((chunk, Ht), np.uint64); pr = np.zeros((chunk, Ht, 32), np.uint8); pc = np.zeros((chunk, Ht, 32), np.uint8)
    ph = np.zeros((chunk, Ht, 32), np.uint8); lc = np.zeros((chunk, 32), np.uint8); lh = np.zeros((chunk, 32), np.uint8)
    got = C.c_uint64()
    w = mine[:min(64, len(mine))].copy()  # warm-up: builds the generator tables of this policy's shapes
    assert L.dapol_prove_batch(sub, len(w), p(w), H, policy, sd, p(out), out.nbytes, C.byref(got)) == 0, L.dapol_last_cuda_error()
    params = ctx.params()
    prove_s = verify_s = write_s = 0.0
    chunk_times, sample = [], []
    n_ok = n_done = n_bytes = 0
    tampered_rejected = None
    barrier(); t_all = time.perf_counter()
    with open(path, "wb") as f:
        for s in range(0, len(mine), chunk):
            li = np.ascontiguousarray(mine[s:s + chunk]); k = len(li)
            t1 = time.perf_counter()
            rc = L.dapol_prove_batch(sub, k, p(li), H, policy, sd, p(out), out.nbytes, C.byref(got))
            assert rc == 0, (rc, L.dapol_last_cuda_error())
            t2 = time.perf_counter()
            f.write(memoryview(out)[:k * size]); n_bytes += k * size
            t3 = time.perf_counter()
            rc = L.dapol_tree_paths(sub, k, p(li), p(pv), p(pr), p(pc), p(ph), p(lc), p(lh))  # the leaves' own proof nodes
            assert rc == 0, rc
            rc = L.dapol_verify_batch(ctx._h, 0, policy, k, p(rcom), p(rhash), p(lc), p(lh), p(out), p(off), p(ok))
            assert rc == 0, rc
            t4 = time.perf_counter()
            prove_s += t2 - t1; write_s += t3 - t2; verify_s += t4 - t3
            if len(chunk_times) < 3 or s + chunk >= len(mine):
                chunk_times.append((round(t2 - t1, 4), round(t3 - t2, 4), round(t4 - t3, 4)))
            n_ok += int(ok[:k].sum()); n_done += k
            if s == 0:
                bad = out[:size].copy(); bad[100] ^= 1  # one tampered proof must be rejected
                o1 = np.zeros(1, np.uint8)
                L.dapol_verify_batch(ctx._h, 0, policy, 1, p(rcom), p(rhash), p(lc), p(lh), p(bad), p(off), p(o1))
                tampered_rejected = not bool(o1[0])
                for j in range(min(n_oracle, k)):  # kept for the CPU oracle's verifier, outside the timed region
                    sample.append((out[j * size:(j + 1) * size].tobytes(), lc[j].tobytes(), lh[j].tobytes()))
    barrier(); all_s = time.perf_counter() - t_all
    if path != os.devnull:
        os.unlink(path)
    oracle_ok = 0
    if sample:
        from oracle import cref
        cref.build(); cref.lib()
        with cf.ThreadPoolExecutor(2) as ex:
            oracle_ok = sum(ex.map(lambda t: bool(cref.verify_inclusion(0, policy, t[0], root.com, root.hash, t[1], t[2])), sample))
    stats = torch.tensor([prove_s, verify_s, all_s, float(n_ok), float(n_done), float(n_bytes), write_s, float(oracle_ok), float(len(sample)),
                          float(n_owned)], dtype=torch.float64, device="cuda")
    mx, sm = stats.clone(), stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        done, okc = int(sm[4].item()), int(sm[3].item())
        print(json.dumps({
            "config": f"north star: 2^{ul} users, height {H}, aggregation_factor {H}, policy {'Splitting' if policy else 'Padding'}, {world} GPU(s); "
                      f"every inclusion proof generated, written out and verified",
            "n_gpus": world, "users": int(sm[9].item()), "proofs": done, "all_verified": okc == done, "verdicts_ok": okc, "tampered_rejected": tampered_rejected,
            "proof_bytes": int(size), "bytes_written": int(sm[5].item()), "sink": "per-rank file" if path != os.devnull else "null device (scratch disk too small)",
            "tree_build_s_e2e": build_s, "tree_phase_ms_rank0": tree.phase_ms, "root": root.com.hex(), "shard0_root_equals_oracle_golden": shard_root_ok,
            "oracle_verified_sample": f"{int(sm[7].item())} of {int(sm[8].item())}",
            "prove_s_max_rank": mx[0].item(), "write_s_max_rank": mx[6].item(), "verify_s_max_rank": mx[1].item(), "prove_write_verify_wall_s": mx[2].item(),
            "prove_per_s": done / mx[0].item(), "verify_per_s": done / mx[1].item(), "proofs_per_s_wall": done / mx[2].item(),
            "total_wall_s_build_plus_proofs": build_s + mx[2].item(),
            "verify_group": verify_group, "verify_fallbacks_rank0": int(ctx.verify_fallbacks), "rangeproof_window": params["rangeproof_window"], "comb_window": params["comb_window"], "chunk": chunk,
            "rank0_chunk_s_prove_write_verify": chunk_times,
            "timing": "wall clock between barrier + synchronize pairs, max over ranks; host buffers (proofs D2H after prove, written, H2D for verify)"}), flush=True)
    tree.close(); native.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
