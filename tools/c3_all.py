"""BASELINE.json configs[2] at full size: inclusion proofs (aggregated Bulletproofs, aggregation_factor = height) for ALL users of
one tree, generated AND verified, sharded over the GPUs of one box (one process per GPU, torchrun).  Every rank proves and
verifies the leaves of its own prefix through the C ABI (host buffers in / out); no communication after the tree build.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c3_all.py [users_log2=20] [height=32]
      [policy=0] [chunk=8192] [limit_per_rank=0]
Timing: device work bracketed by barrier + synchronize, max over ranks (the slowest rank's wall time between the barriers)."""
import ctypes as C
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bench import AUDIT_SEED, PAD_SEED, synth_liabilities
from dapol_b200 import Comm, Context, CudaEngine, ShardedDapol, _ffi

PROVE_SEED = hashlib.sha256(b"dapol-b200-prove").digest()


def main():
    ul = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    policy = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
    limit = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    L = _ffi.lib()
    comm, engine = Comm(), CudaEngine(ctx)
    n_total = 1 << ul
    n = n_total // world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier(); t0 = time.perf_counter()
    tree = ShardedDapol.new(engine, comm, 0, synth_liabilities(n, first=rank * n), AUDIT_SEED, H, H, PAD_SEED, policy=policy)
    barrier(); build_s = time.perf_counter() - t0
    root = tree.root()
    # leaves this rank owns (whole-tree indexes with its prefix), in input order
    lmap = tree.leaf_index_map.cpu().numpy().view(np.uint64)
    mine = lmap[(lmap >> np.uint64(tree.sub_height)) == np.uint64(rank)] if world > 1 else lmap
    if limit:
        mine = mine[:limit]
    size = L.dapol_inclusion_proof_size(H, H, policy)
    sd = (C.c_uint8 * 32).from_buffer_copy(PROVE_SEED)
    rcom = np.frombuffer(root.com, np.uint8).copy(); rhash = np.frombuffer(root.hash, np.uint8).copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    out = np.zeros(chunk * size, np.uint8)
    off = np.arange(chunk + 1, dtype=np.uint64) * np.uint64(size)
    ok = np.zeros(chunk, np.uint8)
    Ht = max(H, 1)
    pv = np.zeros((chunk, Ht), np.uint64); pr = np.zeros((chunk, Ht, 32), np.uint8); pc = np.zeros((chunk, Ht, 32), np.uint8)
    ph = np.zeros((chunk, Ht, 32), np.uint8); lc = np.zeros((chunk, 32), np.uint8); lh = np.zeros((chunk, 32), np.uint8)
    got = C.c_uint64()
    # warm-up: builds the generator tables of this policy's shapes
    w = mine[:min(64, len(mine))].copy()
    assert L.dapol_prove_batch(tree.subtree, len(w), p(w), H, policy, sd, p(out), out.nbytes, C.byref(got)) == 0, L.dapol_last_cuda_error()
    prove_s = verify_s = 0.0
    chunk_times = []
    n_ok = n_done = 0
    tampered_rejected = None
    barrier(); t_all = time.perf_counter()
    for s in range(0, len(mine), chunk):
        li = np.ascontiguousarray(mine[s:s + chunk]); k = len(li)
        t1 = time.perf_counter()
        rc = L.dapol_prove_batch(tree.subtree, k, p(li), H, policy, sd, p(out), out.nbytes, C.byref(got))
        assert rc == 0, (rc, L.dapol_last_cuda_error())
        t2 = time.perf_counter()
        rc = L.dapol_tree_paths(tree.subtree, k, p(li), p(pv), p(pr), p(pc), p(ph), p(lc), p(lh))  # the leaves' own proof nodes
        assert rc == 0, rc
        t3 = time.perf_counter()
        rc = L.dapol_verify_batch(ctx._h, 0, policy, k, p(rcom), p(rhash), p(lc), p(lh), p(out), p(off), p(ok))
        assert rc == 0, rc
        t4 = time.perf_counter()
        prove_s += t2 - t1; verify_s += t4 - t3
        chunk_times.append((round(t2 - t1, 4), round(t3 - t2, 4), round(t4 - t3, 4)))
        n_ok += int(ok[:k].sum()); n_done += k
        if s == 0:  # one tampered proof must be rejected
            bad = out[:size].copy(); bad[100] ^= 1
            o1 = np.zeros(1, np.uint8)
            L.dapol_verify_batch(ctx._h, 0, policy, 1, p(rcom), p(rhash), p(lc), p(lh), p(bad), p(off), p(o1))
            tampered_rejected = not bool(o1[0])
    barrier(); all_s = time.perf_counter() - t_all
    stats = torch.tensor([prove_s, verify_s, all_s, float(n_ok), float(n_done)], dtype=torch.float64, device="cuda")
    mx = stats.clone(); sm = stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        done, okc = int(sm[4].item()), int(sm[3].item())
        print(json.dumps({
            "config": f"C3: inclusion proofs for all users, 2^{ul} users, height {H}, aggregation_factor {H}, policy {'Splitting' if policy else 'Padding'}, {world} GPU(s)",
            "n_gpus": world, "proofs": done, "all_verified": okc == done, "tampered_rejected": tampered_rejected, "proof_bytes": int(size),
            "tree_build_s_e2e": build_s, "prove_s_max_rank": mx[0].item(), "verify_s_max_rank": mx[1].item(), "prove_plus_verify_wall_s": mx[2].item(),
            "prove_per_s": done / mx[0].item(), "verify_per_s": done / mx[1].item(), "prove_plus_verify_per_s": done / mx[2].item(),
            "root": root.com.hex()[:16], "chunk": chunk, "rank0_chunk_s_prove_paths_verify": chunk_times[:3] + chunk_times[-2:],
            "timing": "wall clock between barrier + synchronize pairs, max over ranks; host buffers (proofs D2H after prove, H2D for verify)"}), flush=True)
    tree.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
