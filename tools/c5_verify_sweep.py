"""BASELINE.json configs[4]: batch range-proof verification sweep (1k - 1M proofs, 32- and 64-bit ranges) on one GPU, with the
host CPU (oracle port, all threads) timed beside it on a bounded sample.  1 % of the proofs are corrupted; the GPU verdicts
must be exactly "corrupted <=> rejected", and a sample of them is cross-checked against the oracle verifier (reject parity).
Prints one JSON line per (nbits, K, group).   python tools/c5_verify_sweep.py [max_log2=20] [groups=0] [--gpus handled by torchrun: each rank = K proofs]
groups: comma list of dapol_ctx_set_verify_mode group sizes (0 = every proof on its own by Straus; G > 1 = bucket-method groups with
per-proof fallback) -- the head-to-head of the two verifiers on a batch with 1 % bad proofs."""
import concurrent.futures as cf
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dapol_b200 import Context, _ffi
from oracle import cref

SEED = hashlib.sha256(b"dapol-b200").digest()


def main():
    max_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    groups = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    L = _ffi.lib()
    import ctypes as C
    seed = (C.c_uint8 * 32).from_buffer_copy(SEED)
    cores = os.cpu_count() or 1
    for nbits in (64, 32):
        Kmax = 1 << max_log2
        size = L.dapol_rangeproof_size(nbits, 1)
        rng = np.random.default_rng(99 + nbits + rank)
        vals = rng.integers(0, 1 << (nbits - 1), size=(Kmax, 1), dtype=np.uint64)
        bl = rng.integers(0, 256, size=(Kmax, 1, 32), dtype=np.uint8); bl[:, :, 31] &= 0x0F
        coms = ctx.commit_batch(vals[:, 0], bl[:, 0])
        tv, tb, tc = (torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev) for a in (vals, bl, coms))
        ts = torch.arange(Kmax, dtype=torch.int64, device=dev); tz = torch.zeros(Kmax, dtype=torch.int64, device=dev)
        d_proofs = torch.empty(Kmax * size, dtype=torch.uint8, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        assert L.dapol_rangeproof_prove_batch_dev(ctx._h, nbits, 1, Kmax, tv.data_ptr(), tb.data_ptr(), seed, ts.data_ptr(), tz.data_ptr(), d_proofs.data_ptr()) == 0
        ev[1].record(); torch.cuda.synchronize()
        prove_ms = ev[0].elapsed_time(ev[1])
        bad = rng.random(Kmax) < 0.01
        pv = d_proofs.view(Kmax, size)
        bad_t = torch.from_numpy(bad).to(dev)
        byte = torch.from_numpy(rng.integers(0, size, Kmax)).to(dev)
        rows = torch.nonzero(bad_t).squeeze(1)
        pv[rows, byte[rows]] ^= 1 << 3
        d_ok = torch.empty(Kmax, dtype=torch.uint8, device=dev)
        for lg, G in [(lg, G) for lg in range(10, max_log2 + 1, 2) for G in groups]:
            K = 1 << lg
            ctx.set_verify_mode(G)
            f0 = ctx.verify_fallbacks
            L.dapol_rangeproof_verify_batch_dev(ctx._h, nbits, 1, K, d_proofs.data_ptr(), size, tc.data_ptr(), d_ok.data_ptr())  # warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ev[0].record()
            assert L.dapol_rangeproof_verify_batch_dev(ctx._h, nbits, 1, K, d_proofs.data_ptr(), size, tc.data_ptr(), d_ok.data_ptr()) == 0
            ev[1].record(); torch.cuda.synchronize()
            t = torch.tensor([ev[0].elapsed_time(ev[1])], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            ok = d_ok[:K].cpu().numpy().astype(bool)
            exact = bool((ok == ~bad[:K]).all())
            if world > 1:  # every rank checks the verdicts of its own batch
                e = torch.tensor([1 if exact else 0], device=dev)
                dist.all_reduce(e, op=dist.ReduceOp.MIN)
                exact = bool(e.item())
            line = {"config": "C5 batch range-proof verification", "nbits": nbits, "m": 1, "proofs_per_gpu": K, "n_gpus": world, "verify_ms": ms,
                    "verifies_per_s": world * K / ms * 1e3, "corrupted": int(bad[:K].sum()), "verdicts_exact": exact,
                    "verify_group": G, "reverified_per_call": int(ctx.verify_fallbacks - f0) // 2}
            if lg == 10 and rank == 0 and G == groups[0]:  # CPU leg + reject parity on the first 1024 proofs
                hp = pv[:1024].cpu().numpy(); hc = coms[:1024]
                t0 = time.perf_counter()
                with cf.ThreadPoolExecutor(cores) as ex:
                    cpu_ok = list(ex.map(lambda i: cref.rp_verify(hp[i].tobytes(), [hc[i].tobytes()], nbits), range(1024)))
                dt = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": 1024 / dt, "unit": "verifies/s", "cores": cores, "kind": "port", "sample": "first 1024 proofs of the batch"}
                line["reject_parity_with_oracle"] = bool((np.array(cpu_ok) == ok[:1024]).all())
                line["prove_per_s_for_setup"] = Kmax / prove_ms * 1e3
            if rank == 0:
                print(json.dumps(line), flush=True)
        del d_proofs, pv, tv, tb, tc
    ctx.close()


if __name__ == "__main__":
    main()
