"""BASELINE.json configs[2]: inclusion proofs (aggregated Bulletproofs prove + verify) for the users of a 2^20-user, height-32
tree.  Proves and verifies a batch of K users per policy through the C ABI (host buffers out / in), reports proofs/s and the
projected time for all users on 1 and 8 GPUs (per-user proofs are independent: no communication).
  python tools/c3_inclusion.py [users_log2=20] [height=32] [K=2048]"""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bench import AUDIT_SEED, PAD_SEED, synth_liabilities
from dapol_b200 import Context, Dapol, DapolProof, DapolProofNode

PROVE_SEED = hashlib.sha256(b"dapol-b200-prove").digest()


def main():
    ul = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    n = 1 << ul
    ctx = Context(0)
    t0 = time.perf_counter()
    tree = Dapol.new(ctx, 0, synth_liabilities(n), AUDIT_SEED, H, H, PAD_SEED)
    build_s = time.perf_counter() - t0
    root = tree.root()
    users = np.linspace(0, n - 1, K).astype(np.int64)
    leaves = np.array([tree.leaf_index_of(int(u)) for u in users], np.uint64)
    for policy, name in ((0, "Padding"), (1, "Splitting")):
        tree.policy = policy
        tree.generate_proofs(leaves[:8], PROVE_SEED)  # builds the generator tables of this policy's shapes
        t0 = time.perf_counter()
        proofs = tree.generate_proofs(leaves, PROVE_SEED)
        t1 = time.perf_counter()
        p = tree.paths(leaves)
        nodes = [DapolProofNode(p["leaf_comc"][i].tobytes(), p["leaf_hash"][i].tobytes()) for i in range(K)]
        t2 = time.perf_counter()
        ok = DapolProof.verify_many(ctx, root, nodes, proofs)
        t3 = time.perf_counter()
        bad = DapolProof(bytes(proofs[0].serialize()[:100]) + bytes([proofs[0].serialize()[100] ^ 1]) + proofs[0].serialize()[101:], 0, policy)
        rej = DapolProof.verify_many(ctx, root, nodes[:1], [bad])
        line = {"config": f"C3 inclusion proofs, 2^{ul} users, height {H}, aggregation_factor {H}, policy {name}", "batch": K,
                "proof_bytes": len(proofs[0].serialize()), "prove_per_s": K / (t1 - t0), "verify_per_s": K / (t3 - t2),
                "prove_plus_verify_per_s": K / ((t1 - t0) + (t3 - t2)), "all_verified": bool(ok.all()), "tampered_rejected": not bool(rej[0]),
                "tree_build_s_e2e": build_s,
                "projected_all_users_s": {"1gpu": n * ((t1 - t0) + (t3 - t2)) / K, "8gpu": n * ((t1 - t0) + (t3 - t2)) / K / 8},
                "timing": "wall clock around the C-ABI calls (host buffers: proofs D2H on prove, H2D on verify)"}
        print(json.dumps(line), flush=True)
    tree.close(); ctx.close()


if __name__ == "__main__":
    main()
