#!/usr/bin/env python3
"""Generates tests/golden/dapol_golden.json from the big-int Python restatement (oracle/pyref.py).

The Rust reference cannot run in this container (no cargo, un-vendored crates) and its own tests hold no byte-level
vectors (all thread_rng), so these are ORACLE-generated goldens under the seeded-RNG contract of include/dapol_b200.h:
they freeze the bytes that two independent restatements (pyref big-int, oracle/c 51-bit limbs) and the CUDA path must
all reproduce.  The inputs that ARE pinned by the reference (ids a,b,c,d -> leaves 7,12,2,4; root value 26;
src/dapol/tests.rs:17-84) are the first fixture.  Re-run:  python tests/golden/gen_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyref as o  # noqa: E402

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
PROVE_SEED = hashlib.sha256(b"dapol-b200-prove").digest()


def tree_fixture(name, hash_id, H, liabilities, audit_seed, proofs_for=(), agg=None):
    leaves = o.derive_leaves(hash_id, liabilities, audit_seed, H)
    t = o.build_tree(hash_id, H, [(i, o.node_new(hash_id, v, r)) for i, v, r in leaves], PAD_SEED)
    levels = []
    for h in range(H + 1):
        lvl = t.levels[h]
        levels.append([{"idx": i, "v": lvl[i].v, "com": lvl[i].comc.hex(), "hash": lvl[i].hash.hex(), "pad": int(i in t.is_pad[h])}
                       for i in sorted(lvl)])
    fx = {"name": name, "hash_id": hash_id, "height": H, "audit_seed": audit_seed.hex(),
          "liabilities": [[a.hex(), b.hex(), v] for a, b, v in liabilities], "leaf_idx": [i for i, _, _ in leaves],
          "blindings": [o.sc_bytes(r % o.L).hex() if r < o.L else r.to_bytes(32, "little").hex() for _, _, r in leaves],
          "pad_seed": PAD_SEED.hex(), "levels": levels, "proofs": []}
    for pos, policy in proofs_for:
        leaf = leaves[pos][0]
        p = o.prove_inclusion(t, leaf, agg, policy, PROVE_SEED)
        node = t.levels[H][leaf]
        assert o.verify_inclusion(hash_id, p, policy, (t.root.comc, t.root.hash), (node.comc, node.hash))
        fx["proofs"].append({"leaf_idx": leaf, "policy": policy, "aggregation_factor": agg, "seed": PROVE_SEED.hex(), "bytes": p.hex()})
    return fx


def node_tree_fixture(name, hash_id, H, leaf_idx, values, blind, proofs_for=(), batches=()):
    """Dapol::new_blank + build from ready nodes (mod.rs:196-208) -- the only way to a 64-byte digest (Blake2b, src/tests.rs:104-105) --
    with single proofs and batch proofs (generate_proof_batch, mod.rs:172-190)."""
    t = o.build_tree(hash_id, H, [(i, o.node_new(hash_id, v, r)) for i, v, r in zip(leaf_idx, values, blind)], PAD_SEED)
    levels = [[{"idx": i, "v": t.levels[h][i].v, "com": t.levels[h][i].comc.hex(), "hash": t.levels[h][i].hash.hex(), "pad": int(i in t.is_pad[h])}
               for i in sorted(t.levels[h])] for h in range(H + 1)]
    fx = {"name": name, "hash_id": hash_id, "height": H, "leaf_idx": list(leaf_idx), "values": list(values), "blindings": [o.sc_bytes(r).hex() for r in blind],
          "pad_seed": PAD_SEED.hex(), "levels": levels, "proofs": [], "batch_proofs": []}
    root = (t.root.comc, t.root.hash)
    for leaf, policy, agg in proofs_for:
        p = o.prove_inclusion(t, leaf, agg, policy, PROVE_SEED)
        node = t.levels[H][leaf]
        assert o.verify_inclusion(hash_id, p, policy, root, (node.comc, node.hash))
        fx["proofs"].append({"leaf_idx": leaf, "policy": policy, "aggregation_factor": agg, "seed": PROVE_SEED.hex(), "bytes": p.hex()})
    for leaves, policy, agg in batches:
        p = o.prove_inclusion_batch(t, list(leaves), agg, policy, PROVE_SEED)
        assert o.verify_inclusion_batch(hash_id, p, policy, root, [(t.levels[H][x].comc, t.levels[H][x].hash) for x in leaves])
        fx["batch_proofs"].append({"leaf_idxs": list(leaves), "policy": policy, "aggregation_factor": agg, "seed": PROVE_SEED.hex(), "bytes": p.hex()})
    return fx


def main():
    out = {"generator": "tests/golden/gen_golden.py (oracle/pyref.py)", "trees": [], "range_proofs": [], "node_trees": []}
    kat = [(b"a", b"w", 3), (b"b", b"x", 5), (b"c", b"y", 7), (b"d", b"z", 11)]
    out["trees"].append(tree_fixture("reference KAT src/dapol/tests.rs:17-84 (Blake2s, H=4)", 1, 4, kat, b"test",
                                     proofs_for=[(0, 0), (1, 1)], agg=2))
    liab = [(b"user-%d" % i, b"salt-%d" % (i * 31), (i * 2654435761) & 0xFFFFFFFF) for i in range(12)]
    out["trees"].append(tree_fixture("12 users, blake3, H=8", 0, 8, liab, b"golden-audit-seed", proofs_for=[(5, 0)], agg=3))
    for m, nbits in ((1, 64), (2, 64), (1, 32)):
        values = [(0xDEADBEEFCAFE + 977 * j) % (1 << nbits) for j in range(m)]
        blind = [int.from_bytes(hashlib.sha256(b"blind-%d" % j).digest(), "little") % o.L for j in range(m)]
        proof = o.rp_prove(values, blind, o.ScalarRng(PROVE_SEED, stream=3, base=5), nbits)
        coms = [o.compress(o.pedersen_commit(v, r)) for v, r in zip(values, blind)]
        assert o.rp_verify(proof, coms, nbits)
        out["range_proofs"].append({"nbits": nbits, "m": m, "values": values, "blindings": [o.sc_bytes(r).hex() for r in blind],
                                    "seed": PROVE_SEED.hex(), "stream": 3, "base_block": 5, "commitments": [c.hex() for c in coms],
                                    "proof": proof.hex()})
    idx = [1, 4, 5, 11, 18, 19, 30]
    vals = [(i * 2654435761) & 0xFFFFFFFF for i in range(1, 8)]
    blind = [int.from_bytes(hashlib.sha256(b"node-blind-%d" % j).digest(), "little") % o.L for j in range(7)]
    out["node_trees"].append(node_tree_fixture("7 nodes, Blake2b 64-byte digests, H=5", o.HASH_BLAKE2B, 5, idx, vals, blind,
                                               proofs_for=[(11, 1, 1)], batches=[((4, 5, 18), 1, 1)]))
    out["node_trees"].append(node_tree_fixture("7 nodes, blake3, H=5, batch proof", o.HASH_BLAKE3, 5, idx, vals, blind,
                                               batches=[((1, 19, 30), 0, 0)]))
    json.dump(out, open(os.path.join(HERE, "dapol_golden.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "dapol_golden.json"))


if __name__ == "__main__":
    main()
