"""Worker of the multi-process tests of dapol_b200.sharded (one process per rank, gloo rendezvous on 127.0.0.1).

  python tests/sharded_worker.py <rank> <world> <port> <engine: fake|cuda> <n_users> <height> <hash_id> [uneven]

`fake`: the C-ABI calls are replaced by a test double written with the CPU oracle's primitives (tests may use oracle/),
so the host logic -- slice bookkeeping, the record exchange, padding-stream bases, top-tree assembly -- is covered on CPU.
`cuda`: the real library on cuda:0 (every rank shares the one GPU of the test box; collectives still go through gloo).
Both check the sharded result against the oracle's single-tree build of the whole liability set and print "OK <root>".
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cref  # noqa: E402

PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
PROVE_SEED = hashlib.sha256(b"dapol-b200-prove").digest()
AUDIT_SEED = b"sharded-test"
L_ORDER = 2 ** 252 + 27742317777372353535851937790883648493


def liabilities(n, seed=7):
    rnd = np.random.RandomState(seed)
    ids = [b"user-%d" % i + bytes(rnd.randint(0, 256, rnd.randint(0, 5)).astype(np.uint8)) for i in range(n)]
    eids = [b"ext-%d" % (i * 7919) for i in range(n)]
    vals = rnd.randint(0, 1 << 31, n).astype(np.uint64)
    return ids, eids, vals


def seed_to_index(seed: bytes, H: int) -> int:
    return int.from_bytes(seed[:8], "big") >> (64 - H)


class FakeEngine:
    """Test double of dapol_b200.sharded.CudaEngine on CPU tensors, from the oracle's primitives."""
    device = None

    def __init__(self, global_tree, k, rank, hash_id, positional=False):
        self.g, self.k, self.rank, self.hash_id = global_tree, k, rank, hash_id
        self.positional = positional  # padding keyed by (level, index): no exchange of padding counts

    def to_dev(self, a, dtype):
        return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a).view(dtype).copy())

    def derive(self, hash_id, height, n, iid, io, eid, eo, audit_seed):  # mod.rs:338-386
        iid, io, eid, eo = (t.numpy() for t in (iid, io, eid, eo))
        audit = np.zeros((n, 32), np.uint8); seedst = np.zeros((n, 32), np.uint8); blind = np.zeros((n, 32), np.uint8)
        cand = np.zeros(n, np.int64)
        for i in range(n):
            a = cref.hash(hash_id, audit_seed + iid[io[i]:io[i + 1]].tobytes())
            e = eid[eo[i]:eo[i + 1]].tobytes()
            s = cref.hash(hash_id, cref.hash(hash_id, a + b"index_seed" + e))
            b = bytearray(cref.hash(hash_id, a + b"blind_seed" + e)); b[31] &= 0x7F
            audit[i] = np.frombuffer(a, np.uint8); seedst[i] = np.frombuffer(s, np.uint8); blind[i] = np.frombuffer(bytes(b), np.uint8)
            cand[i] = np.uint64(seed_to_index(s, height)).astype(np.int64)
        return tuple(torch.from_numpy(x) for x in (audit, seedst, cand, blind))

    def assign(self, hash_id, height, n_total, audit, seedst, cand, blind, values, prefix_bits, prefix, cap):  # mod.rs:345-349,408-441
        seen, used, final = set(), set(), []
        for i in range(n_total):
            a = audit[i].numpy().tobytes()
            assert a not in seen
            seen.add(a)
            s = seedst[i].numpy().tobytes()
            for _ in range(128):
                idx = seed_to_index(s, height)
                if idx not in used:
                    break
                s = cref.hash(hash_id, s)
            else:
                raise AssertionError("failed to map")
            used.add(idx); final.append(idx)
            cand[i] = int(np.uint64(idx).astype(np.int64))
        order = sorted(range(n_total), key=lambda i: final[i])
        mine = [i for i in order if (final[i] >> (height - prefix_bits)) == prefix] if prefix_bits else order
        mask = (1 << (height - prefix_bits)) - 1
        idx = torch.tensor([np.uint64(final[i] & mask).astype(np.int64) for i in mine], dtype=torch.int64)
        return idx, values[mine], blind[mine]

    def pad_counts(self, height, idx):
        counts = np.zeros(height + 1, np.uint64)
        cur = np.sort(idx.numpy().astype(np.uint64))
        for h in range(height, 0, -1):
            if len(cur) == 0:
                break
            par = np.unique(cur >> np.uint64(1))
            counts[h] = 2 * len(par) - len(cur)
            cur = par
        return counts

    def _expected_bases(self, height):
        """RNG block of this shard's first padding node per level, read off the oracle's single tree."""
        H = self.g.height
        base = np.zeros(height + 1, np.uint64)
        below = 0
        for h in range(height, 0, -1):
            lvl = self.g.level(self.k + h)
            pads = lvl["idx"][lvl["is_pad"] == 1]
            base[h] = below + int(((pads >> np.uint64(h)) < self.rank).sum())
            below += len(pads)
        return base, below

    def build_shard(self, hash_id, height, idx, values, blind, pad_seed, level_base):
        if self.positional:  # the shard's coordinates inside the whole tree: levels above it, index of its first node per level
            exp = np.array([self.k] + [self.rank << h for h in range(1, height + 1)], np.uint64)
            assert (np.asarray(level_base, np.uint64) == exp).all(), (level_base, exp)
        else:
            exp, _ = self._expected_bases(height)
            assert (np.asarray(level_base, np.uint64)[1:] == exp[1:]).all(), (level_base, exp)
        leaves = self.g.level(self.g.height)
        sel = (leaves["is_pad"] == 0) & ((leaves["idx"] >> np.uint64(height)) == self.rank)
        assert (leaves["idx"][sel] & np.uint64((1 << height) - 1) == idx.numpy().astype(np.uint64)).all()
        assert (leaves["v"][sel] == values.numpy().astype(np.uint64)).all()
        return ("subtree", self.g.get_node(self.k, self.rank))

    def root_record(self, tree):
        nd = tree[1]
        rec = np.zeros(232, np.uint8)  # the half-point part stays empty: the double adds compressed points
        rec[128:160] = np.frombuffer(nd["comc"], np.uint8); rec[160:192] = np.frombuffer(nd["hash"], np.uint8)
        rec[192:224] = np.frombuffer(nd["r"], np.uint8); rec[224:232] = np.frombuffer(int(nd["v"]).to_bytes(8, "little"), np.uint8)
        return rec

    def build_top(self, hash_id, height, idx, records, pad_seed, pad_base):
        if not self.positional:
            _, below = self._expected_bases(self.g.height - self.k)
            assert pad_base == below, (pad_base, below)
        cur = {int(i): dict(comc=r[128:160].tobytes(), hash=r[160:192].tobytes(), r=int.from_bytes(r[192:224].tobytes(), "little"),
                            v=int.from_bytes(r[224:232].tobytes(), "little")) for i, r in zip(idx, records)}
        ordinal = pad_base
        for lvl in range(height, 0, -1):  # smtree build order: level by level, left to right
            nxt = {}
            for x in sorted(cur):
                if (x ^ 1) not in cur:
                    r = cref.rng_scalar(pad_seed, x ^ 1, lvl) if self.positional else cref.rng_scalar(pad_seed, ordinal)
                    ordinal += 1
                    c = cref.commit(0, r)
                    cur[x ^ 1] = dict(comc=c, hash=cref.hash(hash_id, c), r=int.from_bytes(r, "little"), v=0)
            for x in sorted(cur):
                if x & 1:
                    continue
                l, r = cur[x], cur[x ^ 1]
                nxt[x >> 1] = dict(comc=cref.point_add(l["comc"], r["comc"]), hash=cref.hash(hash_id, l["comc"] + r["comc"] + l["hash"] + r["hash"]),
                                   r=(l["r"] + r["r"]) % L_ORDER, v=(l["v"] + r["v"]) & (2 ** 64 - 1))
            cur = nxt
        return ("top", cur[0] if height else cur[int(idx[0])])

    def attach(self, tree, top, prefix):
        assert prefix == self.rank

    def destroy(self, tree):
        pass

    def root_of(self, tree):
        from dapol_b200.api import DapolNode
        nd = tree[1]
        return DapolNode(nd["v"], (nd["r"] % L_ORDER).to_bytes(32, "little"), nd["comc"], nd["hash"])


def main():
    rank, world, port, engine, n, H, hash_id = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
    uneven = "uneven" in sys.argv[8:]
    positional = "positional" in sys.argv[8:]  # padding keyed by (level, index): SURVEY 8(f) N3
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from dapol_b200.sharded import Comm, ShardedDapol
    comm = Comm()
    k = (world - 1).bit_length()
    # the oracle's single tree over the whole liability set
    ids, eids, vals = liabilities(n)
    if "dup" in sys.argv[8:]:
        # duplicated internal ids on different ranks (mod.rs:345-349): every rank must report DuplicatedInternalId at the input
        # position the sequential reference fails at (the oracle's err_pos) -- the exact pass behind the prefix screen decides
        ids[n - 2] = ids[1]; ids[n // 2] = ids[3]
        ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
        rc, _, _, err_pos = cref.derive_leaves(hash_id, ib, io, eb, eo, AUDIT_SEED, H)
        assert rc == 4
        cuts = [0] + [(n * (r + 1)) // world for r in range(world)]
        lo, hi = cuts[rank], cuts[rank + 1]
        sl = (*cref.pack_ids(ids[lo:hi]), *cref.pack_ids(eids[lo:hi]), vals[lo:hi])
        from dapol_b200 import Context, DapolError
        from dapol_b200.sharded import CudaEngine, NativeComm
        E = CudaEngine(Context(0))
        try:
            ShardedDapol.new(E, comm, hash_id, sl, AUDIT_SEED, H, 1, PAD_SEED, native=NativeComm(E.ctx, comm, backend="torch"))
            raise AssertionError("duplicate ids not detected")
        except DapolError as e:
            assert (e.code, e.detail) == (4, err_pos), (e.code, e.detail, err_pos)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        print("OK dup@%d" % err_pos, flush=True)
        return
    ib, io = cref.pack_ids(ids); eb, eo = cref.pack_ids(eids)
    rc, gidx, gbl, _ = cref.derive_leaves(hash_id, ib, io, eb, eo, AUDIT_SEED, H)
    assert rc == 0
    order = np.argsort(gidx, kind="stable")
    g = cref.Tree(hash_id, H, gidx[order], vals[order], gbl[order], PAD_SEED, positional=positional)
    # this rank's slice of the input
    cuts = [0] + [(n * (r + 1)) // world for r in range(world)]
    if uneven and world > 1:
        cuts = [0, 1] + [1 + ((n - 1) * r) // (world - 1) for r in range(1, world)]
    lo, hi = cuts[rank], cuts[rank + 1]
    sl = (*cref.pack_ids(ids[lo:hi]), *cref.pack_ids(eids[lo:hi]), vals[lo:hi])
    sl = (sl[0], sl[1], sl[2], sl[3], sl[4])
    if engine == "fake":
        E = FakeEngine(g, k, rank, hash_id, positional)
    else:
        from dapol_b200 import Context
        from dapol_b200.sharded import CudaEngine
        E = CudaEngine(Context(0))
        E.set_padding_mode(positional)
    agg = min(H, 3)
    native = None
    if engine == "cuda" and "native" in sys.argv[8:]:
        # the one-call path (dapol_sharded_build): the library runs the whole exchange protocol; its collectives go through
        # the host-provided transport (dapol_comm_ops) on top of this test's gloo group, the ranks share the box's one GPU
        from dapol_b200.sharded import NativeComm
        native = NativeComm(E.ctx, comm, backend="torch")
    t = ShardedDapol.new(E, comm, hash_id, sl, AUDIT_SEED, H, agg, PAD_SEED, native=native)
    root, oroot = t.root_raw(), g.root()
    assert (root.value, root.com, root.hash) == (oroot["v"], oroot["comc"], oroot["hash"]), "sharded root != oracle single-tree root"
    assert int.from_bytes(root.blinding, "little") % L_ORDER == int.from_bytes(oroot["r"], "little") % L_ORDER
    for pos in (0, n // 2, n - 1) + tuple(range(lo, min(hi, lo + 3))):
        if native is None or lo <= pos < hi:  # the one-call build keeps the id -> index map of the rank's own slice
            assert t.leaf_index_of(pos) == int(gidx[pos])
        else:
            assert t.leaf_index_of(pos) is None
    if engine == "cuda":
        import ctypes as C
        from dapol_b200 import _ffi
        from dapol_b200.api import DapolProof, DapolProofNode
        L = _ffi.lib()
        Hs = H - k
        if t.subtree:  # every node of the shard's subtree against the oracle's nodes below (level k, idx rank)
            for h in range(Hs + 1):
                nn = L.dapol_tree_level_size(t.subtree, h)
                idx = np.zeros(nn, np.uint64); v = np.zeros(nn, np.uint64); c = np.zeros((nn, 32), np.uint8); hs = np.zeros((nn, 32), np.uint8)
                pad = np.zeros(nn, np.uint8)
                assert L.dapol_tree_level_copy(t.subtree, h, idx.ctypes.data, v.ctypes.data, None, c.ctypes.data, hs.ctypes.data, pad.ctypes.data) == 0
                o = g.level(k + h)
                sel = (o["idx"] >> np.uint64(h)) == rank if h < 64 else np.ones(len(o["idx"]), bool)
                assert (o["idx"][sel] & np.uint64((1 << h) - 1 if h else 0) == idx).all(), h
                for key, arr in (("v", v), ("comc", c), ("hash", hs), ("is_pad", pad)):
                    assert (o[key][sel] == arr).all(), (h, key)
            mine = [int(x) for x in gidx if (int(x) >> Hs) == rank][:3]
            for policy in (0, 1):
                t.policy = policy
                proofs = t.generate_proofs(mine, PROVE_SEED)
                assert proofs is not None
                for x, p in zip(mine, proofs):
                    assert p.serialize() == g.prove_inclusion(x, agg, policy, PROVE_SEED), "sharded proof bytes != oracle"
                    leaf = g.get_node(H, x)
                    assert cref.verify_inclusion(hash_id, policy, p.serialize(), oroot["comc"], oroot["hash"], leaf["comc"], leaf["hash"])
                leaves = [DapolProofNode(g.get_node(H, x)["comc"], g.get_node(H, x)["hash"]) for x in mine]
                ok = DapolProof.verify_many(E.ctx, t.root(), leaves, proofs)
                assert ok.all()
            other = [int(x) for x in gidx if (int(x) >> Hs) != rank][:1]
            if other:
                assert t.generate_proofs(other, PROVE_SEED) is None
    t.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print("OK", root.com.hex()[:16], flush=True)


if __name__ == "__main__":
    main()
