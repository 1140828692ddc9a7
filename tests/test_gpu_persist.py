"""SURVEY 8(f) N4 -- tree persistence and streamed proof output (no reference counterpart: the crate keeps the tree in memory and
leaves "write the proofs to a local file" as a TODO, src/dapol/mod.rs:250).  The bar is the same as for the build: a loaded tree
is the saved tree bit for bit (every level, the id -> index map, inclusion-proof bytes), and the oracle's tree."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PAD_SEED = hashlib.sha256(b"dapol-b200").digest()
PROVE_SEED = hashlib.sha256(b"persist").digest()


@pytest.fixture(scope="module")
def ctx():
    from dapol_b200 import Context
    c = Context(0, 15)
    c.set_rangeproof_window(8)
    yield c
    c.close()


def _liabilities(n):
    return [(b"user-%d" % i, b"ext-%d" % (i * 7), (i * 2654435761) & 0xFFFFFFFF) for i in range(n)]


@pytest.mark.parametrize("hash_id,H,n,policy", [(0, 12, 300, 0), (1, 9, 40, 1), (0, 1, 1, 0)])
def test_saved_tree_loads_bit_identical(ctx, tmp_path, hash_id, H, n, policy):
    from dapol_b200 import Dapol, DapolProof
    agg = min(H, 4)
    t = Dapol.new(ctx, hash_id, _liabilities(n), b"audit", H, agg, PAD_SEED, policy=policy)
    path = str(tmp_path / "tree.dapol")
    t.save(path)
    assert os.path.getsize(path) > 113 * t.num_nodes
    u = Dapol.load(ctx, path, agg, policy)
    assert (u.height, u.hash_id, u.num_nodes, u.num_padding) == (t.height, t.hash_id, t.num_nodes, t.num_padding)
    for h in range(H + 1):
        a, b = t.level(h), u.level(h)
        for k in a:
            assert (a[k] == b[k]).all(), (h, k)
    ra, rb = t.root_raw(), u.root_raw()
    assert (ra.value, ra.blinding, ra.com, ra.hash) == (rb.value, rb.blinding, rb.com, rb.hash)
    leaves = [t.leaf_index_of(i) for i in range(n)]
    assert leaves == [u.leaf_index_of(i) for i in range(n)]
    pick = leaves[:: max(1, n // 7)]
    pa, pb = t.generate_proofs(pick, PROVE_SEED), u.generate_proofs(pick, PROVE_SEED)
    assert [x.serialize() for x in pa] == [x.serialize() for x in pb]
    paths = u.paths(pick)
    from dapol_b200 import DapolProofNode
    nodes = [DapolProofNode(paths["leaf_comc"][i].tobytes(), paths["leaf_hash"][i].tobytes()) for i in range(len(pick))]
    assert DapolProof.verify_many(ctx, u.root(), nodes, pb).all()
    # streamed proofs = the batch call's bytes, proof i at i * size
    out = str(tmp_path / "proofs.bin")
    size = u.generate_proofs_to_file(pick, PROVE_SEED, out, chunk=3)
    blob = open(out, "rb").read()
    assert len(blob) == size * len(pick) and [blob[i * size:(i + 1) * size] for i in range(len(pick))] == [x.serialize() for x in pa]
    t.close(); u.close()


def test_bad_files_are_io_errors(ctx, tmp_path):
    from dapol_b200 import Dapol, DapolError
    t = Dapol.new(ctx, 0, _liabilities(20), b"audit", 8, 4, PAD_SEED)
    path = str(tmp_path / "tree.dapol")
    t.save(path)
    blob = open(path, "rb").read()
    for bad in (blob[:-9], blob[:100], b"not a tree" * 20, blob[:8] + b"\x07" + blob[9:]):
        p = str(tmp_path / "bad.dapol")
        open(p, "wb").write(bad)
        with pytest.raises(DapolError) as e:
            Dapol.load(ctx, p, 4)
        assert e.value.code == 21
    # consistent sizes but corrupted slot maps / indexes (a crafted file): rejected before any kernel can index out of bounds.
    # Layout after the 64-byte header, the 9 x 32-byte level table and the 128-byte root point: idx u64[T], v, r, com, hash, is_pad, pos u32[]
    T = t.num_nodes
    idx_off = 64 + 9 * 32 + 128
    pos_off = idx_off + T * (8 + 8 + 32 + 32 + 32 + 1)
    for off, val in ((idx_off + 8 * 3, b"\xff" * 8), (idx_off + 8 * (T - 1), b"\x00" * 8), (pos_off + 4 * 64, b"\xff\xff\xff\x7f"),
                     (pos_off + 4 * 65, b"\x00\x00\x00\x00")):
        bad = bytearray(blob); bad[off:off + len(val)] = val
        p = str(tmp_path / "crafted.dapol")
        open(p, "wb").write(bytes(bad))
        with pytest.raises(DapolError) as e:
            Dapol.load(ctx, p, 4)
        assert e.value.code == 21, off
    with pytest.raises(DapolError) as e:
        Dapol.load(ctx, str(tmp_path / "missing.dapol"), 4)
    assert e.value.code == 21
    with pytest.raises(DapolError) as e:
        t.save(str(tmp_path / "no-such-dir" / "x"))
    assert e.value.code == 21
    t.close()
