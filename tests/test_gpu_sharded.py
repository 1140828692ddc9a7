"""Sharded build on the GPU: 1, 2 and 4 ranks (processes) share the test box's one GPU, rendezvous over gloo,
and go through the real C ABI (dapol_leaves_derive_dev / _assign_dev, dapol_tree_build_shard_dev, dapol_tree_build_from_records,
dapol_tree_attach_top).  tests/sharded_worker.py checks, per rank: every node of the shard's subtree, the whole-tree root,
the id -> index map, byte-identical inclusion proofs (both policies) and their verification -- all against the oracle's
single-tree build of the same liabilities."""
import pytest

from test_sharded_cpu import run_world

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n,H,hash_id,extra", [(1, 300, 12, 0, ()), (2, 500, 14, 0, ()), (4, 700, 16, 1, ()), (2, 65, 9, 0, ("uneven",)),
                                                      (4, 5, 8, 0, ()), (4, 300, 13, 0, ("positional",)), (2, 40, 8, 1, ("positional", "uneven"))])
def test_sharded_build_matches_single_tree_oracle(world, n, H, hash_id, extra):
    run_world(world, "cuda", n, H, hash_id, extra, timeout=900)


@pytest.mark.parametrize("world,n,H,hash_id,extra", [(1, 300, 12, 0, ()), (2, 500, 14, 0, ()), (4, 700, 16, 1, ()), (2, 65, 9, 0, ("uneven",)),
                                                      (4, 5, 8, 0, ()), (4, 300, 13, 0, ("positional",)), (2, 40, 8, 1, ("positional", "uneven")),
                                                      (4, 128, 8, 0, ()), (2, 512, 10, 1, ("uneven",)), (8, 2000, 16, 0, ()), (4, 600, 14, 0, ("dup",)),
                                                      (2, 300, 40, 0, ()), (4, 200, 64, 1, ())])
def test_one_call_sharded_build_matches_single_tree_oracle(world, n, H, hash_id, extra):
    """dapol_sharded_build (the whole exchange protocol inside the library: all-to-all of claims, per-prefix collision
    resolution with only the losers travelling again, padding bases, subtree, root gather, top tree), its collectives on the
    host-provided transport.  Same checks as above; (4, 128, 8) and (2, 512, 10) are the densest legal trees (2^H = 2 N):
    a quarter of the users collide and re-hash for several rounds."""
    run_world(world, "cuda", n, H, hash_id, tuple(extra) + ("native",), timeout=900)
