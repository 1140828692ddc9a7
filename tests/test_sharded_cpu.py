"""Host logic of the sharded (one-process-per-GPU) build on CPU: world_size-2 and -4 gloo runs of
dapol_b200.sharded with the C-ABI calls replaced by a test double made of the oracle's primitives
(tests/sharded_worker.py), checked against the oracle's single-tree root; plus the padding-stream
base arithmetic on its own."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(world, engine, n, H, hash_id, extra=(), timeout=600):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "sharded_worker.py"), str(r), str(world), str(port), engine, str(n), str(H),
                               str(hash_id), *extra], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"
    roots = {out.strip().splitlines()[-1] for out in outs}
    assert len(roots) == 1 and next(iter(roots)).startswith("OK "), outs
    return roots


@pytest.mark.parametrize("world,n,H,hash_id,extra", [(2, 48, 9, 0, ()), (4, 40, 10, 1, ()), (2, 33, 8, 0, ("uneven",)), (1, 20, 7, 0, ()),
                                                      (4, 36, 9, 0, ("positional",))])
def test_sharded_host_logic_gloo(world, n, H, hash_id, extra):
    run_world(world, "fake", n, H, hash_id, extra)


def test_shard_pad_bases():
    from dapol_b200.sharded import shard_pad_bases
    # 2 shards, subtree height 3; counts[r][h]
    counts = np.array([[0, 1, 2, 3], [0, 0, 4, 5]], np.uint64)
    b0, top0 = shard_pad_bases(counts, 0, pad_base=10)
    b1, top1 = shard_pad_bases(counts, 1, pad_base=10)
    # leaf level (h=3) first: shard 0 then shard 1; then h=2; then h=1
    assert list(b0[1:]) == [10 + 8 + 6, 10 + 8, 10] and list(b1[1:]) == [10 + 8 + 6 + 1, 10 + 8 + 2, 10 + 3]
    assert top0 == top1 == 10 + 15
