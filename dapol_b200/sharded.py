"""One DAPOL+ tree over several GPUs, one process per GPU.

The reference builds its tree on one CPU thread (src/dapol/mod.rs:100-128); it has no multi-device
counterpart, so this module keeps the reference's names (`new`, `root`, `root_raw`, `generate_proof*`)
on an object every rank holds.  With 2^k ranks the tree splits at level k:

  1. every rank hashes ITS slice of the liabilities (build_leaf_nodes stage 1, mod.rs:338-386);
  2. the per-user records (112 B) are all-gathered over NCCL/NVLink and every rank runs the global
     duplicate / shuffle_index rules (mod.rs:345-349, 408-441) and keeps the sorted leaves whose index
     starts with its k-bit prefix;
  3. the per-level padding counts are all-gathered so each shard draws its padding blindings from the
     blocks of the seeded stream the single-tree build would use (level H..1, left to right);
  4. each rank builds the height-(H-k) subtree below node `rank` of level k -- no data-path collective;
  5. the 2^k subtree-root records (232 B) are all-gathered and every rank builds the top k levels.

Result: bit-identical root, nodes and proofs with `Dapol.new` on one GPU.  The CUDA work goes through
the C ABI (include/dapol_b200.h, "sharded build"); torch is used for device buffers and collectives only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi
from .api import DapolError, DapolNode, DapolProof, DapolProofNode, POLICY_PADDING, _check, _p

RECORD_BYTES = 232


class Comm:
    """torch.distributed plumbing (nccl on GPUs, gloo in the CPU tests); world 1 needs no process group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.group = group
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.backend = self.dist.get_backend(group) if self.dist else "none"

    def all_gather_host(self, arr: np.ndarray, device=None) -> np.ndarray:
        """Small fixed-size array -> [world, ...] on the host."""
        if self.world == 1:
            return arr[None].copy()
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy())
        if self.backend == "nccl":
            t = t.to(device)
        out = torch.empty((self.world,) + t.shape, dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group) if self.backend == "nccl" else \
            self.dist.all_gather(list(out.unbind(0)), t, group=self.group)
        return out.cpu().numpy().view(arr.dtype).reshape((self.world,) + arr.shape)

    def all_gather_rows(self, t: torch.Tensor, counts) -> torch.Tensor:
        """Concatenate per-rank row blocks (counts[r] rows from rank r) in rank order, on t's device."""
        if self.world == 1:
            return t
        dev = t.device
        nmax = int(max(counts))
        if t.shape[0] < nmax:
            pad = torch.zeros((nmax - t.shape[0],) + t.shape[1:], dtype=t.dtype, device=dev)
            t = torch.cat([t, pad])
        src = t.contiguous() if self.backend == "nccl" else t.cpu().contiguous()
        out = torch.empty((self.world * nmax,) + t.shape[1:], dtype=t.dtype, device=src.device)
        if self.backend == "nccl":
            self.dist.all_gather_into_tensor(out, src, group=self.group)
        else:
            self.dist.all_gather(list(out.view((self.world, nmax) + t.shape[1:]).unbind(0)), src, group=self.group)
        out = out.to(dev)
        if all(int(c) == nmax for c in counts):
            return out
        blocks = out.view((self.world, nmax) + t.shape[1:])
        return torch.cat([blocks[r, : int(counts[r])] for r in range(self.world)])


class _DevView:
    """Zero-copy torch view of raw device memory handed to a transport callback."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _dev_view(ptr, nbytes, device):
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


class NativeComm:
    """dapol_comm (include/dapol_b200.h) for dapol_sharded_build.
    `nccl`: the library's own NCCL communicator over NVLink (the 128-byte unique id travels through torch.distributed);
    `torch`: the host-provided transport (dapol_comm_ops) on top of torch.distributed -- NCCL collectives on the device
    buffers when the process group is nccl, host-staged gloo collectives otherwise (several ranks sharing one test GPU)."""

    def __init__(self, ctx, comm: Comm, backend: str = "nccl"):
        L = _ffi.lib()
        self.comm, self.ctx, self.h = comm, ctx, C.c_void_p()
        self.device = torch.device("cuda", ctx.device)
        dist = comm.dist
        if backend == "nccl":
            ident = np.zeros(128, np.uint8)
            if comm.rank == 0:
                _check(L.dapol_comm_nccl_unique_id(_p(ident)))
            if comm.world > 1:
                box = [ident.tobytes()]
                dist.broadcast_object_list(box, src=0, group=comm.group)
                ident = np.frombuffer(box[0], np.uint8).copy()
            _check(L.dapol_comm_nccl_create(ctx._h, _p(ident), comm.rank, comm.world, C.byref(self.h)))
        else:
            self._ag = _ffi.ALL_GATHER_FN(self._all_gather)
            self._a2a = _ffi.ALL_TO_ALL_FN(self._all_to_all)
            self._ops = _ffi.CommOps(None, self._ag, self._a2a)
            _check(L.dapol_comm_create(comm.rank, comm.world, C.byref(self._ops), C.byref(self.h)))

    # transport callbacks: device pointers in, ordered on the given stream (here: synchronise, do it, return when done)
    def _all_gather(self, user, d_send, d_recv, nbytes, stream):
        try:
            dist, W = self.comm.dist, self.comm.world
            torch.cuda.synchronize(self.device)
            src = _dev_view(d_send, nbytes, self.device)
            dst = _dev_view(d_recv, nbytes * W, self.device)
            if W == 1:
                dst.copy_(src)
            elif self.comm.backend == "nccl":
                dist.all_gather_into_tensor(dst, src, group=self.comm.group)
            else:
                h = src.cpu()
                out = [torch.empty_like(h) for _ in range(W)]
                dist.all_gather(out, h, group=self.comm.group)
                dst.copy_(torch.cat(out))
            torch.cuda.synchronize(self.device)
            return 0
        except Exception as e:  # never unwind through the C frames
            print("dapol transport all_gather failed:", e, flush=True)
            return 19

    def _all_to_all(self, user, d_send, send_off, send_bytes, d_recv, recv_off, recv_bytes, stream):
        try:
            dist, W = self.comm.dist, self.comm.world
            torch.cuda.synchronize(self.device)
            so, sb = [int(send_off[r]) for r in range(W)], [int(send_bytes[r]) for r in range(W)]
            ro, rb = [int(recv_off[r]) for r in range(W)], [int(recv_bytes[r]) for r in range(W)]
            # the library packs both buffers densely in rank order (offsets = running sums): one all_to_all_single
            assert so == [sum(sb[:r]) for r in range(W)] and ro == [sum(rb[:r]) for r in range(W)]
            onhost = self.comm.backend != "nccl"
            empty = torch.empty(0, dtype=torch.uint8, device=self.device)
            src = _dev_view(d_send, sum(sb), self.device) if sum(sb) else empty
            dst = _dev_view(d_recv, sum(rb), self.device) if sum(rb) else empty
            if W == 1:
                dst.copy_(src)
            elif onhost:
                out = torch.empty(sum(rb), dtype=torch.uint8)
                dist.all_to_all_single(out, src.cpu(), rb, sb, group=self.comm.group)
                dst.copy_(out)
            else:
                dist.all_to_all_single(dst, src, rb, sb, group=self.comm.group)
            torch.cuda.synchronize(self.device)
            return 0
        except Exception as e:
            print("dapol transport all_to_all failed:", e, flush=True)
            return 19

    def close(self):
        if getattr(self, "h", None) and self.h:
            _ffi.lib().dapol_comm_destroy(self.h)
            self.h = None

    __del__ = close


class CudaEngine:
    """The C-ABI calls of the sharded build, on torch device tensors."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)
        self.L = _ffi.lib()
        self.last_shard_times = None  # device time per phase of the last dapol_tree_build_shard_dev on this engine
        self.positional = False       # padding blindings keyed by position instead of the creation-order stream

    def set_padding_mode(self, positional: bool):
        self.ctx.set_padding_mode(positional)
        self.positional = bool(positional)

    def to_dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(self.device)
        return torch.from_numpy(np.ascontiguousarray(a).view(dtype)).to(self.device)

    def derive(self, hash_id, height, n, iid, io, eid, eo, audit_seed):
        dev = self.device
        if n == 0:  # a rank may hold an empty slice
            z = torch.empty((0, 32), dtype=torch.uint8, device=dev)
            return z, z.clone(), torch.empty(0, dtype=torch.int64, device=dev), z.clone()
        audit = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        seedst = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        blind = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        cand = torch.empty(n, dtype=torch.int64, device=dev)
        aseed = (C.c_uint8 * max(len(audit_seed), 1)).from_buffer_copy(audit_seed or b"\0")
        torch.cuda.current_stream(dev).synchronize()
        _check(self.L.dapol_leaves_derive_dev(self.ctx._h, hash_id, height, n, iid.data_ptr(), io.data_ptr(), eid.data_ptr(), eo.data_ptr(),
                                              aseed, len(audit_seed), audit.data_ptr(), seedst.data_ptr(), cand.data_ptr(), blind.data_ptr()))
        return audit, seedst, cand, blind

    def assign(self, hash_id, height, n_total, audit, seedst, cand, blind, values, prefix_bits, prefix, cap):
        dev = self.device
        torch.cuda.current_stream(dev).synchronize()
        while True:
            o_idx = torch.empty(cap, dtype=torch.int64, device=dev)
            o_val = torch.empty(cap, dtype=torch.int64, device=dev)
            o_bl = torch.empty((cap, 32), dtype=torch.uint8, device=dev)
            n_out, err = C.c_uint64(), C.c_uint64()
            torch.cuda.current_stream(dev).synchronize()
            rc = self.L.dapol_leaves_assign_dev(self.ctx._h, hash_id, height, n_total, audit.data_ptr(), seedst.data_ptr(), cand.data_ptr(),
                                                blind.data_ptr(), values.data_ptr(), prefix_bits, prefix, o_idx.data_ptr(), o_val.data_ptr(),
                                                o_bl.data_ptr(), cap, C.byref(n_out), C.byref(err))
            if rc == 18:  # buffer too small: the fix-point is already reached, the second pass only re-sorts
                cap = int(n_out.value)
                continue
            _check(rc, err.value if rc in (4, 5) else None)
            m = int(n_out.value)
            return o_idx[:m], o_val[:m], o_bl[:m]

    def pad_counts(self, height, idx):
        counts = np.zeros(height + 1, np.uint64)
        if len(idx):
            _check(self.L.dapol_tree_level_pad_counts_dev(self.ctx._h, height, len(idx), idx.data_ptr(), _p(counts)))
        return counts

    def build_shard(self, hash_id, height, idx, values, blind, pad_seed, level_base):
        h = C.c_void_p()
        seed = (C.c_uint8 * 32).from_buffer_copy(pad_seed)
        lb = np.ascontiguousarray(level_base, np.uint64)
        _check(self.L.dapol_tree_build_shard_dev(self.ctx._h, hash_id, height, len(idx), idx.data_ptr(), values.data_ptr(), blind.data_ptr(),
                                                 seed, _p(lb), C.byref(h)))
        self.last_shard_times = self.ctx.last_build_times()
        return h

    def root_record(self, tree) -> np.ndarray:
        rec = np.zeros(RECORD_BYTES, np.uint8)
        _check(self.L.dapol_tree_root_record(tree, _p(rec)))
        return rec

    def build_top(self, hash_id, height, idx, records, pad_seed, pad_base):
        h = C.c_void_p()
        seed = (C.c_uint8 * 32).from_buffer_copy(pad_seed)
        idx = np.ascontiguousarray(idx, np.uint64)
        records = np.ascontiguousarray(records, np.uint8)
        _check(self.L.dapol_tree_build_from_records(self.ctx._h, hash_id, height, len(idx), _p(idx), _p(records), seed, pad_base, C.byref(h)))
        return h

    def attach(self, tree, top, prefix):
        _check(self.L.dapol_tree_attach_top(tree, top, prefix))

    def destroy(self, tree):
        if tree and getattr(self.ctx, "_h", None):  # trees die with their context at the latest
            self.L.dapol_tree_destroy(tree)

    def root_of(self, tree) -> DapolNode:
        com = np.zeros(32, np.uint8); hs = np.zeros(32, np.uint8); bl = np.zeros(32, np.uint8)
        v = C.c_uint64()
        _check(self.L.dapol_tree_root(tree, _p(com), _p(hs), C.byref(v), _p(bl)))
        return DapolNode(v.value, bl.tobytes(), com.tobytes(), hs.tobytes())


def shard_pad_bases(counts_all: np.ndarray, rank: int, pad_base: int = 0):
    """counts_all[r][h] = padding nodes of shard r at its level h (h = 0..Hs, level 0 = the shard's root).
    Returns (level_base[h] for `rank`, pad_base of the top tree): the single-tree build creates the padding nodes
    level by level from the leaves up and left to right inside a level, i.e. shard by shard inside a level."""
    counts_all = np.asarray(counts_all, dtype=np.uint64)
    world, L = counts_all.shape
    tot = counts_all.sum(axis=0)
    base = np.zeros(L, np.uint64)
    below = 0
    for h in range(L - 1, 0, -1):
        base[h] = pad_base + below + int(counts_all[:rank, h].sum())
        below += int(tot[h])
    return base, pad_base + below


class ShardedDapol:
    """The Dapol<D, R> surface (src/dapol/mod.rs:100-190) of one tree whose leaf ranges live on several GPUs."""

    def __init__(self, engine, comm, hash_id, height, aggregation_factor, policy):
        self.engine, self.comm = engine, comm
        self.hash_id, self.height, self.aggregation_factor, self.policy = hash_id, height, aggregation_factor, policy
        self.k = max(comm.world - 1, 0).bit_length()
        if (1 << self.k) != comm.world:
            raise DapolError(16, "world size must be a power of two")
        self.sub_height = height - self.k
        self.subtree = None
        self.top = None
        self.leaf_index_map = None
        self.n_total = 0
        self.n_mine = 0
        self.first_pos = 0
        self.n_local = 0
        self.native = False
        self.phase_ms = None

    @classmethod
    def new(cls, engine, comm, hash_id, liabilities, audit_seed: bytes, tree_height: int, aggregation_factor: int, pad_seed: bytes,
            policy=POLICY_PADDING, pad_base: int = 0, native: "NativeComm | None" = None):
        """Dapol::new(liabilities, options) (mod.rs:100-128) where `liabilities` is THIS rank's slice of the input, packed
        as (iid_blob, iid_off[n+1], eid_blob, eid_off[n+1], values[n]) (numpy, or torch tensors already on the device);
        input order = rank 0's slice, then rank 1's, ..."""
        self = cls(engine, comm, hash_id, tree_height, aggregation_factor, policy)
        if tree_height > 64:
            raise DapolError(1)
        if self.sub_height < 1:
            raise DapolError(16)
        ib, io, eb, eo, vals = liabilities
        n = len(io) - 1
        if native is not None:
            # the whole protocol behind ONE C-ABI call (dapol_sharded_build): all-to-all of claims, per-prefix collision
            # resolution, padding bases, subtree, root gather, top tree -- nothing of it runs in Python
            E, L = engine, engine.L
            d = (E.to_dev(ib, np.uint8), E.to_dev(io, np.int64), E.to_dev(eb, np.uint8), E.to_dev(eo, np.int64), E.to_dev(vals, np.int64))
            sub, top = C.c_void_p(), C.c_void_p()
            n_total, first, err = C.c_uint64(), C.c_uint64(), C.c_uint64()
            ms = np.zeros(4, np.float32)
            seed = (C.c_uint8 * 32).from_buffer_copy(pad_seed)
            aseed = (C.c_uint8 * max(len(audit_seed), 1)).from_buffer_copy(audit_seed or b"\0")
            torch.cuda.current_stream(E.device).synchronize()
            rc = L.dapol_sharded_build(E.ctx._h, native.h, hash_id, tree_height, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                       d[4].data_ptr(), aseed, len(audit_seed), seed, pad_base, C.byref(sub), C.byref(top), C.byref(n_total),
                                       C.byref(first), C.byref(err), _p(ms))
            _check(rc, err.value if rc in (4, 5) else None)
            self.subtree = sub if sub.value else None
            self.top = top if top.value else None
            self.n_total, self.first_pos, self.n_local = int(n_total.value), int(first.value), n
            self.phase_ms = dict(zip(("hash", "exchange", "subtree", "top"), ms.tolist()))
            E.last_shard_times = E.ctx.last_build_times() if self.subtree else None
            self.native = True
            return self
        counts = comm.all_gather_host(np.array([n], np.uint64), getattr(engine, "device", None))[:, 0]
        self.n_total = int(counts.sum())
        self.first_pos = int(counts[: comm.rank].sum())
        if tree_height < 64 and (1 << tree_height) < 2 * self.n_total:
            raise DapolError(2)
        E = engine
        audit, seedst, cand, blind = E.derive(hash_id, tree_height, n, E.to_dev(ib, np.uint8), E.to_dev(io, np.int64), E.to_dev(eb, np.uint8),
                                              E.to_dev(eo, np.int64), audit_seed)
        d_vals = E.to_dev(vals, np.int64)
        if comm.world > 1:  # one exchange: 112 B per user
            audit, seedst, blind = (comm.all_gather_rows(t, counts) for t in (audit, seedst, blind))
            cand, d_vals = (comm.all_gather_rows(t, counts) for t in (cand, d_vals))
        cap = min(self.n_total, self.n_total // comm.world + self.n_total // (4 * comm.world) + 1024)
        idx, v, bl = E.assign(hash_id, tree_height, self.n_total, audit, seedst, cand, blind, d_vals, self.k, comm.rank, cap)
        self.leaf_index_map = cand  # final leaf index of every input position (id_to_idx_map, mod.rs:80,389)
        self._build(idx, v, bl, pad_seed, pad_base)
        return self

    @classmethod
    def build_from_nodes(cls, engine, comm, hash_id, height, aggregation_factor, leaf_idx, values, blindings, pad_seed: bytes,
                         policy=POLICY_PADDING, pad_base: int = 0):
        """Dapol::new_blank + build (mod.rs:196-208) from THIS rank's sorted leaves: whole-tree indexes that all start with
        the rank's prefix."""
        self = cls(engine, comm, hash_id, height, aggregation_factor, policy)
        if self.sub_height < 1:
            raise DapolError(16)
        idx = np.ascontiguousarray(leaf_idx, np.uint64)
        if len(idx) and self.k and not ((idx >> np.uint64(self.sub_height)) == np.uint64(comm.rank)).all():
            raise DapolError(16, "leaf outside this rank's prefix")
        mask = np.uint64((1 << self.sub_height) - 1) if self.sub_height < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
        E = engine
        self._build(E.to_dev(idx & mask, np.int64), E.to_dev(np.ascontiguousarray(values, np.uint64), np.int64),
                    E.to_dev(np.ascontiguousarray(blindings, np.uint8).reshape(-1, 32), np.uint8), pad_seed, pad_base)
        return self

    def _build(self, idx, v, bl, pad_seed, pad_base):
        E, comm, Hs = self.engine, self.comm, self.sub_height
        self.n_mine = len(idx)
        dev = getattr(E, "device", None)
        if getattr(E, "positional", False):
            # position-keyed padding: a padding blinding depends on (level, index) only, so the shards need
            # no exchange of padding counts -- the shard passes its coordinates inside the whole tree instead
            level_base = np.array([self.k] + [comm.rank << h for h in range(1, Hs + 1)], dtype=np.uint64)
            top_base = 0
        else:
            counts_all = comm.all_gather_host(E.pad_counts(Hs, idx), dev)
            level_base, top_base = shard_pad_bases(counts_all, comm.rank, pad_base)
        rec = np.zeros(RECORD_BYTES + 8, np.uint8)
        if self.n_mine:
            self.subtree = E.build_shard(self.hash_id, Hs, idx, v, bl, pad_seed, level_base)
            rec[:RECORD_BYTES] = E.root_record(self.subtree)
            rec[RECORD_BYTES] = 1
        recs = comm.all_gather_host(rec, dev)  # the only exchange of the build proper: 2^k root records
        present = np.nonzero(recs[:, RECORD_BYTES])[0].astype(np.uint64)
        if len(present) == 0:
            raise DapolError(16, "empty tree")
        self.top = E.build_top(self.hash_id, self.k, present, recs[present.astype(np.int64), :RECORD_BYTES], pad_seed, top_base)
        if self.subtree:
            E.attach(self.subtree, self.top, comm.rank)

    # -- accessors (identical on every rank) ------------------------------------------------------
    def root_raw(self) -> DapolNode:
        return self.engine.root_of(self.top or self.subtree)  # one-rank native build: the whole tree is the `subtree`

    def root(self) -> DapolProofNode:
        return self.root_raw().get_proof_node()

    def owner_of(self, leaf_idx: int) -> int:
        return int(leaf_idx) >> self.sub_height if self.sub_height < 64 else 0

    def leaf_index_of(self, input_pos: int):
        """id_to_idx_map lookup (mod.rs:148-151) by global input position."""
        if getattr(self, "native", False):  # the native build keeps the map of the LOCAL slice only (on the top-tree handle)
            if not (self.first_pos <= input_pos < self.first_pos + self.n_local):
                return None
            x = C.c_uint64()
            _check(_ffi.lib().dapol_tree_leaf_index_of(self.top or self.subtree, input_pos, C.byref(x)))
            return x.value
        if self.leaf_index_map is None or not (0 <= input_pos < self.n_total):
            return None
        return int(self.leaf_index_map[input_pos].item()) & 0xFFFFFFFFFFFFFFFF

    # -- inclusion proofs for leaves this rank owns -------------------------------------------------
    def generate_proofs(self, leaf_idx, seed: bytes):
        """[Dapol::generate_proof(idx)] (mod.rs:167-190) for whole-tree leaf indexes owned by this rank; None otherwise."""
        li = np.ascontiguousarray(leaf_idx, np.uint64)
        if self.subtree is None or any(self.owner_of(x) != self.comm.rank for x in li.tolist()):
            return None
        L = _ffi.lib()
        size = L.dapol_inclusion_proof_size(self.height, self.aggregation_factor, self.policy)
        if size == 0:
            raise DapolError(16)
        out = np.zeros(len(li) * size, np.uint8)
        got = C.c_uint64()
        sd = (C.c_uint8 * 32).from_buffer_copy(seed)
        rc = L.dapol_prove_batch(self.subtree, len(li), _p(li), self.aggregation_factor, self.policy, sd, _p(out), out.nbytes, C.byref(got))
        if rc == 17:
            return None
        _check(rc)
        return [DapolProof(out[i * size:(i + 1) * size].tobytes(), self.hash_id, self.policy) for i in range(len(li))]

    def generate_proof(self, leaf_idx: int, seed: bytes):
        r = self.generate_proofs([leaf_idx], seed)
        return None if r is None else r[0]

    def close(self):
        if self.subtree:
            self.engine.destroy(self.subtree)
            self.subtree = None
        if self.top:
            self.engine.destroy(self.top)
            self.top = None

    __del__ = close
