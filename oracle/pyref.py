"""TEST INFRASTRUCTURE ONLY -- CPU oracle, pure-Python big-int restatement of the DAPOL+ hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
package.  Nothing under dapol_b200/ (the product) imports it.

PARITY STATUS: *parity unpinned* against the Rust reference for commitment / hash /
root / proof BYTES: the reference holds no byte-level golden vectors and cannot be
built here (no cargo, un-vendored crates).  What IS pinned, and is
checked in tests/test_oracle_pins.py:
  * the reference's own KAT  id -> leaf index  a,b,c,d -> 7,12,2,4   (src/dapol/tests.rs:38-84)
  * root value 26                                                    (src/dapol/tests.rs:24)
  * SINGLE_PROOF_BYTE_NUM = 672                                      (src/range/mod.rs:18)
  * RFC 9496 ristretto255 vectors, the bulletproofs B_blinding constant, merlin's
    published transcript test vector (third-party algorithms the reference calls).

Each function cites the reference file:line (in /root/reference) or the upstream crate
whose published algorithm it restates:
  curve25519-dalek-ng ^4.1.1, bulletproofs ^4.0.0, merlin ^3.0.0, smtree ^0.1.2,
  blake3 ^0.3.8, blake2 ^0.9, rand_chacha (ChaCha20Rng) for the injected-RNG contract.
"""
from __future__ import annotations

import hashlib
import struct

try:  # D = blake3 (benches/dapol.rs:38) -- python package used only as a pin for oracle/c
    import blake3 as _blake3
except Exception:  # pragma: no cover
    _blake3 = None

# --------------------------------------------------------------------------------------
# GF(2^255-19)  (dalek field.rs)
# --------------------------------------------------------------------------------------
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


def inv(x):
    return pow(x, P - 2, P)


D = (-121665 * inv(121666)) % P
D2 = (2 * D) % P
SQRT_M1 = 19681161376707505956807079304988542015446066515923890162744021073123829784752
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
INVSQRT_A_MINUS_D = 54469307008909316920995813868745141605393597292927456921205312896311721017578
ONE_MINUS_D_SQ = 1159843021668779879193775521855586647937357759715417654439879720876111806838
D_MINUS_ONE_SQ = 40440834346308536858101042469323190826248399146238708352240133220865137265952


def is_neg(x):
    return (x % P) & 1


def fabs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u, v):
    """RFC 9496 4.2 / dalek FieldElement::sqrt_ratio_i."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = (u * v3) % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    return (correct or flipped), fabs(r)


# --------------------------------------------------------------------------------------
# ristretto255 (dalek ristretto.rs == RFC 9496).  Points are extended (X, Y, Z, T).
# --------------------------------------------------------------------------------------
IDENT = (0, 1, 1, 0)
_BX = 15112221349535400772501151409588531511454012693041857206046113283949847762202
_BY = 46316835694926478169428394003475163141307993866256225615783033603165251855960
BASEPOINT = (_BX, _BY, 1, _BX * _BY % P)


def pt_add(p, q):
    X1, Y1, Z1, T1 = p
    X2, Y2, Z2, T2 = q
    A = (Y1 - X1) * (Y2 - X2) % P
    B = (Y1 + X1) * (Y2 + X2) % P
    C = T1 * D2 % P * T2 % P
    Dd = 2 * Z1 * Z2 % P
    E, F, G, H = B - A, Dd - C, Dd + C, B + A
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def pt_neg(p):
    return ((-p[0]) % P, p[1], p[2], (-p[3]) % P)


def pt_sub(p, q):
    return pt_add(p, pt_neg(q))


def pt_mul(k, p):
    k %= L
    acc = IDENT
    while k:
        if k & 1:
            acc = pt_add(acc, p)
        p = pt_add(p, p)
        k >>= 1
    return acc


def pt_eq(p, q):
    """RistrettoPoint::ct_eq."""
    return (p[0] * q[1] - p[1] * q[0]) % P == 0 or (p[0] * q[0] - p[1] * q[1]) % P == 0


def pt_is_identity(p):
    return pt_eq(p, IDENT)


def compress(p) -> bytes:
    """RistrettoPoint::compress (RFC 9496 4.3.2)."""
    x0, y0, z0, t0 = p
    u1 = (z0 + y0) * (z0 - y0) % P
    u2 = x0 * y0 % P
    _, invsqrt = sqrt_ratio_m1(1, u1 * u2 % P * u2 % P)
    den1 = invsqrt * u1 % P
    den2 = invsqrt * u2 % P
    z_inv = den1 * den2 % P * t0 % P
    ix0 = x0 * SQRT_M1 % P
    iy0 = y0 * SQRT_M1 % P
    ench = den1 * INVSQRT_A_MINUS_D % P
    rotate = is_neg(t0 * z_inv)
    if rotate:
        x, y, den_inv = iy0, ix0, ench
    else:
        x, y, den_inv = x0, y0, den2
    if is_neg(x * z_inv):
        y = (-y) % P
    s = fabs(den_inv * (z0 - y))
    return s.to_bytes(32, "little")


def decompress(b: bytes):
    """CompressedRistretto::decompress (RFC 9496 4.3.1); None on failure."""
    s = int.from_bytes(b, "little")
    if s >= P or (s & 1):
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2s = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2s) % P
    ok, invsqrt = sqrt_ratio_m1(1, v * u2s % P)
    den_x = invsqrt * u2 % P
    den_y = invsqrt * den_x % P * v % P
    x = fabs(2 * s * den_x)
    y = u1 * den_y % P
    t = x * y % P
    if (not ok) or is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def elligator(t):
    """RistrettoPoint::elligator_ristretto_flavor (RFC 9496 4.3.4 MAP)."""
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) % P * ((r + D) % P) % P
    ok, s = sqrt_ratio_m1(u, v)
    s_prime = (-fabs(s * t)) % P
    if not ok:
        s = s_prime
    c = (P - 1) if ok else r
    N = (c * (r - 1) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform_bytes(b: bytes):
    """RistrettoPoint::from_uniform_bytes: two field elements, bit 255 masked."""
    assert len(b) == 64
    t1 = (int.from_bytes(b[:32], "little") & ((1 << 255) - 1)) % P
    t2 = (int.from_bytes(b[32:], "little") & ((1 << 255) - 1)) % P
    return pt_add(elligator(t1), elligator(t2))


# --------------------------------------------------------------------------------------
# Scalars mod l (dalek scalar.rs)
# --------------------------------------------------------------------------------------
def sc_from_bits(b: bytes) -> int:
    """Scalar::from_bits: clears bit 255 only, NOT reduced (src/dapol/mod.rs:385)."""
    return int.from_bytes(b, "little") & ((1 << 255) - 1)


def sc_wide(b: bytes) -> int:
    """Scalar::from_bytes_mod_order_wide."""
    return int.from_bytes(b, "little") % L


def sc_bytes(x: int) -> bytes:
    return (x % L).to_bytes(32, "little")


def sc_canonical(b: bytes):
    """Scalar::from_canonical_bytes; None if >= l."""
    x = int.from_bytes(b, "little")
    return x if x < L else None


def sc_inv(x):
    return pow(x, L - 2, L)


# --------------------------------------------------------------------------------------
# Seeded RNG contract (replaces thread_rng): ChaCha20Rng::from_seed(seed),
# 64-bit block counter (words 12,13), 64-bit stream id (words 14,15), LE output.
# The k-th Scalar::random draw of a stream = wide-reduce of keystream block k.
# --------------------------------------------------------------------------------------
def _rotl32(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def chacha20_block(key: bytes, counter: int, stream: int = 0) -> bytes:
    st = list(struct.unpack("<4I", b"expand 32-byte k")) + list(struct.unpack("<8I", key))
    st += [counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, stream & 0xFFFFFFFF, (stream >> 32) & 0xFFFFFFFF]
    w = st[:]

    def qr(a, b, c, d):
        w[a] = (w[a] + w[b]) & 0xFFFFFFFF; w[d] = _rotl32(w[d] ^ w[a], 16)
        w[c] = (w[c] + w[d]) & 0xFFFFFFFF; w[b] = _rotl32(w[b] ^ w[c], 12)
        w[a] = (w[a] + w[b]) & 0xFFFFFFFF; w[d] = _rotl32(w[d] ^ w[a], 8)
        w[c] = (w[c] + w[d]) & 0xFFFFFFFF; w[b] = _rotl32(w[b] ^ w[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return struct.pack("<16I", *[(w[i] + st[i]) & 0xFFFFFFFF for i in range(16)])


class ScalarRng:
    """Stream of Scalar::random draws: draw k = block (base + k) of ChaCha20(seed, stream)."""

    def __init__(self, seed: bytes, stream: int = 0, base: int = 0):
        assert len(seed) == 32
        self.seed, self.stream, self.k = seed, stream, base

    def scalar(self) -> int:
        b = chacha20_block(self.seed, self.k, self.stream)
        self.k += 1
        return sc_wide(b)


def rng_scalar(seed: bytes, k: int, stream: int = 0) -> int:
    return sc_wide(chacha20_block(seed, k, stream))


# --------------------------------------------------------------------------------------
# Digests D (digest::Digest users in the reference)
# --------------------------------------------------------------------------------------
HASH_BLAKE3 = 0
HASH_BLAKE2S = 1
HASH_BLAKE2B = 2


def digest(hash_id: int, *parts: bytes) -> bytes:
    data = b"".join(parts)
    if hash_id == HASH_BLAKE3:
        return _blake3.blake3(data).digest()
    if hash_id == HASH_BLAKE2S:
        return hashlib.blake2s(data).digest()
    if hash_id == HASH_BLAKE2B:  # blake2::Blake2b, 64-byte digests (src/tests.rs:104-105; new_blank + build only)
        return hashlib.blake2b(data).digest()
    raise ValueError("hash id")


def dlen(hash_id: int) -> int:
    return 64 if hash_id == HASH_BLAKE2B else 32


# --------------------------------------------------------------------------------------
# Keccak-f[1600], STROBE-128, merlin transcript (merlin strobe.rs / transcript.rs)
# --------------------------------------------------------------------------------------
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def keccak_f(state: bytearray):
    a = [[int.from_bytes(state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]

    def rol(v, n):
        n %= 64
        return ((v << n) | (v >> (64 - n))) & _M64 if n else v

    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")


class Strobe128:
    R = 166
    I, A, C, T, M, K = 1, 2, 4, 8, 16, 32

    def __init__(self, label: bytes):
        self.st = bytearray(200)
        self.st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        self.st[6:18] = b"STROBEv1.0.2"
        keccak_f(self.st)
        self.pos = self.pos_begin = self.cur_flags = 0
        self.meta_ad(label, False)

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[self.R + 1] ^= 0x80
        keccak_f(self.st)
        self.pos = self.pos_begin = 0

    def _absorb(self, data):
        for b in data:
            self.st[self.pos] ^= b
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray()
        for _ in range(n):
            out.append(self.st[self.pos])
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        old = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old, flags]))
        if flags & (self.C | self.K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.M | self.A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.A, more)
        self._absorb(data)

    def prf(self, n, more=False):
        self._begin_op(self.I | self.A | self.C, more)
        return self._squeeze(n)


class Transcript:
    """merlin::Transcript + bulletproofs TranscriptProtocol (bulletproofs transcript.rs)."""

    def __init__(self, label: bytes = b""):
        # reference uses Transcript::new(&[]) -- the EMPTY label (src/range/mod.rs:51,67,86,105)
        self.s = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, msg):
        self.s.meta_ad(label, False)
        self.s.meta_ad(struct.pack("<I", len(msg)), True)
        self.s.ad(msg, False)

    def append_u64(self, label, x):
        self.append_message(label, struct.pack("<Q", x))

    def challenge_bytes(self, label, n):
        self.s.meta_ad(label, False)
        self.s.meta_ad(struct.pack("<I", n), True)
        return self.s.prf(n)

    def challenge_scalar(self, label):
        return sc_wide(self.challenge_bytes(label, 64))

    def rangeproof_domain_sep(self, n, m):
        self.append_message(b"dom-sep", b"rangeproof v1")
        self.append_u64(b"n", n)
        self.append_u64(b"m", m)

    def innerproduct_domain_sep(self, n):
        self.append_message(b"dom-sep", b"ipp v1")
        self.append_u64(b"n", n)

    def append_scalar(self, label, x):
        self.append_message(label, sc_bytes(x))

    def append_point(self, label, pbytes):
        self.append_message(label, pbytes)

    def validate_and_append_point(self, label, pbytes) -> bool:
        if pbytes == bytes(32):
            return False
        self.append_message(label, pbytes)
        return True


# --------------------------------------------------------------------------------------
# Generators (bulletproofs generators.rs)
# --------------------------------------------------------------------------------------
B_BLINDING = from_uniform_bytes(hashlib.sha3_512(compress(BASEPOINT)).digest())


def pedersen_commit(v: int, r: int):
    """PedersenGens::default().commit  (src/dapol/node.rs:31)."""
    return pt_add(pt_mul(v, BASEPOINT), pt_mul(r, B_BLINDING))


_GEN_CACHE: dict = {}


def bp_gens(n: int, m: int):
    """BulletproofGens::new(n, m): (G[j][i], H[j][i]) via SHAKE256 chains."""
    out = {}
    for tag in (b"G", b"H"):
        rows = []
        for j in range(m):
            key = (tag, j)
            have = _GEN_CACHE.get(key, [])
            if len(have) < n:
                stream = hashlib.shake_256(b"GeneratorsChain" + tag + struct.pack("<I", j)).digest(64 * n)
                have = have + [from_uniform_bytes(stream[64 * i: 64 * i + 64]) for i in range(len(have), n)]
                _GEN_CACHE[key] = have
            rows.append(have[:n])
        out[tag] = rows
    return out[b"G"], out[b"H"]


# --------------------------------------------------------------------------------------
# Node algebra (src/dapol/node.rs)
# --------------------------------------------------------------------------------------
class Node:
    __slots__ = ("v", "r", "com", "comc", "hash")

    def __init__(self, v, r, com, comc, h):
        self.v, self.r, self.com, self.comc, self.hash = v, r, com, comc, h


def node_new(hash_id, v: int, r: int) -> Node:
    """DapolNode::new (node.rs:29-45): com = v*B + r*B_blinding; hash = D(compress(com))."""
    com = pedersen_commit(v, r)
    cc = compress(com)
    return Node(v, r, com, cc, digest(hash_id, cc))


def node_merge(hash_id, l: Node, r: Node) -> Node:
    """Mergeable::merge (node.rs:64-80): hash = D(C(L)||C(R)||H(L)||H(R)); sums."""
    h = digest(hash_id, l.comc, r.comc, l.hash, r.hash)
    com = pt_add(l.com, r.com)
    return Node((l.v + r.v) & _M64, (l.r + r.r) % L, com, compress(com), h)


def proofnode_merge(hash_id, l, r):
    """DapolProofNode::merge (src/proof/node.rs:56-69) on (point, comc, hash) triples."""
    com = pt_add(l[0], r[0])
    return (com, compress(com), digest(hash_id, l[1], r[1], l[2], r[2]))


# --------------------------------------------------------------------------------------
# Leaf derivation (src/dapol/mod.rs:323-441)
# --------------------------------------------------------------------------------------
MAX_TREE_HEIGHT, MIN_SPARSITY, MAX_INDEX_RETRIES = 64, 2, 128

ERR_OK, ERR_TREE_HEIGHT_TOO_BIG, ERR_SPARSITY_TOO_SMALL, ERR_INVALID_DIGEST_SIZE = 0, 1, 2, 3
ERR_DUPLICATED_INTERNAL_ID, ERR_FAILED_TO_MAP_INDEX = 4, 5


class DapolError(Exception):
    def __init__(self, code, detail=None):
        super().__init__(f"DapolError({code}, {detail})")
        self.code, self.detail = code, detail


def derive_leaves(hash_id, liabilities, audit_seed: bytes, height: int):
    """build_leaf_nodes + shuffle_index: returns [(idx, value, blinding_from_bits)] in INPUT order.

    liabilities: [(internal_id bytes, external_id bytes, value)].
    """
    seen_ids, used = set(), set()
    out = []
    for pos, (iid, eid, value) in enumerate(liabilities):
        if iid in seen_ids:
            raise DapolError(ERR_DUPLICATED_INTERNAL_ID, pos)
        audit_id = digest(hash_id, audit_seed, iid)
        seed = digest(hash_id, audit_id, b"index_seed", eid)
        idx = None
        for _ in range(MAX_INDEX_RETRIES):
            seed = digest(hash_id, seed)
            cand = int.from_bytes(seed[:8], "big") >> (64 - height)
            if cand not in used:
                used.add(cand)
                idx = cand
                break
        if idx is None:
            raise DapolError(ERR_FAILED_TO_MAP_INDEX, pos)
        blind_seed = digest(hash_id, audit_id, b"blind_seed", eid)
        seen_ids.add(iid)
        out.append((idx, value, sc_from_bits(blind_seed)))
    return out


def check_options(n_liab: int, height: int):
    """Dapol::new argument checks (mod.rs:101-116)."""
    if height > MAX_TREE_HEIGHT:
        raise DapolError(ERR_TREE_HEIGHT_TOO_BIG, height)
    if (1 << height) < n_liab * MIN_SPARSITY:
        raise DapolError(ERR_SPARSITY_TOO_SMALL, height)


# --------------------------------------------------------------------------------------
# Padded sparse Merkle tree build (smtree SparseMerkleTree::build):
# bottom-up, one level at a time, left to right; lone node -> padding sibling.
# Padding draw order under the RNG contract = creation order (level H..1, left to right).
# --------------------------------------------------------------------------------------
class Tree:
    def __init__(self, hash_id, height):
        self.hash_id, self.height = hash_id, height
        self.levels = [dict() for _ in range(height + 1)]  # levels[h][idx] = Node ; h=0 root
        self.is_pad = [set() for _ in range(height + 1)]

    @property
    def root(self) -> Node:
        return self.levels[0][0]

    def path_siblings(self, leaf_idx: int):
        """Siblings leaf-level first, root-child last (single-leaf get_merkle_path_ref_batch)."""
        H = self.height
        return [self.levels[h][(leaf_idx >> (H - h)) ^ 1] for h in range(H, 0, -1)]


def build_tree(hash_id, height, leaves, pad_seed: bytes, pad_draw_base: int = 0, positional: bool = False) -> Tree:
    """leaves: iterable of (idx, Node) at leaf level.  Dapol::build (mod.rs:206-208).
    positional (opt-in): the padding node at (level h, index i) draws block i of stream h instead of the next
    block of the creation-order stream -- Paddable::padding(idx, secret) as a function of its arguments (node.rs:85-88 TODO)."""
    t = Tree(hash_id, height)
    cur = dict(sorted((i, n) for i, n in leaves))
    draw = pad_draw_base
    for h in range(height, 0, -1):
        lvl = t.levels[h]
        parents = {}
        for idx in sorted(cur):
            lvl[idx] = cur[idx]
        for idx in sorted(cur):
            sib = idx ^ 1
            if sib not in cur:
                # DapolNode::padding (node.rs:86-88): new(0, Scalar::random(rng))
                lvl[sib] = node_new(hash_id, 0, rng_scalar(pad_seed, sib, h) if positional else rng_scalar(pad_seed, draw))
                t.is_pad[h].add(sib)
                draw += 1
            elif idx & 1:
                continue  # pair handled at its left member
            l, r = (lvl[idx], lvl[sib]) if idx & 1 == 0 else (lvl[sib], lvl[idx])
            parents[idx >> 1] = node_merge(hash_id, l, r)
        cur = parents
    t.levels[0] = cur if height > 0 else dict(cur)
    return t


# --------------------------------------------------------------------------------------
# Bulletproofs range proof (bulletproofs range_proof/{mod,party,dealer}.rs,
# inner_product_proof.rs).  n = bit size, m = parties (power of two).
# --------------------------------------------------------------------------------------
def _msm(scalars, points):
    acc = IDENT
    for s, p in zip(scalars, points):
        if s % L:
            acc = pt_add(acc, pt_mul(s, p))
    return acc


def rp_prove(values, blindings, rng: ScalarRng, n: int = 64) -> bytes:
    """RangeProof::prove_multiple_with_rng with Transcript::new(&[]) (src/range/mod.rs:48-78)."""
    m = len(values)
    assert m and m & (m - 1) == 0 and n in (8, 16, 32, 64)
    G, Hh = bp_gens(n, m)
    tr = Transcript(b"")
    tr.rangeproof_domain_sep(n, m)
    a_bl, s_bl, sL, sR, V, A, S = [], [], [], [], [], IDENT, IDENT
    for j in range(m):  # Party::assign_position_with_rng
        ab = rng.scalar()
        Aj = pt_mul(ab, B_BLINDING)
        for i in range(n):
            Aj = pt_add(Aj, G[j][i]) if (values[j] >> i) & 1 else pt_sub(Aj, Hh[j][i])
        sb = rng.scalar()
        sl = [rng.scalar() for _ in range(n)]
        sr = [rng.scalar() for _ in range(n)]
        Sj = pt_add(pt_mul(sb, B_BLINDING), pt_add(_msm(sl, G[j]), _msm(sr, Hh[j])))
        a_bl.append(ab); s_bl.append(sb); sL.append(sl); sR.append(sr)
        V.append(compress(pedersen_commit(values[j], blindings[j])))
        A, S = pt_add(A, Aj), pt_add(S, Sj)
    for vj in V:
        tr.append_point(b"V", vj)
    Ab, Sb = compress(A), compress(S)
    tr.append_point(b"A", Ab)
    tr.append_point(b"S", Sb)
    y = tr.challenge_scalar(b"y")
    z = tr.challenge_scalar(b"z")
    # Party::apply_challenge_with_rng
    l0, l1, r0, r1, t1b, t2b = [], [], [], [], [], []
    T1, T2 = IDENT, IDENT
    t0s = t1s = t2s = 0
    for j in range(m):
        offset_zz = z * z % L * pow(z, j, L) % L
        exp_y = pow(y, j * n, L)
        exp_2 = 1
        pl0, pl1, pr0, pr1 = [], [], [], []
        for i in range(n):
            aL = (values[j] >> i) & 1
            aR = (aL - 1) % L
            pl0.append((aL - z) % L)
            pl1.append(sL[j][i])
            pr0.append((exp_y * ((aR + z) % L) + offset_zz * exp_2) % L)
            pr1.append(exp_y * sR[j][i] % L)
            exp_y = exp_y * y % L
            exp_2 = exp_2 * 2 % L
        t0 = sum(a * b for a, b in zip(pl0, pr0)) % L
        t2 = sum(a * b for a, b in zip(pl1, pr1)) % L
        t1 = (sum((a + b) * (c + d) for a, b, c, d in zip(pl0, pl1, pr0, pr1)) - t0 - t2) % L
        b1 = rng.scalar()
        b2 = rng.scalar()
        T1 = pt_add(T1, pedersen_commit(t1, b1))
        T2 = pt_add(T2, pedersen_commit(t2, b2))
        l0.append(pl0); l1.append(pl1); r0.append(pr0); r1.append(pr1)
        t1b.append(b1); t2b.append(b2)
        t0s, t1s, t2s = t0s + t0, t1s + t1, t2s + t2
    T1b, T2b = compress(T1), compress(T2)
    tr.append_point(b"T_1", T1b)
    tr.append_point(b"T_2", T2b)
    x = tr.challenge_scalar(b"x")
    assert x != 0
    t_x = (t0s + t1s * x + t2s * x * x) % L
    t_x_bl = e_bl = 0
    lvec, rvec = [], []
    for j in range(m):
        offset_zz = z * z % L * pow(z, j, L) % L
        t_x_bl += offset_zz * blindings[j] + x * (t1b[j] + x * t2b[j])
        e_bl += a_bl[j] + s_bl[j] * x
        lvec += [(a + b * x) % L for a, b in zip(l0[j], l1[j])]
        rvec += [(a + b * x) % L for a, b in zip(r0[j], r1[j])]
    t_x_bl %= L
    e_bl %= L
    tr.append_scalar(b"t_x", t_x)
    tr.append_scalar(b"t_x_blinding", t_x_bl)
    tr.append_scalar(b"e_blinding", e_bl)
    w = tr.challenge_scalar(b"w")
    Q = pt_mul(w, BASEPOINT)
    # InnerProductProof::create with G_factors = 1, H_factors = y^-i
    N = n * m
    Gv = [g for row in G for g in row]
    y_inv = sc_inv(y)
    Hv = [pt_mul(pow(y_inv, i, L), h) for i, h in enumerate(h for row in Hh for h in row)]
    tr.innerproduct_domain_sep(N)
    a, b = lvec, rvec
    LR = []
    while N > 1:
        N //= 2
        aL, aR, bL, bR = a[:N], a[N:], b[:N], b[N:]
        GL, GR, HL, HR = Gv[:N], Gv[N:], Hv[:N], Hv[N:]
        cL = sum(p * q for p, q in zip(aL, bR)) % L
        cR = sum(p * q for p, q in zip(aR, bL)) % L
        Lp = pt_add(pt_add(_msm(aL, GR), _msm(bR, HL)), pt_mul(cL, Q))
        Rp = pt_add(pt_add(_msm(aR, GL), _msm(bL, HR)), pt_mul(cR, Q))
        Lb, Rb = compress(Lp), compress(Rp)
        LR += [Lb, Rb]
        tr.append_point(b"L", Lb)
        tr.append_point(b"R", Rb)
        u = tr.challenge_scalar(b"u")
        ui = sc_inv(u)
        a = [(p * u + q * ui) % L for p, q in zip(aL, aR)]
        b = [(p * ui + q * u) % L for p, q in zip(bL, bR)]
        Gv = [pt_add(pt_mul(ui, p), pt_mul(u, q)) for p, q in zip(GL, GR)]
        Hv = [pt_add(pt_mul(u, p), pt_mul(ui, q)) for p, q in zip(HL, HR)]
    return b"".join([Ab, Sb, T1b, T2b, sc_bytes(t_x), sc_bytes(t_x_bl), sc_bytes(e_bl)] + LR + [sc_bytes(a[0]), sc_bytes(b[0])])


def rp_parse(proof: bytes):
    """RangeProof::from_bytes + InnerProductProof::from_bytes; None = FormatError."""
    if len(proof) % 32 or len(proof) < 7 * 32:
        return None
    w = [proof[i: i + 32] for i in range(0, len(proof), 32)]
    sc = [sc_canonical(x) for x in w[4:7]]
    ipp = w[7:]
    if len(ipp) < 2 or (len(ipp) - 2) % 2:
        return None
    lg = (len(ipp) - 2) // 2
    if lg >= 32:
        return None
    a, b = sc_canonical(ipp[-2]), sc_canonical(ipp[-1])
    if None in sc or a is None or b is None:
        return None
    return dict(A=w[0], S=w[1], T1=w[2], T2=w[3], t_x=sc[0], t_x_bl=sc[1], e_bl=sc[2],
                L=ipp[0:2 * lg:2], R=ipp[1:2 * lg:2], a=a, b=b, lg=lg)


def rp_verify(proof: bytes, commitments, n: int = 64, c: int | None = None) -> bool:
    """RangeProof::verify_multiple with Transcript::new(&[]) (src/range/mod.rs:83-119)."""
    m = len(commitments)
    if n not in (8, 16, 32, 64) or m == 0 or m & (m - 1):
        return False
    pr = rp_parse(proof)
    if pr is None:
        return False
    tr = Transcript(b"")
    tr.rangeproof_domain_sep(n, m)
    for V in commitments:
        tr.append_point(b"V", V)
    if not tr.validate_and_append_point(b"A", pr["A"]) or not tr.validate_and_append_point(b"S", pr["S"]):
        return False
    y = tr.challenge_scalar(b"y")
    z = tr.challenge_scalar(b"z")
    if not tr.validate_and_append_point(b"T_1", pr["T1"]) or not tr.validate_and_append_point(b"T_2", pr["T2"]):
        return False
    x = tr.challenge_scalar(b"x")
    tr.append_scalar(b"t_x", pr["t_x"])
    tr.append_scalar(b"t_x_blinding", pr["t_x_bl"])
    tr.append_scalar(b"e_blinding", pr["e_bl"])
    w = tr.challenge_scalar(b"w")
    N, lg = n * m, pr["lg"]
    if (1 << lg) != N:
        return False
    tr.innerproduct_domain_sep(N)
    us = []
    for Lb, Rb in zip(pr["L"], pr["R"]):
        if not tr.validate_and_append_point(b"L", Lb) or not tr.validate_and_append_point(b"R", Rb):
            return False
        us.append(tr.challenge_scalar(b"u"))
    if c is None:  # batching weight; any value (thread_rng in the reference)
        c = sc_wide(hashlib.sha512(proof).digest()) or 1
    pts = [decompress(b) for b in [pr["A"], pr["S"], pr["T1"], pr["T2"]] + list(pr["L"]) + list(pr["R"]) + list(commitments)]
    if any(p is None for p in pts):
        return False
    A, S, T1, T2 = pts[:4]
    Ls, Rs, Vs = pts[4:4 + lg], pts[4 + lg:4 + 2 * lg], pts[4 + 2 * lg:]
    u_sq = [u * u % L for u in us]
    u_inv = [sc_inv(u) for u in us]
    u_inv_sq = [v * v % L for v in u_inv]
    s = [0] * N
    s[0] = 1
    for v in u_inv:
        s[0] = s[0] * v % L
    for i in range(1, N):
        lg_i = i.bit_length() - 1
        s[i] = s[i - (1 << lg_i)] * u_sq[lg - 1 - lg_i] % L
    a, b = pr["a"], pr["b"]
    zz = z * z % L
    y_inv = sc_inv(y)
    sum_y = sum(pow(y, i, L) for i in range(N)) % L
    sum_z = sum(pow(z, j, L) for j in range(m)) % L
    delta = ((z - zz) * sum_y - z * zz % L * ((1 << n) - 1) % L * sum_z) % L
    G, Hh = bp_gens(n, m)
    acc = pt_add(A, pt_mul(x, S))
    acc = pt_add(acc, pt_mul(c * x, T1))
    acc = pt_add(acc, pt_mul(c * x * x, T2))
    for k in range(lg):
        acc = pt_add(acc, pt_mul(u_sq[k], Ls[k]))
        acc = pt_add(acc, pt_mul(u_inv_sq[k], Rs[k]))
    acc = pt_add(acc, pt_mul(-pr["e_bl"] - c * pr["t_x_bl"], B_BLINDING))
    acc = pt_add(acc, pt_mul(w * (pr["t_x"] - a * b) + c * (delta - pr["t_x"]), BASEPOINT))
    for i in range(N):
        j, ii = divmod(i, n)
        gs = (-z - a * s[i]) % L
        hs = (z + pow(y_inv, i, L) * (zz * pow(z, j, L) % L * (1 << ii) - b * s[N - 1 - i])) % L
        acc = pt_add(acc, pt_mul(gs, G[j][ii]))
        acc = pt_add(acc, pt_mul(hs, Hh[j][ii]))
    for j in range(m):
        acc = pt_add(acc, pt_mul(c * zz % L * pow(z, j, L), Vs[j]))
    return pt_is_identity(acc)


# --------------------------------------------------------------------------------------
# Aggregation policies + wire formats (src/range/padding.rs, splitting.rs, mod.rs:16-21)
# --------------------------------------------------------------------------------------
POLICY_PADDING, POLICY_SPLITTING = 0, 1
SINGLE_PROOF_BYTE_NUM = 672


def next_pow2(x):
    return 1 if x <= 1 else 1 << (x - 1).bit_length()


def policy_plan(n_siblings: int, agg: int, policy: int):
    """Returns [(start, count, padded_m)] aggregated groups then individual singles.

    Padding (padding.rs:88-118): ONE aggregated proof over the first `agg` siblings padded
    to next_power_of_two with (0, Scalar::one()); Splitting (splitting.rs:100-129): one
    aggregated proof per set bit of `agg`, largest first.  Rest: singles.
    The reference panics (slice OOB) if agg > n_siblings; callers must reject that.
    """
    assert agg <= n_siblings
    groups = []
    if policy == POLICY_PADDING:
        groups.append((0, agg, next_pow2(agg)))  # agg == 0 -> m = 1 of pure padding
        pos = agg
    else:
        base, pos = next_pow2(agg), 0
        while pos < agg:
            if agg & base:
                groups.append((pos, base, base))
                pos += base
            base >>= 1
    singles = list(range(pos, n_siblings))
    return groups, singles


def policy_prove(values, blindings, agg, policy, seed: bytes, stream: int):
    """R::generate_proof.  RNG contract: proof #q of this DapolProof uses ChaCha20(seed,
    stream) starting at block q << 32 (aggregated proofs first, then singles)."""
    groups, singles = policy_plan(len(values), agg, policy)
    aggregated, individual, q = [], [], 0
    for start, count, m in groups:
        vs = list(values[start:start + count]) + [0] * (m - count)
        bs = list(blindings[start:start + count]) + [1] * (m - count)
        aggregated.append(rp_prove(vs, bs, ScalarRng(seed, stream, q << 32)))
        q += 1
    for pos in singles:
        individual.append(rp_prove([values[pos]], [blindings[pos]], ScalarRng(seed, stream, q << 32)))
        q += 1
    return aggregated, individual


def policy_serialize(aggregated, individual, policy) -> bytes:
    """Serializable::serialize (padding.rs:40-53 / splitting.rs:38-59); usize_to_bytes = big-endian."""
    out = b""
    if policy == POLICY_PADDING:
        assert len(aggregated) == 1
    else:
        out += len(aggregated).to_bytes(2, "big")
    for p in aggregated:
        out += len(p).to_bytes(8, "big") + p
    out += len(individual).to_bytes(8, "big")
    for p in individual:
        out += p
    return out


def policy_deserialize(b: bytes, policy, begin=0):
    """deserialize_as_a_unit; returns (aggregated, individual, new_begin) or None."""
    def take(k):
        nonlocal begin
        if len(b) - begin < k:
            raise ValueError
        r = b[begin:begin + k]
        begin += k
        return r
    try:
        n_agg = 1 if policy == POLICY_PADDING else int.from_bytes(take(2), "big")
        aggregated = []
        for _ in range(n_agg):
            size = int.from_bytes(take(8), "big")
            p = take(size)
            if rp_parse(p) is None:
                return None
            aggregated.append(p)
        k = int.from_bytes(take(8), "big")
        individual = []
        for _ in range(k):
            p = take(SINGLE_PROOF_BYTE_NUM)
            if rp_parse(p) is None:
                return None
            individual.append(p)
    except ValueError:
        return None
    return aggregated, individual, begin


def policy_verify(aggregated, individual, commitments, policy) -> bool:
    """RangeVerifiable::verify (padding.rs:168-197 / splitting.rs:180-211)."""
    if len(individual) > len(commitments):
        return False  # reference: usize underflow panic
    n_agg = len(commitments) - len(individual)
    if policy == POLICY_PADDING:
        if len(aggregated) != 1:
            return False
        m = next_pow2(n_agg)
        coms = list(commitments[:n_agg]) + [compress(B_BLINDING)] * (m - n_agg)
        if not rp_verify(aggregated[0], coms):
            return False
    else:
        base, pos, idx = next_pow2(n_agg), 0, 0
        while pos < n_agg:
            if n_agg & base:
                if idx >= len(aggregated) or not rp_verify(aggregated[idx], commitments[pos:pos + base]):
                    return False
                idx += 1
                pos += base
            base >>= 1
    for k, p in enumerate(individual):
        if not rp_verify(p, [commitments[n_agg + k]]):
            return False
    return True


# --------------------------------------------------------------------------------------
# Inclusion proofs (src/dapol/mod.rs:167-190, src/proof/mod.rs:41-95)
# Merkle wire format (smtree, UPSTREAM-RECALL, SURVEY App. A.6 -- widths unverified):
#   be16(height) || be64(batch_num) || path bytes (ceil(h/8), MSB-first) each
#   || be64(sibling_num) || siblings (com 32 || hash 32)
# --------------------------------------------------------------------------------------
def merkle_serialize(height, leaf_idx, siblings) -> bytes:
    out = height.to_bytes(2, "big") + (1).to_bytes(8, "big")
    nbytes = (height + 7) // 8
    out += ((leaf_idx << (8 * nbytes - height)) if height else 0).to_bytes(nbytes, "big")
    out += len(siblings).to_bytes(8, "big")
    for comc, h in siblings:
        out += comc + h
    return out


def merkle_deserialize(b: bytes, begin=0, dlen=32):
    try:
        height = int.from_bytes(b[begin:begin + 2], "big"); begin += 2
        batch = int.from_bytes(b[begin:begin + 8], "big"); begin += 8
        if batch != 1 or height > 64:
            return None
        nbytes = (height + 7) // 8
        if len(b) - begin < nbytes + 8:
            return None
        idx = int.from_bytes(b[begin:begin + nbytes], "big") >> (8 * nbytes - height) if height else 0
        begin += nbytes
        k = int.from_bytes(b[begin:begin + 8], "big"); begin += 8
        if len(b) - begin < k * (32 + dlen):
            return None
        sibs = []
        for _ in range(k):
            comc, h = b[begin:begin + 32], b[begin + 32:begin + 32 + dlen]
            if decompress(comc) is None:  # proof/node.rs:81-102 rejects non-canonical points
                return None
            sibs.append((comc, h))
            begin += 32 + dlen
    except Exception:
        return None
    return height, idx, sibs, begin


PROVER_KEY_LABEL = b"dapol-b200 prover nonce key v1"


def prover_nonce_key(seed: bytes, root_com: bytes, root_hash: bytes, policy: int, agg: int, height: int) -> bytes:
    """ChaCha20 key of the prover's nonce streams of one tree (RNG contract, include/dapol_b200.h): the caller's seed bound to
    the tree (its root commits to every witness), the policy, the aggregation factor and the height, so that re-using a seed
    for another tree / policy / factor never re-uses a nonce with a different witness.  Always BLAKE3, whatever D is."""
    return digest(HASH_BLAKE3, PROVER_KEY_LABEL, seed, root_com, root_hash, struct.pack("<QQQ", policy, agg, height))


def prove_inclusion(tree: Tree, leaf_idx: int, agg: int, policy: int, seed: bytes) -> bytes:
    """Dapol::generate_proof -> DapolProof::serialize = range || merkle (proof/mod.rs:68-73)."""
    sibs = tree.path_siblings(leaf_idx)
    values = [s.v for s in sibs]
    blindings = [s.r % L for s in sibs]
    root = tree.root  # property
    key = prover_nonce_key(seed, root.comc, root.hash, policy, agg, tree.height)
    aggregated, individual = policy_prove(values, blindings, agg, policy, key, leaf_idx)
    return policy_serialize(aggregated, individual, policy) + merkle_serialize(
        tree.height, leaf_idx, [(s.comc, s.hash) for s in sibs])


def verify_inclusion(hash_id, proof: bytes, policy: int, root, leaf) -> bool:
    """DapolProof::deserialize + verify(root, leaf); root/leaf = (comc, hash)."""
    r = policy_deserialize(proof, policy)
    if r is None:
        return False
    aggregated, individual, begin = r
    mk = merkle_deserialize(proof, begin, dlen(hash_id))
    if mk is None:
        return False
    height, idx, sibs, end = mk
    if len(sibs) != height:
        return False
    pt = decompress(leaf[0])
    if pt is None:
        return False
    cur = (pt, leaf[0], leaf[1])
    for lvl, (comc, h) in enumerate(sibs):  # MerkleProof::verify folds upward from the leaf
        sib = (decompress(comc), comc, h)
        bit = (idx >> lvl) & 1
        cur = proofnode_merge(hash_id, sib, cur) if bit else proofnode_merge(hash_id, cur, sib)
    if cur[1] != root[0] or cur[2] != root[1]:
        return False
    return policy_verify(aggregated, individual, [c for c, _ in sibs], policy)


# --------------------------------------------------------------------------------------
# Batch proofs: ONE DapolProof for several leaves (Dapol::generate_proof_batch, src/dapol/mod.rs:172-190;
# DapolProof::verify_batch, src/proof/mod.rs:49-54; test shape src/proof/tests.rs:6-35).
# smtree get_merkle_path_ref_batch / MerkleProof::verify_batch (UPSTREAM-RECALL): level by level from
# the leaves up, left to right, a node's sibling is listed only if it is not itself on the way up from the batch.
# --------------------------------------------------------------------------------------
def batch_sibling_plan(height: int, leaf_idxs):
    """[(level, index)] of the siblings a batch Merkle proof carries, in proof order; leaf_idxs strictly increasing."""
    assert all(a < b for a, b in zip(leaf_idxs, leaf_idxs[1:])), "batch indexes must be strictly increasing"
    plan, cur = [], list(leaf_idxs)
    for h in range(height, 0, -1):
        have = set(cur)
        plan += [(h, x ^ 1) for x in cur if (x ^ 1) not in have]
        nxt = []
        for x in cur:
            if not nxt or nxt[-1] != x >> 1:
                nxt.append(x >> 1)
        cur = nxt
    return plan


BATCH_KEY_LABEL = b"dapol-b200 batch proof nonce key v1"


def batch_nonce_key(key: bytes, leaf_idxs) -> bytes:
    """Nonce key of a batch proof of more than one leaf: the tree's prover key (prover_nonce_key) bound to the batch --
    BLAKE3 chained over the leaf indexes, 64 of them (512 B) per link; the proof's range proofs draw from stream 0."""
    st = digest(HASH_BLAKE3, BATCH_KEY_LABEL, key, struct.pack("<Q", len(leaf_idxs)))
    for i in range(0, len(leaf_idxs), 64):
        st = digest(HASH_BLAKE3, st, b"".join(struct.pack("<Q", x) for x in leaf_idxs[i:i + 64]))
    return st


def merkle_serialize_batch(height, leaf_idxs, siblings) -> bytes:
    out = height.to_bytes(2, "big") + len(leaf_idxs).to_bytes(8, "big")
    nbytes = (height + 7) // 8
    for x in leaf_idxs:
        out += ((x << (8 * nbytes - height)) if height else 0).to_bytes(nbytes, "big")
    out += len(siblings).to_bytes(8, "big")
    for comc, h in siblings:
        out += comc + h
    return out


def prove_inclusion_batch(tree: Tree, leaf_idxs, agg: int, policy: int, seed: bytes) -> bytes:
    """Dapol::generate_proof_batch(leaf_indexes) -> DapolProof::serialize; one leaf = prove_inclusion (mod.rs:167-169)."""
    leaf_idxs = list(leaf_idxs)
    if len(leaf_idxs) == 1:
        return prove_inclusion(tree, leaf_idxs[0], agg, policy, seed)
    sibs = [tree.levels[h][i] for h, i in batch_sibling_plan(tree.height, leaf_idxs)]
    root = tree.root
    key = batch_nonce_key(prover_nonce_key(seed, root.comc, root.hash, policy, agg, tree.height), leaf_idxs)
    aggregated, individual = policy_prove([s.v for s in sibs], [s.r % L for s in sibs], agg, policy, key, 0)
    return policy_serialize(aggregated, individual, policy) + merkle_serialize_batch(tree.height, leaf_idxs, [(s.comc, s.hash) for s in sibs])


def verify_inclusion_batch(hash_id, proof: bytes, policy: int, root, leaves) -> bool:
    """DapolProof::deserialize + verify_batch(root, leaves); root / leaves = (comc, hash), leaves in index order."""
    dlen = 64 if hash_id == HASH_BLAKE2B else 32
    r = policy_deserialize(proof, policy)
    if r is None:
        return False
    aggregated, individual, begin = r
    try:
        height = int.from_bytes(proof[begin:begin + 2], "big"); begin += 2
        nb = int.from_bytes(proof[begin:begin + 8], "big"); begin += 8
        nbytes = (height + 7) // 8
        if height > 64 or nb != len(leaves) or nb == 0 or len(proof) - begin < nb * nbytes + 8:
            return False
        idxs = []
        for _ in range(nb):
            idxs.append(int.from_bytes(proof[begin:begin + nbytes], "big") >> (8 * nbytes - height) if height else 0)
            begin += nbytes
        k = int.from_bytes(proof[begin:begin + 8], "big"); begin += 8
        if len(proof) - begin < k * (32 + dlen):
            return False
        sibs = []
        for _ in range(k):
            sibs.append((proof[begin:begin + 32], proof[begin + 32:begin + 32 + dlen]))
            begin += 32 + dlen
    except Exception:
        return False
    if any(a >= b for a, b in zip(idxs, idxs[1:])):
        return False
    plan = batch_sibling_plan(height, idxs)
    if len(plan) != len(sibs):
        return False
    pts = [decompress(c) for c, _ in leaves] + [decompress(c) for c, _ in sibs]
    if any(p is None for p in pts):
        return False
    cur = {x: (decompress(c), c, h) for x, (c, h) in zip(idxs, leaves)}
    given = {hi: (decompress(c), c, h) for hi, (c, h) in zip(plan, sibs)}
    for lvl in range(height, 0, -1):
        nxt = {}
        for x in sorted(cur):
            if x >> 1 in nxt:
                continue
            other = cur.get(x ^ 1) or given[(lvl, x ^ 1)]
            l, r_ = (cur[x], other) if not x & 1 else (other, cur[x])
            nxt[x >> 1] = proofnode_merge(hash_id, l, r_)
        cur = nxt
    top = cur[0] if height else cur[idxs[0]]
    if top[1] != root[0] or top[2] != root[1]:
        return False
    return policy_verify(aggregated, individual, [c for c, _ in sibs], policy)
