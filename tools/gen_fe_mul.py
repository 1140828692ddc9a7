#!/usr/bin/env python3
"""Generates dapol_b200/csrc/fe_mulsqr_gen.inc: 256x256->512-bit multiply / square and the
2^256 = 38 (mod 2^255-19) fold, as PTX carry chains on 8x32-bit limbs.

Scheme: products a[j]*b[i] are (lo,hi) word pairs at position i+j.  Pairs at even positions
accumulate into E, pairs at odd positions into O (O is E shifted by one word), so that inside
one row the pairs of a chain never overlap and a single carry chain
(mad.lo.cc / madc.hi.cc ...) runs through them.  Result = E + (O << 32).

One instruction list is emitted twice: as an inline-asm block per chain for the device
(the carry flag never crosses an asm statement), and as EMU_* macro calls for a host build
(tests/host_emu) so the chain logic is checked on CPU against big integers.
"""
import os
import sys

# Tuning knobs (defaults = the variant measured fastest on B200, see profiles/fe_variants_r01.txt):
#   DAPOL_FE_PLAIN_MUL  : comma list of "row:parity" half-rows of the 8x8 product computed as plain (carry-free,
#                         full-rate) IMAD.WIDE products and merged with ALU-pipe add chains instead of the
#                         half-rate carry-predicated IMAD.WIDE.X chain
#   DAPOL_FE_PLAIN_SQR  : same for the off-diagonal rows of the squaring
#   DAPOL_FE_SQR_DIAG   : "chain" (mad chain) or "plain" (plain products + add chain)
#   DAPOL_FE_FOLD       : "chain" (original) or "plain"
def _pairs(env, default):
    v = os.environ.get(env, default)
    return set(tuple(int(x) for x in t.split(":")) for t in v.split(",") if t)


PLAIN_MUL = _pairs("DAPOL_FE_PLAIN_MUL", "")
PLAIN_SQR = _pairs("DAPOL_FE_PLAIN_SQR", "")
SQR_DIAG = os.environ.get("DAPOL_FE_SQR_DIAG", "chain")
FOLD = os.environ.get("DAPOL_FE_FOLD", "plain")
OUT = os.environ.get("DAPOL_FE_OUT") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dapol_b200", "csrc", "fe_mulsqr_gen.inc")


class Chain:
    """A straight-line block of carry-chained instructions."""

    def __init__(self):
        self.ins = []  # (op, dst, [srcs])

    def add(self, op, dst, *srcs):
        self.ins.append((op, dst, list(srcs)))

    def emit(self, fresh):
        """fresh: set of lvalues first written in this chain (output-only operands)."""
        if not self.ins:
            return ""
        # ---- device asm
        ops = []  # operand lvalues in order: outputs first
        outs = []
        for _, dst, _ in self.ins:
            if dst not in outs:
                outs.append(dst)
        ins_only = []
        for _, _, srcs in self.ins:
            for s in srcs:
                if s not in outs and s not in ins_only and not s.isdigit():
                    ins_only.append(s)
        ops = outs + ins_only
        num = {o: i for i, o in enumerate(ops)}
        # an output that is read before being written in this chain must be "+r"
        read_before_write = set()
        written = set()
        for _, dst, srcs in self.ins:
            for s in srcs:
                if s in num and s in outs and s not in written:
                    read_before_write.add(s)
            written.add(dst)
        lines = []
        for op, dst, srcs in self.ins:
            args = ", ".join(["%%%d" % num[dst]] + [s if s.isdigit() else "%%%d" % num[s] for s in srcs])
            lines.append('"%s.u32 %s;\\n\\t"' % (op, args))
        # early-clobber: an "=r" output written before all inputs are consumed must not share a register with an input
        out_c = ", ".join('"%s"(%s)' % ("+r" if o in read_before_write else "=&r", o) for o in outs)
        in_c = ", ".join('"r"(%s)' % o for o in ins_only)
        dev = "    asm(" + "\n        ".join(lines) + "\n        : " + out_c + "\n        : " + in_c + ");\n"
        # ---- host emulation
        emu = "    { uint32_t cf_ = 0; (void)cf_;\n"
        for op, dst, srcs in self.ins:
            emu += "      EMU_%s(%s, %s);\n" % (op.replace(".", "_").upper(), dst, ", ".join(srcs))
        emu += "    }\n"
        return "#ifdef __CUDA_ARCH__\n" + dev + "#else\n" + emu + "#endif\n"


class Acc:
    def __init__(self, name, limit):
        self.name, self.limit, self.live = name, limit, set()

    def w(self, i):
        return "%s[%d]" % (self.name, i)


def product_chain(acc: Acc, start: int, prods):
    """prods: list of (x, y) lvalues; k-th product -> words start+2k (lo), start+2k+1 (hi)."""
    ch = Chain()
    carry = False
    fresh = set()
    last_hi_was_live = False
    for k, (x, y) in enumerate(prods):
        for word, part in ((start + 2 * k, "lo"), (start + 2 * k + 1, "hi")):
            assert word < acc.limit, (acc.name, word)
            live = word in acc.live
            dst = acc.w(word)
            if live:
                ch.add(("madc.%s.cc" if carry else "mad.%s.cc") % part, dst, x, y, dst)
                carry = True
            else:
                if carry:
                    ch.add("madc.%s.cc" % part, dst, x, y, "0")
                else:
                    ch.add("mul.%s" % part, dst, x, y)
                fresh.add(dst)
                acc.live.add(word)
            last_hi_was_live = live
    if carry and last_hi_was_live:
        word = start + 2 * len(prods)
        while word < acc.limit:
            dst = acc.w(word)
            if word in acc.live:
                ch.add("addc.cc", dst, dst, "0")
                word += 1
            else:
                ch.add("addc", dst, "0", "0")
                fresh.add(dst)
                acc.live.add(word)
                break
    # the final instruction never needs to set the flag
    if ch.ins:
        op, dst, srcs = ch.ins[-1]
        if op.endswith(".cc") and op.startswith(("madc", "addc")):
            ch.ins[-1] = (op[:-3], dst, srcs)
    return ch.emit(fresh)


_tmp_counter = [0]


def plain_product_chain(acc: Acc, start: int, prods):
    """Same contract as product_chain, but the products are carry-free 64-bit multiplies (full-rate IMAD.WIDE)
    and only the accumulation is a carry chain (ALU pipe)."""
    _tmp_counter[0] += 1
    t = "pt%d_" % _tmp_counter[0]
    n = len(prods)
    s = "    uint32_t %s[%d];\n" % (t, 2 * n)
    for k, (x, y) in enumerate(prods):
        s += "    { uint64_t p_ = (uint64_t)%s * (uint64_t)%s; %s[%d] = (uint32_t)p_; %s[%d] = (uint32_t)(p_ >> 32); }\n" % (x, y, t, 2 * k, t, 2 * k + 1)
    ch = Chain()
    carry = False
    fresh = set()
    last_was_live = False
    for k in range(2 * n):
        word = start + k
        assert word < acc.limit
        dst, src = acc.w(word), "%s[%d]" % (t, k)
        live = word in acc.live
        if live:
            ch.add("addc.cc" if carry else "add.cc", dst, dst, src)
            carry = True
        else:
            if carry:
                ch.add("addc.cc", dst, src, "0")
            else:
                s += "    %s = %s;\n" % (dst, src)
            fresh.add(dst)
            acc.live.add(word)
        last_was_live = live
    if carry and last_was_live:
        word = start + 2 * n
        while word < acc.limit:
            dst = acc.w(word)
            if word in acc.live:
                ch.add("addc.cc", dst, dst, "0")
                word += 1
            else:
                ch.add("addc", dst, "0", "0")
                fresh.add(dst)
                acc.live.add(word)
                break
    if ch.ins:
        op, dst, srcs = ch.ins[-1]
        if op.endswith(".cc") and op.startswith("addc"):
            ch.ins[-1] = (op[:-3], dst, srcs)
    return s + ch.emit(fresh)


def merge_EO(R, E: Acc, O: Acc, n):
    """R = E + (O << 32), n words."""
    out = "    %s[0] = %s;\n" % (R, E.w(0))
    ch = Chain()
    first = True
    for k in range(1, n):
        e = E.w(k) if k in E.live else None
        o = O.w(k - 1) if (k - 1) in O.live else None
        dst = "%s[%d]" % (R, k)
        a, b = (e or "0"), (o or "0")
        if a == "0":
            a, b = b, a
        if a == "0":  # both dead
            ch.add("add" if first else "addc.cc", dst, "0", "0")
        else:
            ch.add("add.cc" if first else "addc.cc", dst, a, b)
        first = False
    op, dst, srcs = ch.ins[-1]
    ch.ins[-1] = ("addc", dst, srcs)
    return out + ch.emit(set())


def gen_mul4(name):
    """256-bit product of two 128-bit halves: 16 wide multiply-accumulates on the E/O carry chains."""
    s = "DAPOL_HD_INLINE void %s(uint32_t R[8], const uint32_t a[4], const uint32_t b[4]) {\n" % name
    s += "    uint32_t E[8], O[7];\n"
    E, O = Acc("E", 8), Acc("O", 7)
    for i in range(4):
        for parity in (0, 1):
            js = [j for j in range(4) if j % 2 == parity]
            p0 = i + js[0]
            acc, start = (E, p0) if p0 % 2 == 0 else (O, p0 - 1)
            s += product_chain(acc, start, [("a[%d]" % j, "b[%d]" % i) for j in js])
    assert E.live == set(range(8)) and O.live == set(range(7)), (E.live, O.live)
    s += merge_EO("R", E, O, 8)
    s += "}\n\n"
    return s


def gen_mul_karatsuba():
    """a * b with ONE level of subtractive Karatsuba: 3 x 16 = 48 wide multiply-accumulates instead of 64.
    a = a0 + a1 2^128, b = b0 + b1 2^128:  a b = z0 + (z0 + z2 + (a0 - a1)(b1 - b0)) 2^128 + z2 2^256 with z0 = a0 b0, z2 = a1 b1.
    The differences are taken as absolute values with a sign mask, so every product is 128 x 128 bits; the additions run on the
    ALU pipe, which the schoolbook product leaves two thirds idle, while the multiply pipe (the bound) does a quarter less."""
    s = gen_mul4("mul_wide_4x4")
    s += "DAPOL_HD_INLINE void mul_wide_8x8(uint32_t R[16], const uint32_t a[8], const uint32_t b[8]) {\n"
    s += "    uint32_t Z0[8], Z2[8], ZM[8], da[4], db[4], M[9], sa, sb, sg, cin;\n"
    s += "    mul_wide_4x4(Z0, a, b);\n    mul_wide_4x4(Z2, a + 4, b + 4);\n"
    # da = |a0 - a1| with sa = all-ones if a0 < a1; db = |b1 - b0| with sb
    for d, x, y, m in (("da", "a[%d]", "a[%d]", "sa"), ("db", "b[%d]", "b[%d]", "sb")):
        lo, hi = (0, 4) if d == "da" else (4, 0)
        ch = Chain()
        for k in range(4):
            ch.add("sub.cc" if k == 0 else "subc.cc", "%s[%d]" % (d, k), x % (lo + k), y % (hi + k))
        ch.add("subc", m, "0", "0")  # 0 - 0 - borrow: all ones iff the difference is negative
        s += ch.emit(set())
        s += "    " + " ".join("%s[%d] ^= %s;" % (d, k, m) for k in range(4)) + "\n"
        ch = Chain()
        for k in range(4):  # (x ^ m) - m = -x when m is all ones
            ch.add("sub.cc" if k == 0 else ("subc.cc" if k < 3 else "subc"), "%s[%d]" % (d, k), "%s[%d]" % (d, k), m)
        s += ch.emit(set())
    s += "    mul_wide_4x4(ZM, da, db);\n"
    s += "    sg = sa ^ sb;  // all ones: the middle product is negative\n"
    # M = Z0 + Z2 (9 words)
    ch = Chain()
    for k in range(8):
        ch.add("add.cc" if k == 0 else "addc.cc", "M[%d]" % k, "Z0[%d]" % k, "Z2[%d]" % k)
    ch.add("addc", "M[8]", "0", "0")
    s += ch.emit(set())
    # M += sg ? -ZM : ZM   (two's complement over 9 words; the true value is non-negative and below 2^257)
    s += "    " + " ".join("ZM[%d] ^= sg;" % k for k in range(8)) + "\n"
    ch = Chain()
    ch.add("add.cc", "cin", "sg", "1")  # carry-in = 1 iff sg is all ones
    for k in range(8):
        ch.add("addc.cc", "M[%d]" % k, "M[%d]" % k, "ZM[%d]" % k)
    ch.add("addc", "M[8]", "M[8]", "sg")
    s += ch.emit(set())
    # R = Z0 | Z2 with M added at word 4
    s += "    " + " ".join("R[%d] = Z0[%d];" % (k, k) for k in range(4)) + "\n"
    ch = Chain()
    for k in range(4):
        ch.add("add.cc" if k == 0 else "addc.cc", "R[%d]" % (4 + k), "Z0[%d]" % (4 + k), "M[%d]" % k)
    for k in range(4):
        ch.add("addc.cc", "R[%d]" % (8 + k), "Z2[%d]" % k, "M[%d]" % (4 + k))
    ch.add("addc.cc", "R[12]", "Z2[4]", "M[8]")
    for k in range(13, 16):
        ch.add("addc.cc" if k < 15 else "addc", "R[%d]" % k, "Z2[%d]" % (k - 8), "0")
    s += ch.emit(set())
    s += "    (void)cin;\n}\n\n"
    return s


def gen_mul():
    # DAPOL_FE_MUL=karatsuba: measured and REJECTED on B200 (profiles/r02_variants.txt): 48 + 8 wide multiply-accumulates instead of
    # 64 + 8, but ptxas turns the extra carry chains into IMAD.X / IMAD.MOV on the same multiply pipe and the product needs 207
    # instructions instead of 139: fe_mul 112 -> 91 G/s, tree build 25.9 -> 28.7 ms.  The schoolbook product stays the default.
    if os.environ.get("DAPOL_FE_MUL", "schoolbook") == "karatsuba":
        return gen_mul_karatsuba()
    s = "DAPOL_HD_INLINE void mul_wide_8x8(uint32_t R[16], const uint32_t a[8], const uint32_t b[8]) {\n"
    s += "    uint32_t E[16], O[15];\n"
    E, O = Acc("E", 16), Acc("O", 15)
    for i in range(8):
        for parity in (0, 1):
            js = [j for j in range(8) if j % 2 == parity]
            p0 = i + js[0]
            acc, start = (E, p0) if p0 % 2 == 0 else (O, p0 - 1)
            fn = plain_product_chain if (i, parity) in PLAIN_MUL else product_chain
            s += fn(acc, start, [("a[%d]" % j, "b[%d]" % i) for j in js])
    assert E.live == set(range(16)) and O.live == set(range(15)), (E.live, O.live)
    s += merge_EO("R", E, O, 16)
    s += "}\n\n"
    return s


def gen_sqr():
    s = "DAPOL_HD_INLINE void sqr_wide_8(uint32_t R[16], const uint32_t a[8]) {\n"
    s += "    uint32_t E[16], O[15], U[16];\n"
    E, O = Acc("E", 16), Acc("O", 15)
    for i in range(7):
        for parity in (0, 1):
            js = [j for j in range(i + 1, 8) if (i + j) % 2 == parity]
            if not js:
                continue
            p0 = i + js[0]
            acc, start = (E, p0) if p0 % 2 == 0 else (O, p0 - 1)
            fn = plain_product_chain if (i, parity) in PLAIN_SQR else product_chain
            s += fn(acc, start, [("a[%d]" % i, "a[%d]" % j) for j in js])
    # off-diagonal sum U = E + (O<<32); words never written are zero
    for k in range(16):
        if k not in E.live:
            s += "    E[%d] = 0;\n" % k
            E.live.add(k)
    for k in range(15):
        if k not in O.live:
            s += "    O[%d] = 0;\n" % k
            O.live.add(k)
    s += merge_EO("U", E, O, 16)
    # double
    ch = Chain()
    for k in range(16):
        ch.add("add.cc" if k == 0 else ("addc.cc" if k < 15 else "addc"), "U[%d]" % k, "U[%d]" % k, "U[%d]" % k)
    s += ch.emit(set())
    # add diagonal squares
    if SQR_DIAG == "plain":
        s += "    uint32_t dg_[16];\n"
        for i in range(8):
            s += "    { uint64_t p_ = (uint64_t)a[%d] * (uint64_t)a[%d]; dg_[%d] = (uint32_t)p_; dg_[%d] = (uint32_t)(p_ >> 32); }\n" % (i, i, 2 * i, 2 * i + 1)
        ch = Chain()
        for k in range(16):
            ch.add("add.cc" if k == 0 else ("addc.cc" if k < 15 else "addc"), "R[%d]" % k, "U[%d]" % k, "dg_[%d]" % k)
        s += ch.emit(set())
        s += "}\n\n"
        return s
    ch = Chain()
    for i in range(8):
        ch.add("mad.lo.cc" if i == 0 else "madc.lo.cc", "R[%d]" % (2 * i), "a[%d]" % i, "a[%d]" % i, "U[%d]" % (2 * i))
        ch.add("madc.hi.cc" if i < 7 else "madc.hi", "R[%d]" % (2 * i + 1), "a[%d]" % i, "a[%d]" % i, "U[%d]" % (2 * i + 1))
    s += ch.emit(set())
    s += "}\n\n"
    return s


def gen_fold():
    """r = R[0..8) + 38 * R[8..16)  mod 2^256-38 (i.e. a representative < 2^256 of the value mod p)."""
    s = "DAPOL_HD_INLINE void fold38(uint32_t r[8], const uint32_t R[16]) {\n"
    if FOLD == "plain":
        s += "    uint32_t t8, fl_[8], fh_[8];\n"
        for k in range(8):
            s += "    { uint64_t p_ = (uint64_t)R[%d] * 38ull; fl_[%d] = (uint32_t)p_; fh_[%d] = (uint32_t)(p_ >> 32); }\n" % (8 + k, k, k)
        ch = Chain()
        for k in range(8):
            ch.add("add.cc" if k == 0 else "addc.cc", "r[%d]" % k, "R[%d]" % k, "fl_[%d]" % k)
        ch.add("addc", "t8", "0", "0")
        s += ch.emit(set())
        ch = Chain()
        for k in range(7):
            ch.add("add.cc" if k == 0 else "addc.cc", "r[%d]" % (k + 1), "r[%d]" % (k + 1), "fh_[%d]" % k)
        ch.add("addc", "t8", "t8", "fh_[7]")
        s += ch.emit(set())
        s += "    uint32_t m_ = t8 * 38u;\n"
        ch = Chain()
        ch.add("add.cc", "r[0]", "r[0]", "m_")
        for k in range(1, 8):
            ch.add("addc.cc", "r[%d]" % k, "r[%d]" % k, "0")
        ch.add("addc", "t8", "0", "0")
        s += ch.emit(set())
        s += "    r[0] += t8 * 38u;\n"
        s += "}\n\n"
        return s
    s += "    uint32_t t8, c38 = 38u;\n"
    ch = Chain()
    for k in range(8):
        ch.add("mad.lo.cc" if k == 0 else "madc.lo.cc", "r[%d]" % k, "R[%d]" % (8 + k), "c38", "R[%d]" % k)
    ch.add("addc", "t8", "0", "0")
    s += ch.emit(set())
    ch = Chain()
    for k in range(7):
        ch.add("mad.hi.cc" if k == 0 else "madc.hi.cc", "r[%d]" % (k + 1), "R[%d]" % (8 + k), "c38", "r[%d]" % (k + 1))
    ch.add("madc.hi", "t8", "R[15]", "c38", "t8")
    s += ch.emit(set())
    # t8 < 2^7: fold again; a second carry-out can only leave a tiny value, so +38 cannot overflow again
    ch = Chain()
    ch.add("mad.lo.cc", "r[0]", "t8", "c38", "r[0]")
    for k in range(1, 8):
        ch.add("addc.cc", "r[%d]" % k, "r[%d]" % k, "0")
    ch.add("addc", "t8", "0", "0")
    s += ch.emit(set())
    s += "    r[0] += t8 * 38u;\n"
    s += "}\n\n"
    return s


def main():
    hdr = "// GENERATED by tools/gen_fe_mul.py -- do not edit.\n"
    hdr += "// 8x32-bit limb multiply/square carry chains (E/O split accumulators) + fold by 38.\n\n"
    body = gen_mul() + gen_sqr() + gen_fold()
    with open(OUT, "w") as f:
        f.write(hdr + body)
    n_mad = body.count("mad") + body.count("mul.")
    print("wrote", os.path.normpath(OUT), "lines:", body.count("\n"), file=sys.stderr)


if __name__ == "__main__":
    main()
