// TEST HARNESS ONLY: the range-proof kernel bodies (dapol_b200/csrc/rp_kernels.cuh) compiled for the HOST and driven
// in serial loops that mirror the CUDA orchestration of dapol_rp.cu (same passes, same order, several emulated
// "threads" per CTA so the strided loops and the partial-sum reductions are exercised).  Checked against the oracle by
// tests/test_host_emu_rp.py.  Never linked into or loaded by the product library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
// the switch to the folded-generator rounds (hybrid inner-product argument) is a run-time knob here, so that the small
// shapes a CPU can afford exercise both the original-generator rounds and the hybrid ones
static int g_hybrid_min_n = 1024;
#define RP_HYBRID_MIN_N g_hybrid_min_n
#include "../../dapol_b200/csrc/rp_kernels.cuh"

#define EX extern "C" __attribute__((visibility("default")))
constexpr int W = 4;
constexpr int NW = 253 / W + 1;
constexpr uint32_t HALF = 1u << (W - 1);

struct Tables {
    int mcap = 0;
    std::vector<ge_niels> tab;
    std::vector<uint32_t> ext;
};
static Tables g_tab;

static void ensure_tables(int m) {
    if (g_tab.mcap >= m) return;
    int mc = 1;
    while (mc < m) mc <<= 1;
    uint64_t nb = 128ull * mc + 2;
    std::vector<uint32_t> uniform(2ull * mc * 64 * 16);
    for (int t = 0; t < 2 * mc; t++) rp_gen_chain_body((uint32_t)(t % mc), t / mc, uniform.data() + (uint64_t)t * 64 * 16);
    g_tab.ext.assign(nb * 32, 0);
    for (uint64_t g = 0; g < 128ull * mc; g++) rp_gen_point_body(g, uniform.data(), g_tab.ext.data());
    ge B, Bb;
    ge_basepoint(B); ge_bblinding(Bb);
    rp_store_ext(g_tab.ext.data() + 32 * (128ull * mc), B);
    rp_store_ext(g_tab.ext.data() + 32 * (128ull * mc + 1), Bb);
    std::vector<uint32_t> wb(nb * NW * 32);
    for (uint64_t g = 0; g < nb; g++) rp_tab_windows_body<W>(g, g_tab.ext.data(), wb.data());
    g_tab.tab.resize(nb * NW * HALF);
    constexpr uint32_t C = HALF < RP_TAB_CHUNK ? HALF : RP_TAB_CHUNK;
    for (uint64_t it = 0; it < nb * NW * (HALF / C); it++) rp_tab_chunk_body<W>(it, wb.data(), g_tab.tab.data());
    g_tab.mcap = mc;
}
// compressed generator (is_h, party, i) -- checked against the oracle's BulletproofGens
EX void emu_rp_gen(int is_h, uint32_t party, uint32_t i, uint8_t out[32]) {
    ensure_tables((int)party + 1);
    ge p;
    rp_load_ext(p, g_tab.ext.data() + 32 * ((uint64_t)(is_h ? 64 * g_tab.mcap : 0) + 64ull * party + i));
    uint32_t c[8];
    ge_compress(c, p);
    memcpy(out, c, 32);
}
// s * P via the fixed-base table of generator g == via variable-base scalar mult of the same point
EX int emu_rp_table_check(uint32_t g, const uint8_t s[32]) {
    ensure_tables(1);
    sc k;
    memcpy(k.v, s, 32);
    sc_reduce256(k, k);
    ge a, b, p;
    ge_identity(a);
    rp_fixed_mul_acc<W>(a, g_tab.tab.data(), g, k);
    rp_load_ext(p, g_tab.ext.data() + 32ull * g);
    ge_scalarmult_var(b, k, p);
    uint32_t ca[8], cb[8];
    ge_compress(ca, a); ge_compress(cb, b);
    return memcmp(ca, cb, 32) == 0;
}
EX void emu_merlin_test(const uint8_t *label, uint32_t llen, const uint8_t *mlabel, uint32_t mllen, const uint8_t *msg, uint32_t mlen,
                        const uint8_t *clabel, uint32_t cllen, uint8_t out[32]) {
    merlin t;
    strobe_init(t);
    tr_append_bytes(t, "dom-sep", 7, label, llen);
    tr_append_bytes(t, reinterpret_cast<const char *>(mlabel), mllen, msg, mlen);
    sc c;
    tr_challenge_scalar(t, reinterpret_cast<const char *>(clabel), cllen, c);
    memcpy(out, c.v, 32);
}

struct Bufs {
    std::vector<merlin> tr;
    std::vector<uint32_t> Vc, blr, chal, zpow, mult, vecA, vecB, ypow, svec, cu0, cu1, cui0, cui1, pts, ptc, varpts, varsc, vartab, gfold, proof;
    std::vector<int> status;
};
static void setup(RpBatch &b, Bufs &u, int nbits, int m, uint64_t K) {
    ensure_tables(m);
    memset(&b, 0, sizeof b);
    b.nbits = nbits; b.m = m; b.N = nbits * m; b.K = K;
    b.lg = 0;
    while ((1 << b.lg) < b.N) b.lg++;
    b.plen = 32 * (9 + 2 * b.lg);
    int nv = rp_nvar(b.lg, m);
    uint64_t N = b.N;
    u.tr.resize(K); u.Vc.resize(K * m * 8); u.blr.resize(K * m * 8); u.chal.resize(K * CH_COUNT * 8); u.zpow.resize(K * m * 8);
    u.mult.resize(K * 3 * 32 * 8); u.vecA.resize(K * N * 8); u.vecB.resize(K * N * 8); u.ypow.resize(K * N * 8); u.svec.resize(K * N * 8);
    u.cu0.resize(K * N / 2 * 8); u.cu1.resize(K * N / 2 * 8); u.cui0.resize(K * N / 2 * 8); u.cui1.resize(K * N / 2 * 8);
    u.pts.resize(K * 2 * 32); u.ptc.resize(K * 2 * 8); u.gfold.resize(K * 2 * RP_FOLD_N * 32); u.varpts.resize(K * nv * 32); u.varsc.resize(K * nv * 8); u.vartab.resize(K * nv * 8 * 32); u.proof.resize(K * b.plen / 4);
    u.status.resize(K);
    b.tr = u.tr.data(); b.Vc = u.Vc.data(); b.blr = u.blr.data(); b.chal = u.chal.data(); b.zpow = u.zpow.data(); b.mult = u.mult.data();
    b.vecA = u.vecA.data(); b.vecB = u.vecB.data(); b.ypow = u.ypow.data(); b.svec = u.svec.data();
    b.cu[0] = u.cu0.data(); b.cu[1] = u.cu1.data(); b.cui[0] = u.cui0.data(); b.cui[1] = u.cui1.data();
    b.pts = u.pts.data(); b.ptc = u.ptc.data(); b.gfold = u.gfold.data(); b.varpts = u.varpts.data(); b.varsc = u.varsc.data(); b.vartab = u.vartab.data(); b.proof = u.proof.data(); b.status = u.status.data();
    uint64_t per = (uint64_t)NW * HALF;
    b.tabG = g_tab.tab.data();
    b.tabH = b.tabG + 64ull * g_tab.mcap * per;
    b.tabB = b.tabG + 128ull * g_tab.mcap * per;
    b.tabBbl = b.tabB + per;
}
static void expand(const RpBatch &b, uint32_t *vec, int slot) {
    for (uint64_t p = 0; p < b.K; p++)
        for (int s = 0; s < b.lg; s++)
            for (uint32_t i = 0; i < (1u << s); i++) rp_expand_step(vec + p * b.N * 8, b.mult + (p * 3 + slot) * 32 * 8, s, i);
}

// comb tables of the tree (window 4) for V_j = commit(v_j, r_j)
static std::vector<ge_niels> g_tb, g_tbbl;
template <int WT>
static void comb_entry(uint64_t t, ge_niels *table, int nw, int which) {  // same construction as tree_kernels.cuh comb_table_body
    uint32_t half = 1u << (WT - 1);
    uint32_t k = (uint32_t)(t / half), e = (uint32_t)(t % half);
    if ((int)k >= nw) return;
    ge base;
    if (which == 0) ge_basepoint_half(base); else ge_bblinding(base);  // tab_b = multiples of B/2 (tree_kernels.cuh)
    for (uint32_t i = 0; i < k * WT; i++) ge_dbl(base, base);
    ge acc = base;
    for (uint32_t i = 0; i < e; i++) ge_add(acc, acc, base);
    ge_niels n;
    ge_to_niels(n, acc);
    table[t] = n;
}

EX void emu_rp_set_hybrid_min_n(int n) { g_hybrid_min_n = n; }
EX int emu_rp_prove(int nbits, int m, uint64_t K, const uint64_t *values, const uint8_t *blindings, const uint8_t seed[32],
                    const uint64_t *stream, const uint64_t *base_block, int T, uint8_t *out) {
    RpBatch b;
    Bufs u;
    setup(b, u, nbits, m, K);
    if (g_tb.empty()) {
        constexpr int NWR = 253 / 4 + 1;
        g_tb.resize(NWR * 8); g_tbbl.resize(NWR * 8);
        for (uint64_t t = 0; t < g_tb.size(); t++) { comb_entry<4>(t, g_tb.data(), NWR, 0); comb_entry<4>(t, g_tbbl.data(), NWR, 1); }
    }
    b.values = values; b.blind = reinterpret_cast<const uint32_t *>(blindings); b.stream = stream; b.base_block = base_block;
    memcpy(b.seed, seed, 32);
    const uint32_t N = b.N;
    for (uint64_t p = 0; p < K; p++) rp_p0_body(b, p);
    for (uint64_t p = 0; p < K; p++) for (int j = 0; j < m; j++) rp_p1_body<4>(b, p, j, g_tb.data(), g_tbbl.data());
    for (uint64_t p = 0; p < K; p++) for (uint32_t k = 0; k < N; k++) rp_p2_body(b, p, k);
    auto msm = [&](auto partial) {  // sum of the per-thread partial points of a CTA
        for (uint64_t p = 0; p < K; p++)
            for (int which = 0; which < 2; which++) {
                ge sum, part;
                ge_identity(sum);
                for (int t = 0; t < T; t++) { partial(part, p, which, (uint32_t)t); ge_add(sum, sum, part); }
                rp_store_point(b, p, which, sum);
            }
    };
    msm([&](ge &acc, uint64_t p, int which, uint32_t t) { rp_p3_partial<W>(acc, b, p, which, t, (uint32_t)T); });
    auto compress_pts = [&]() { for (uint64_t pw = 0; pw < 2 * K; pw++) rp_compress_point_body(b, pw); };
    compress_pts();
    for (uint64_t p = 0; p < K; p++) rp_p4_body(b, p);
    expand(b, b.ypow, 0);
    for (uint64_t p = 0; p < K; p++) {
        sc t0, t1, t2, a0, a1, a2;
        sc_set_u64(t0, 0); t1 = t0; t2 = t0;
        for (int t = 0; t < T; t++) { rp_p5_partial(a0, a1, a2, b, p, (uint32_t)t, (uint32_t)T); sc_add(t0, t0, a0); sc_add(t1, t1, a1); sc_add(t2, t2, a2); }
        rp_st(rp_ch(b, p, CH_T0), t0); rp_st(rp_ch(b, p, CH_T1), t1); rp_st(rp_ch(b, p, CH_T2), t2);
    }
    for (uint64_t p = 0; p < K; p++) for (int which = 0; which < 2; which++) rp_p6_body<W>(b, p, which);
    compress_pts();
    for (uint64_t p = 0; p < K; p++) rp_p7_body(b, p);
    for (uint64_t p = 0; p < K; p++) for (uint32_t k = 0; k < N; k++) rp_p8_body(b, p, k);
    expand(b, b.ypow, 1);
    const int sw = rp_switch_round(b.N, b.lg);
    for (int rnd = 1; rnd <= b.lg; rnd++) {
        for (uint64_t p = 0; p < K; p++) {
            sc cl, cr, a0, a1;
            sc_set_u64(cl, 0); cr = cl;
            for (int t = 0; t < T; t++) { rp_p9_partial(a0, a1, b, p, rnd, (uint32_t)t, (uint32_t)T); sc_add(cl, cl, a0); sc_add(cr, cr, a1); }
            rp_st(rp_ch(b, p, CH_CL), cl); rp_st(rp_ch(b, p, CH_CR), cr);
        }
        uint32_t h = N >> rnd, cnt = h > (1u << (rnd - 1)) ? h : (1u << (rnd - 1));
        if (rnd == sw)
            for (uint64_t p = 0; p < K; p++) for (int which = 0; which < 2; which++) for (uint32_t j = 0; j < RP_FOLD_N; j++) rp_pm_body<W>(b, p, which, j, sw);
        if (rnd >= sw) {
            for (uint64_t p = 0; p < K; p++)
                for (int which = 0; which < 2; which++) {
                    ge sum, part;
                    ge_identity(sum);
                    for (uint32_t t = 0; t < 2 * h; t++) { rp_pv_partial<W>(part, b, p, rnd, which, t, 2 * h); ge_add(sum, sum, part); }
                    rp_store_point(b, p, which, sum);
                }
        } else {
            msm([&](ge &acc, uint64_t p, int which, uint32_t t) { rp_p10_partial<W>(acc, b, p, rnd, which, t, (uint32_t)T); });
        }
        compress_pts();
        for (uint64_t p = 0; p < K; p++) rp_p11_body(b, p, rnd);
        // a thread reads a[i], a[h+i] and writes a[i]; table growth reads cur, writes nxt: any order is fine
        for (uint64_t p = 0; p < K; p++) for (uint32_t i = 0; i < cnt; i++) rp_p12_body(b, p, rnd, i);
        if (rnd >= sw && rnd < b.lg)
            for (uint64_t p = 0; p < K; p++) for (int which = 0; which < 2; which++) for (uint32_t i = 0; i < h; i++) rp_pf_body(b, p, rnd, which, i);
    }
    memcpy(out, b.proof, K * b.plen);
    int rc = 0;
    for (uint64_t p = 0; p < K; p++) rc |= b.status[p];
    return rc;
}

static int g_vgroups = 8;
EX void emu_rp_set_vgroups(int g) { g_vgroups = g; }
EX void emu_rp_verify(int nbits, int m, uint64_t K, const uint8_t *proofs, const uint8_t *coms, int T, uint8_t *ok) {
    RpBatch b;
    Bufs u;
    setup(b, u, nbits, m, K);
    b.proof_in = reinterpret_cast<const uint32_t *>(proofs);
    b.coms = reinterpret_cast<const uint32_t *>(coms);
    const int nv = rp_nvar(b.lg, m);
    for (uint64_t p = 0; p < K; p++) rp_v0_body(b, p);
    expand(b, b.svec, 2);
    expand(b, b.ypow, 1);
    // threads per proof of V1 as rp_plan picks them for a small batch (8, or more when a thread would exceed RP_V1_PMAX points);
    // odd proofs exercise other group counts through the same body
    int g = g_vgroups;
    while (g * RP_V1_PMAX < nv) g <<= 1;
    b.vgroups = g < nv ? g : nv;
    for (uint64_t p = 0; p < K; p++) for (int q = 0; q < b.vgroups; q++) rp_v1_body(b, p, q);
    for (uint64_t p = 0; p < K; p++) {
        ge sum, part;
        ge_identity(sum);
        for (int t = 0; t < T; t++) { rp_v2_partial<W>(part, b, p, (uint32_t)t, (uint32_t)T); ge_add(sum, sum, part); }
        ok[p] = (uint8_t)(b.status[p] && ge_is_identity(sum));
    }
}

// Batched verifier (bucket method, "batched verification" in rp_kernels.cuh) in the order rp_verify_chunk_batched runs it:
// V0 + expansions, weights, terms, a stable sort of the terms by bucket id, buckets, chunk fold, windows, combined scalars,
// fixed part with T emulated threads per group, group verdicts.  gok[g] = verdict of group g.
#include <algorithm>
#include <numeric>
EX void emu_rp_verify_batched(int nbits, int m, uint64_t K, const uint8_t *proofs, const uint8_t *coms, int T, int G, int cbits, const uint8_t wseed[32],
                              int *gok) {
    RpBatch b;
    Bufs u;
    setup(b, u, nbits, m, K);
    b.proof_in = reinterpret_cast<const uint32_t *>(proofs);
    b.coms = reinterpret_cast<const uint32_t *>(coms);
    const uint64_t nv = (uint64_t)rp_nvar(b.lg, m), N = (uint64_t)b.N;
    for (uint64_t p = 0; p < K; p++) rp_v0_body(b, p);
    expand(b, b.svec, 2);
    expand(b, b.ypow, 1);
    RpbPlan pl;
    memset(&pl, 0, sizeof pl);
    pl.G = G; pl.groups = (K + G - 1) / G; pl.P = 2;
    rpb_set_windows(pl, cbits);
    const uint64_t nb = 1ull << (cbits - 1);
    pl.L = (int)std::min<uint64_t>(4, nb);  // small chunks so that several chunks per window are exercised
    const uint64_t n_terms = K * nv * pl.NW, n_buckets = pl.groups * pl.NW * nb, n_chunks = n_buckets / pl.L, n_gw = pl.groups * pl.NW;
    memcpy(pl.wseed, wseed, 32);
    std::vector<uint32_t> rho(K * 8), cached(K * nv * 32), keys_in(n_terms), keys(n_terms), vals_in(n_terms), vals(n_terms), bucket(n_buckets * 32),
        crun(n_chunks * 32), ctot(n_chunks * 32), window(n_gw * 32), gsc(pl.groups * (2 * N + 2) * 8), gfix(pl.groups * 32),
        gpart(pl.groups * rpb_parts(pl) * (2 * N + 2) * 8);
    std::vector<int> gbad(pl.groups, 0);
    pl.gpart = gpart.data(); pl.gbad = gbad.data();
    pl.rho = rho.data(); pl.cached = cached.data(); pl.keys_in = keys_in.data(); pl.keys = keys.data(); pl.vals_in = vals_in.data(); pl.vals = vals.data();
    pl.bucket = bucket.data(); pl.chunk_run = crun.data(); pl.chunk_tot = ctot.data(); pl.window = window.data(); pl.gsc = gsc.data(); pl.gfix = gfix.data();
    pl.gok = gok;
    for (uint64_t p = 0; p < K; p++) rpb_weight_body(pl, p);
    for (uint64_t p = 0; p < K; p++) for (uint64_t q = 0; q < nv; q++) rpb_terms_body(b, pl, p, (int)q);
    std::vector<uint64_t> order(n_terms);
    std::iota(order.begin(), order.end(), 0ull);
    std::stable_sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) { return keys_in[x] < keys_in[y]; });
    for (uint64_t j = 0; j < n_terms; j++) { keys[j] = keys_in[order[j]]; vals[j] = vals_in[order[j]]; }
    for (uint64_t bk = 0; bk < n_buckets; bk++) rpb_bucket_body(pl, bk, n_terms);
    for (uint64_t ch = 0; ch < n_chunks; ch++) rpb_chunk_body(pl, ch);
    for (uint64_t gw = 0; gw < n_gw; gw++) rpb_window_body(pl, gw);
    for (uint64_t g = 0; g < pl.groups; g++)
        for (uint64_t part = 0; part < rpb_parts(pl); part++) for (uint32_t t = 0; t < 2 * N + 2; t++) rpb_combine_body(b, pl, g, part, t);
    for (uint64_t g = 0; g < pl.groups; g++) for (uint32_t t = 0; t < 2 * N + 2; t++) rpb_combine_sum_body(b, pl, g, t);
    for (uint64_t p = 0; p < K; p++) rpb_status_body(b, pl, p);
    for (uint64_t g = 0; g < pl.groups; g++) {
        ge sum, part;
        ge_identity(sum);
        for (int t = 0; t < T; t++) { rpb_fixed_partial<W>(part, b, pl, g, (uint32_t)t, (uint32_t)T); ge_add(sum, sum, part); }
        rp_store_ext(pl.gfix + g * 32, sum);
    }
    for (uint64_t g = 0; g < pl.groups; g++) rpb_group_body(b, pl, g);
}
