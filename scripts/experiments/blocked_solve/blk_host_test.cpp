// Host emulation test of the blocked substitution solve (metmhn_b200/csrc/mmh_blk.cuh): the SAME per-lane functions the
// CUDA kernel runs are compiled for the host, a warp is emulated lane by lane (shared memory = an array, reads of a step
// come from a snapshot taken at the start of the step, exactly what __syncwarp guarantees on the device) and the result
// is compared with a plain sequential substitution in index order.  Built and run by tests/test_blk_host.py (no GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "mmh_blk.cuh"

using namespace mmh;

struct Space {
    int K = 0;
    BlkCtx ctx{};
    std::vector<std::vector<double>> P, Q;     // tables behind the descriptors
    std::vector<double> d1, d2, rhs;
    std::vector<int> group;                    // group of every bit
};

// rate tables of a group: T[t][u] = base_t * prod_{j in u, j != t} W[t][j] over the group's bits (local indices)
static std::vector<double> group_table(int bits, int t, const std::vector<std::vector<double>>& W, double base, const std::vector<int>& ids)
{
    std::vector<double> T((size_t)1 << bits);
    for (uint32_t u = 0; u < (1u << bits); ++u) {
        double r = base;
        for (int j = 0; j < bits; ++j) if (((u >> j) & 1u) && ids[j] != ids[t]) r *= W[ids[t]][ids[j]];
        T[u] = r;
    }
    return T;
}

// kind 0: one group, split tables (P = low 8 bits, Q = the rest); kind 1: pair, group A = bits 0..KA-1, B = the rest
static Space make_space(int K, int kind, int KA, std::mt19937_64& rng)
{
    Space sp;
    sp.K = K;
    std::uniform_real_distribution<double> U(0.5, 1.6);
    std::vector<std::vector<double>> W(K, std::vector<double>(K));
    for (auto& r : W) for (auto& v : r) v = U(rng);
    std::vector<double> base(K);
    for (auto& v : base) v = U(rng);
    sp.P.resize(K); sp.Q.resize(K);
    BlkCtx& c = sp.ctx;
    std::memset(&c, 0, sizeof(c));
    if (kind == 0) {
        std::vector<int> lo(8), hi(K - 8);
        for (int j = 0; j < 8; ++j) lo[j] = j;
        for (int j = 8; j < K; ++j) hi[j - 8] = j;
        for (int t = 0; t < K; ++t) {
            // rate = T1[u & 255] * T2[u >> 8]; the own bit is skipped by the ids comparison
            std::vector<int> ids_lo = lo, ids_hi = hi;
            std::vector<double> T1((size_t)256), T2((size_t)1 << (K - 8));
            for (uint32_t u = 0; u < 256; ++u) { double r = base[t]; for (int j = 0; j < 8; ++j) if (((u >> j) & 1u) && j != t) r *= W[t][j]; T1[u] = r; }
            for (uint32_t u = 0; u < (1u << (K - 8)); ++u) { double r = 1.0; for (int j = 8; j < K; ++j) if (((u >> (j - 8)) & 1u) && j != t) r *= W[t][j]; T2[u] = r; }
            sp.P[t] = T1; sp.Q[t] = T2;
        }
        for (int t = 0; t < K; ++t) c.bit[t] = {sp.P[t].data(), sp.Q[t].data(), 255u, (1u << (K - 8)) - 1u, 8u, -1};
        sp.d1.resize((size_t)1 << K);
        for (auto& v : sp.d1) v = 1.0 + U(rng);
        c.d1 = sp.d1.data(); c.m1 = (1u << K) - 1u; c.d2 = nullptr; c.m2 = 0; c.sh2 = 0;
        blk_ctx_layout(c, K, 8);
    } else {
        const int KB = K - KA;
        for (int t = 0; t < K; ++t) {
            const bool a = t < KA;
            const int bits = a ? KA : KB, off = a ? 0 : KA;
            std::vector<double> T((size_t)1 << bits);
            for (uint32_t u = 0; u < (1u << bits); ++u) { double r = base[t]; for (int j = 0; j < bits; ++j) if (((u >> j) & 1u) && j + off != t) r *= W[t][j + off]; T[u] = r; }
            if (a || KA == 0) { sp.P[t] = T; c.bit[t] = {sp.P[t].data(), nullptr, (1u << bits) - 1u, 0u, 8u, -1}; }
            else { sp.Q[t] = T; c.bit[t] = {nullptr, sp.Q[t].data(), 1u, (1u << KB) - 1u, (uint32_t)KA, -1}; }
        }
        if (KA == 0) {
            sp.d1.resize((size_t)1 << KB); sp.d2.assign(1, 0.25);
            for (auto& v : sp.d1) v = 1.0 + U(rng);
            c.d1 = sp.d1.data(); c.m1 = (1u << KB) - 1u; c.d2 = sp.d2.data(); c.m2 = 0; c.sh2 = 0;
        } else {
            sp.d1.resize((size_t)1 << KA); sp.d2.resize((size_t)1 << KB);
            for (auto& v : sp.d1) v = 0.5 + U(rng);
            for (auto& v : sp.d2) v = 0.5 + U(rng);
            c.d1 = sp.d1.data(); c.m1 = (1u << KA) - 1u; c.d2 = sp.d2.data(); c.m2 = (1u << KB) - 1u; c.sh2 = (uint32_t)KA;
        }
        blk_ctx_layout(c, K, KA > 8 ? KA : 8);
    }
    sp.rhs.assign((size_t)1 << K, 0.0);
    std::uniform_int_distribution<uint32_t> pick(0, (1u << K) - 1u);
    for (int i = 0; i < 40; ++i) sp.rhs[pick(rng)] = U(rng);
    sp.rhs[0] = 1.0; sp.rhs[(1u << K) - 1u] = 0.7;
    return sp;
}

static double rate(const Space& sp, int t, uint32_t u)
{
    const BlkBit& b = sp.ctx.bit[t];
    double r = 1.0;
    if (b.P) r *= b.P[u & b.mP];
    if (b.Q) r *= b.Q[(u >> b.shQ) & b.mQ];
    return r;
}
static double diag(const Space& sp, uint32_t u)
{
    const BlkCtx& c = sp.ctx;
    return c.d1[u & c.m1] + (c.d2 ? c.d2[(u >> c.sh2) & c.m2] : 0.0);
}

static std::vector<double> reference(const Space& sp, bool adj)
{
    const uint32_t N = 1u << sp.K;
    std::vector<double> v(N);
    if (!adj) {
        for (uint32_t s = 0; s < N; ++s) {
            double a = sp.rhs[s];
            for (int t = 0; t < sp.K; ++t) if ((s >> t) & 1u) a += rate(sp, t, s ^ (1u << t)) * v[s ^ (1u << t)];
            v[s] = a / diag(sp, s);
        }
    } else {
        for (uint32_t s = N; s-- > 0;) {
            double a = sp.rhs[s];
            for (int t = 0; t < sp.K; ++t) if (!((s >> t) & 1u)) a += rate(sp, t, s) * v[s | (1u << t)];
            v[s] = a / diag(sp, s);
        }
    }
    return v;
}

struct RhsArr {
    const double* b;
    void operator()(uint32_t s0, double (&acc)[8]) const { for (int j = 0; j < 8; ++j) acc[j] = b[s0 + blk_joff(j)]; }
};

// mirrors the orchestration of k_blk (mmh_device.cuh): ring of BLK_NS source rows filled in consumption order, OUTER phase
// with all lanes on the row of the iteration, INNER phase skewed
template <bool ADJ>
static std::vector<double> emulate(const Space& sp)
{
    const BlkCtx& c = sp.ctx;
    const uint32_t N = 1u << sp.K;
    std::vector<double> v(N, std::nan(""));                 // unsolved entries poison whatever reads them too early
    std::vector<double> sm(BLK_DOUBLES, 0.0), snap(BLK_DOUBLES), ring((size_t)BLK_NS * BLK_ROW);    // the device zeroes a warp's block once
    std::vector<double> ctab((size_t)BLK_MAXC * BLK_ROW, std::nan("")), sc(BLK_SC_DOUBLES);
    if (c.nC > BLK_MAXC) { std::printf("too many column profiles\n"); std::exit(2); }
    for (int t = 0; t < c.K; ++t)
        if (c.bit[t].cidx >= 0) for (int col = 0; col < BLK_ROW; ++col) blk_ctab_entry(c, t, col, ctab.data());
    RhsArr rhs{sp.rhs.data()};
    const int KO = c.KO;
    if (KO > BLK_MAXKO) { std::printf("too many outer bits\n"); std::exit(2); }
    BlkLane L[32];
    for (int lane = 0; lane < 32; ++lane) blk_lane_consts<ADJ>(c, lane, L[lane]);      // once per CTA on the device
    uint32_t cons = 0;                                      // running over the blocks of the "warp", like on the device
    for (int lv = 0; lv <= KO; ++lv) {
        const int level = ADJ ? KO - lv : lv;
        for (uint32_t o = 0; o < (1u << KO); ++o) {
            if (__builtin_popcount(o) != level) continue;
            BlkPlan plan;
            blk_plan(c, o, ADJ, plan);
            const uint32_t total = (uint32_t)BLK_Q * plan.nE, cons0 = cons;
            uint32_t issued = 0;
            auto issue = [&]() {
                const uint32_t src = blk_chunk_row<ADJ>(c, plan, issued / plan.nE, (int)(issued % plan.nE));
                std::memcpy(&ring[(size_t)((cons0 + issued) % BLK_NS) * BLK_ROW], &v[src], BLK_ROW * sizeof(double));
                ++issued;
            };
            while (issued < total && issued < (uint32_t)BLK_NS) issue();
            for (auto& x : sc) x = std::nan("");
            for (int idx = 0; idx < BLK_Q * (BLK_SC_OUT + plan.nE); ++idx) blk_sc_entry<ADJ>(c, plan, idx, sc.data());
            for (int lane = 0; lane < 32; ++lane) blk_lane_block(c, plan.base, lane, L[lane]);
            for (int t = 0; t < BLK_ITERS; ++t) {
                if (t < BLK_Q) {
                    const uint32_t q = ADJ ? BLK_Q - 1 - t : t;
                    double acc[32][8];
                    for (int lane = 0; lane < 32; ++lane) {
                        for (int j = 0; j < 8; ++j) acc[lane][j] = 0.0;
                        rhs(L[lane].base | blk_seq_mask(c, q), acc[lane]);
                    }
                    for (int k = 0; k < plan.nE; ++k) {
                        const double* slot = &ring[(size_t)(cons % BLK_NS) * BLK_ROW];
                        for (int lane = 0; lane < 32; ++lane)
                            blk_outer_edge(lane, sc[q * BLK_SCW + BLK_SC_OUT + k],
                                           plan.ecidx[k] >= 0 ? &ctab[(size_t)plan.ecidx[k] * BLK_ROW] : nullptr, slot, acc[lane]);
                        ++cons;
                        if (issued < total) issue();
                    }
                    for (int lane = 0; lane < 32; ++lane) blk_sts8(&sm[(size_t)q * BLK_ROW + lane * 2], acc[lane]);
                }
                snap = sm;
                for (int lane = 0; lane < 32; ++lane) {
                    if (c.simple) blk_inner<ADJ, true>(c, L[lane], lane, t, v.data(), snap.data(), sm.data(), sc.data(), ctab.data());
                    else blk_inner<ADJ, false>(c, L[lane], lane, t, v.data(), snap.data(), sm.data(), sc.data(), ctab.data());
                }
            }
        }
    }
    return v;
}

static double max_rel(const std::vector<double>& a, const std::vector<double>& b)
{
    double m = 0.0;
    for (size_t i = 0; i < a.size(); ++i) {
        const double e = std::fabs(a[i] - b[i]) / std::fmax(std::fabs(b[i]), 1e-300);
        if (!(e <= m)) m = std::isnan(e) ? 1e300 : e;
    }
    return m;
}

int main()
{
    std::mt19937_64 rng(12345);
    struct Case { int K, kind, KA; };
    const Case cases[] = {{13, 0, 0}, {15, 0, 0}, {13, 1, 9}, {14, 1, 8}, {15, 1, 11}, {14, 1, 12}, {13, 1, 3}, {14, 1, 5},
                          {13, 1, 0}, {14, 1, 14}, {16, 1, 10}, {13, 1, 13}, {16, 0, 0}, {13, 1, 1}, {14, 1, 2}, {14, 1, 6},
                          {15, 1, 7}, {17, 1, 9}};
    int bad = 0;
    for (const Case& cs : cases) {
        Space sp = make_space(cs.K, cs.kind, cs.KA, rng);
        const double ef = max_rel(emulate<false>(sp), reference(sp, false));
        const double ea = max_rel(emulate<true>(sp), reference(sp, true));
        std::printf("K=%d kind=%d KA=%d simple=%d seq=[%d %d %d %d] nC=%d d1row=%d d2mode=%d  fwd %.2e  adj %.2e\n", cs.K, cs.kind, cs.KA,
                    sp.ctx.simple, sp.ctx.seq[0], sp.ctx.seq[1], sp.ctx.seq[2], BLK_SB > 3 ? sp.ctx.seq[BLK_SB - 1] : -1, sp.ctx.nC, sp.ctx.d1row, sp.ctx.d2mode, ef, ea);
        if (!(ef < 1e-12) || !(ea < 1e-12)) ++bad;
    }
    std::printf(bad ? "FAILED %d\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
