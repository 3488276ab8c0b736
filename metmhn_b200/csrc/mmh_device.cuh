// Device-side data model and kernels of metmhn_b200 (sm_100a, FP64).
//
// Everything the reference does with reshape/transposing Kronecker shuffles
// (metmhn/jx/kronvec.py, vanilla.py) is expressed here directly on the subset lattice:
// a *space* is a lattice over K = KA + KB bits (group A = low bits, group B = high bits,
// canonical order: events shared by PT and MT first).  Transition rates come from small
// per-group tables (row = event, column = sub-state of the group), the diagonal of (D - Q)
// is separable over the two groups, the resolvent solve is an exact forward / backward
// substitution ordered by popcount, and the gradient is assembled from marginal statistics.
// See DESIGN.md section 3 for the derivation and oracle/lattice_direct.py for the executable
// specification these kernels follow.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mmh {

constexpr int NR      = 32;   // table rows per group (events 0..28, then the three special rows)
constexpr int ROW_D   = 29;   // full diagonal part of the group  (D_g + outflow)
constexpr int ROW_DP  = 30;   // PT-diagnosis rate table of the group (0 where it does not apply)
constexpr int ROW_DM  = 31;   // MT-diagnosis rate table
constexpr int MAXT    = 16;   // bits per group with one explicit table; wider groups use a product of two tables
constexpr int MAXG    = 26;   // bits per group
constexpr int BIGK    = 13;   // spaces with K >= BIGK are solved by per-level launches
constexpr int SEGB    = 32;   // blocks of 32 states per big-tier segment (one CTA)
constexpr int FIN_U   = 64;   // sub-states per finish work item
#ifndef FIN_U_WIDE
#define FIN_U_WIDE 256        // ... of a group with 2^15 sub-states or more (item.c carries the size)
#endif

enum Kind : uint8_t { K_PRE = 0, K_JOINT = 1, K_PF = 2, K_MF = 3, K_S1 = 4, K_S2 = 5 };

struct SpaceDev {
    uint8_t  kind, KA, KB, nb;
    uint8_t  has_pf, has_mf, cls, n_tot;
    int32_t  patient;
    int32_t  joint, pre, pf, mf;     // linked spaces (index into the chunk's space array), -1 = none
    uint32_t ptmask, mtmask;         // joint only: events present in PT / MT
    uint32_t slices;                 // joint only: partial-sum slices of the group-A statistics
    uint8_t  splitA, splitB;         // 0, or the number of low bits held by the first of two product tables
    uint8_t  pad[2];
    uint32_t slicesB, pad2;          // joint only: partial-sum slices of the group-B statistics
    uint64_t stPB;                   // joint only: group-B partials (slicesB x (KB+1) x NB)
    uint64_t y_off, x_off;           // scratch offsets (in doubles)
    uint64_t tabA, tabB;             // NR x NA and NR x NB rate / diagonal tables
    uint64_t stA, stB, stP;          // joint only: marginal statistics ((KA+1) x NA, (KB+1) x NB, partials)
    uint64_t tabR;                   // pairs of the row-block kernel: extra rate factors (mmh_rowblock.cuh)
    uint8_t  evA[MAXG], evB[MAXG];   // event id of every bit
};

struct EvalPar {                     // recomputed from the parameter vector at every evaluation
    double W[3][NR][NR];             // 0: exp(theta)  1: exp(theta_pt - d_p)  2: exp(theta - d_m)   (kronvec.py:7-21)
    double base[4][NR];              // 0: Th_ii  1: Th_ii Th_in  2: Th_ii Th_in / dm_n  3: Th_ii / dp_n
    double dp[NR], dm[NR];
};

struct Item { uint32_t space, a, b, c; };

// A group's tables.  Narrow group (KG <= MAXT): one NR x NG table whose rows 29..31 hold the diagonal part
// and the two diagnosis-rate tables.  Wide group: rate(i, u) = T1[i][u_lo] * T2[i][u_hi] (the rates are
// products over the set bits, so they factor over any split of the bits) and the three special rows are
// stored as full vectors behind the two tables.
struct Side {
    const double* t;
    int KG, K1;
    __device__ __forceinline__ uint32_t NG() const { return 1u << KG; }
    __device__ __forceinline__ double rate(int row, uint32_t u) const {
        if (K1 == 0) return t[((uint64_t)row << KG) + u];
        const int K2 = KG - K1;
        return t[((uint64_t)row << K1) + (u & ((1u << K1) - 1u))] *
               t[((uint64_t)NR << K1) + ((uint64_t)row << K2) + (u >> K1)];
    }
    __device__ __forceinline__ double special(int row, uint32_t u) const {
        if (K1 == 0) return t[((uint64_t)row << KG) + u];
        const int K2 = KG - K1;
        return t[((uint64_t)NR << K1) + ((uint64_t)NR << K2) + ((uint64_t)(row - ROW_D) << KG) + u];
    }
};
__device__ __forceinline__ Side side_of(const SpaceDev& sp, int g, const double* S)
{
    Side v;
    v.t = S + (g ? sp.tabB : sp.tabA);
    v.KG = g ? sp.KB : sp.KA;
    v.K1 = g ? sp.splitB : sp.splitA;
    return v;
}

// ------------------------------------------------------------------------------------------
__global__ void k_prep(const double* __restrict__ params, int n_tot, EvalPar* __restrict__ P)
{
    const int n = n_tot - 1;
    const double* th = params;
    const double* ldp = params + n_tot * n_tot;
    const double* ldm = ldp + n_tot;
    for (int t = threadIdx.x; t < NR * NR; t += blockDim.x) {
        int i = t / NR, j = t % NR;
        double w0 = 1.0, w1 = 1.0, w2 = 1.0;
        if (i < n_tot && j < n_tot) {
            double v = th[i * n_tot + j];
            double vpt = (j == n && i < n) ? 0.0 : v;       // likelihood.py:400 (seeding does not act on the PT)
            w0 = exp(v);
            w1 = exp(vpt - ldp[j]);
            w2 = exp(v - ldm[j]);
        }
        P->W[0][i][j] = w0; P->W[1][i][j] = w1; P->W[2][i][j] = w2;
    }
    for (int i = threadIdx.x; i < NR; i += blockDim.x) {
        double b0 = 0, b1 = 0, b2 = 0, b3 = 0, p = 1, m = 1;
        if (i < n_tot) {
            double d = th[i * n_tot + i];
            double tn = (i < n) ? th[i * n_tot + n] : 0.0;
            b0 = exp(d);
            b1 = exp(d + tn);
            b2 = exp(d + tn - ldm[n]);
            b3 = exp(d - ldp[n]);
            p = exp(ldp[i]); m = exp(ldm[i]);
        }
        P->base[0][i] = b0; P->base[1][i] = b1; P->base[2][i] = b2; P->base[3][i] = b3;
        P->dp[i] = p; P->dm[i] = m;
    }
}

// ------------------------------------------------------------------------------------------
// Group tables: for every sub-state u of a group and every event i
//   T[i][u] = base_i * prod_{b in u, ev(b) != i} W[i][ev(b)]          (rate of event i in sub-state u)
//   T[ROW_D][u] = D_g(u) + sum_{i not in u} T[i][u]                    (group part of diag(D - Q))
// which is what kron_diag / diag_scal_* / the k* factor products of kronvec.py:713-999 evaluate
// one shuffle pass at a time.  One CTA per (space, group, part); part 1 is the high-bit factor table of a
// wide group (no base rate, bits K1..KG-1).
struct SetupSel { int wid, bid, nrows, dg; };
__device__ __forceinline__ SetupSel setup_sel(const SpaceDev& sp, int g)
{
    const int n_tot = sp.n_tot, n = n_tot - 1;
    switch (sp.kind) {                        // dg: 0 one, 1 DP0, 2 DPA, 3 DMB, 4 S2 mixed
        case K_PRE:   return {0, 0, n_tot, 1};
        case K_JOINT: return {0, g ? 1 : 0, n, g ? 3 : 2};
        case K_PF:    return {2, 2, n, 0};
        case K_MF:    return {1, 3, n, 0};
        case K_S1:    return {1, 0, n_tot, 0};
        default:      return {0, 0, n_tot, 4};
    }
}

// D_g(u) and the two diagnosis tables of sub-state u (kronvec.py:574-602, 646-671; vanilla.py:125-142)
__device__ __forceinline__ void diag_rates(int dg, int KG, const uint8_t* sev, int n, const EvalPar* __restrict__ P,
                                           uint32_t u, double& d, double& vdp, double& vdm)
{
    d = 1.0; vdp = 0.0; vdm = 0.0;
    if (dg == 1 || dg == 2) {
        d = (dg == 2) ? P->dp[n] : 1.0;
        for (int b = 0; b < KG; ++b) if ((u >> b) & 1u) d *= P->dp[sev[b]];
        vdp = d;
    } else if (dg == 3) {
        d = P->dm[n];
        for (int b = 0; b < KG; ++b) if ((u >> b) & 1u) d *= P->dm[sev[b]];
        vdm = d;
    } else if (dg == 4) {                                   // seeding is the top bit
        const bool seeded = (u >> (KG - 1)) & 1u;
        for (int b = 0; b < KG; ++b) if ((u >> b) & 1u) d *= seeded ? P->dm[sev[b]] : P->dp[sev[b]];
        if (seeded) vdm = d; else vdp = d;
    }
}

__global__ void k_setup(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items,
                        const EvalPar* __restrict__ P, double* __restrict__ S)
{
    const Item it = items[blockIdx.x];                      // a = group | part << 1, b = first entry of this block
    const SpaceDev& sp = spaces[it.space];
    const int g = it.a & 1, part = it.a >> 1;
    const int KG = g ? sp.KB : sp.KA;
    const int K1 = g ? sp.splitB : sp.splitA;
    const int b0 = part ? K1 : 0;                           // first bit of this table
    const int KT = K1 ? (part ? KG - K1 : K1) : KG;         // bits of this table
    const uint32_t NT = 1u << KT;
    const uint8_t* ev = (g ? sp.evB : sp.evA) + b0;
    double* tab = S + (g ? sp.tabB : sp.tabA) + (part ? ((uint64_t)NR << K1) : 0);
    const int n = sp.n_tot - 1;
    const SetupSel sel = setup_sel(sp, g);
    __shared__ uint8_t sev[MAXG];
    if (threadIdx.x < MAXG) sev[threadIdx.x] = (int)threadIdx.x < KT ? ev[threadIdx.x] : 255;
    __syncthreads();
    const uint32_t u = it.b + threadIdx.x;
    if (u >= NT) return;
    double dsum = 0.0;
    for (int i = 0; i < sel.nrows; ++i) {
        double r = part ? 1.0 : P->base[sel.bid][i];
        bool in_u = false;
        for (uint32_t m = u; m; m &= m - 1) {                  // set bits only, in ascending order (same product as before)
            const int e = sev[__ffs(m) - 1];
            if (e == i) in_u = true; else r *= P->W[sel.wid][i][e];
        }
        tab[(uint64_t)i * NT + u] = r;
        if (!in_u) dsum += r;
    }
    if (K1 == 0) {
        double d, vdp, vdm;
        diag_rates(sel.dg, KG, sev, n, P, u, d, vdp, vdm);
        tab[(uint64_t)ROW_D * NT + u]  = d + dsum;
        tab[(uint64_t)ROW_DP * NT + u] = vdp;
        tab[(uint64_t)ROW_DM * NT + u] = vdm;
    } else {
        // product tables: unused rows are zero; for type 2 (dg == 4) rows 30 / 31 hold the two factors of the
        // PT / MT diagnosis rates (the seeding bit is the top bit of the high part and selects between them)
        for (int i = sel.nrows; i < NR; ++i) tab[(uint64_t)i * NT + u] = 0.0;
        if (sel.dg == 4) {
            double vp = 1.0, vm = 1.0;
            for (int b = 0; b < KT; ++b) if ((u >> b) & 1u) { vp *= P->dp[sev[b]]; vm *= P->dm[sev[b]]; }
            if (part) { const bool seeded = (u >> (KT - 1)) & 1u; if (seeded) vp = 0.0; else vm = 0.0; }
            tab[(uint64_t)ROW_DP * NT + u] = vp;
            tab[(uint64_t)ROW_DM * NT + u] = vm;
        }
    }
}

// special-row vectors of a wide group (after both factor tables exist); item.b = first sub-state of 1024
__global__ void k_setup_wide(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items,
                             const EvalPar* __restrict__ P, double* __restrict__ S)
{
    const Item it = items[blockIdx.x];
    const SpaceDev& sp = spaces[it.space];
    const int g = it.a;
    const Side sd = side_of(sp, g, S);
    const int KG = sd.KG, K1 = sd.K1, K2 = KG - K1;
    const uint8_t* ev = g ? sp.evB : sp.evA;
    const int n = sp.n_tot - 1;
    const SetupSel sel = setup_sel(sp, g);
    __shared__ uint8_t sev[MAXG];
    __shared__ int8_t bit_of[NR];
    if (threadIdx.x < NR) bit_of[threadIdx.x] = -1;
    __syncthreads();
    if (threadIdx.x < MAXG) {
        sev[threadIdx.x] = threadIdx.x < KG ? ev[threadIdx.x] : 255;
        if ((int)threadIdx.x < KG) bit_of[ev[threadIdx.x]] = (int8_t)threadIdx.x;
    }
    __syncthreads();
    double* vec = S + (g ? sp.tabB : sp.tabA) + ((uint64_t)NR << K1) + ((uint64_t)NR << K2);
    const uint32_t NG = 1u << KG;
    const uint32_t u = it.b + threadIdx.x;
    if (u >= NG) return;
    double dsum = 0.0;
    for (int i = 0; i < sel.nrows; ++i) {
        const int b = bit_of[i];
        if (b >= 0 && ((u >> b) & 1u)) continue;
        dsum += sd.rate(i, u);
    }
    double d, vdp, vdm;
    diag_rates(sel.dg, KG, sev, n, P, u, d, vdp, vdm);
    vec[u] = d + dsum;
    if (sel.dg != 0) {                                      // kinds without diagnosis tables keep only the diagonal
        vec[(uint64_t)NG + u] = vdp;
        vec[2ull * NG + u] = vdm;
    }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double rate_of(const SpaceDev& sp, const Side& A, const Side& B, int a, uint32_t s)
{
    const int KA = sp.KA;
    if (a < KA) return A.rate(sp.evA[a], s & ((1u << KA) - 1u));
    return B.rate(sp.evB[a - KA], s >> KA);
}

// right-hand side of the forward solve at state s
__device__ __forceinline__ double rhs_fwd(const SpaceDev& sp, const SpaceDev* __restrict__ spaces,
                                          const double* __restrict__ S, uint32_t s)
{
    switch (sp.kind) {
        case K_JOINT: {
            const uint32_t uA = s & ((1u << sp.KA) - 1u), uB = s >> sp.KA;
            if (uA != uB || uA >= (1u << sp.nb)) return 0.0;
            const SpaceDev& pre = spaces[sp.pre];            // seeding inflow from the pre-seeding lattice
            return side_of(pre, 0, S).rate(pre.n_tot - 1, uA) * S[pre.y_off + uA];
        }
        case K_PF: {                                         // PT observed first: D_P y on states with every PT bit set
            const SpaceDev& j = spaces[sp.joint];
            const uint32_t NAj = 1u << j.KA;
            const double cP = side_of(j, 0, S).special(ROW_DP, NAj - 1u);
            return cP * S[j.y_off + (((uint64_t)s << j.KA) | (NAj - 1u))];
        }
        case K_MF: {
            const SpaceDev& j = spaces[sp.joint];
            const uint32_t NBj = 1u << j.KB;
            const double cM = side_of(j, 1, S).special(ROW_DM, NBj - 1u);
            return cM * S[j.y_off + (((uint64_t)(NBj - 1u) << j.KA) | s)];
        }
        default: return s == 0u ? 1.0 : 0.0;
    }
}

__device__ __forceinline__ double joint_score(const SpaceDev& j, const SpaceDev* __restrict__ spaces,
                                              const double* __restrict__ S)
{
    double sc = 0.0;
    if (j.has_pf) { const SpaceDev& q = spaces[j.pf]; sc += S[q.y_off + ((1u << q.KA) - 1u)]; }
    if (j.has_mf) { const SpaceDev& q = spaces[j.mf]; sc += S[q.y_off + ((1u << q.KA) - 1u)]; }
    return sc;
}

// right-hand side of the adjoint solve at state s
__device__ __forceinline__ double rhs_adj(const SpaceDev& sp, const SpaceDev* __restrict__ spaces,
                                          const double* __restrict__ S, uint32_t s)
{
    const uint32_t N = 1u << (sp.KA + sp.KB);
    switch (sp.kind) {
        case K_S1: case K_S2:
            return s == N - 1u ? 1.0 / S[sp.y_off + N - 1u] : 0.0;
        case K_PF: case K_MF:
            return s == N - 1u ? 1.0 / joint_score(spaces[sp.joint], spaces, S) : 0.0;
        case K_JOINT: {
            const uint32_t NA = 1u << sp.KA, NB = 1u << sp.KB;
            const uint32_t uA = s & (NA - 1u), uB = s >> sp.KA;
            double q = 0.0;
            if (sp.has_pf && uA == NA - 1u)
                q += side_of(sp, 0, S).special(ROW_DP, uA) * S[spaces[sp.pf].x_off + uB];
            if (sp.has_mf && uB == NB - 1u)
                q += side_of(sp, 1, S).special(ROW_DM, uB) * S[spaces[sp.mf].x_off + uA];
            return q;
        }
        default: {                                           // K_PRE: seeding edge into the joint lattice
            const SpaceDev& j = spaces[sp.joint];
            return side_of(sp, 0, S).rate(sp.n_tot - 1, s) * S[j.x_off + (((uint64_t)s << j.KA) | s)];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Resolvent solve.  Per-space addressing is resolved ONCE per CTA (big tier) / warp (small tier) into a
// shared-memory context: for every bit the two table-row pointers whose product is the rate of adding
// that bit, and the pointers of the diagonal parts.  The per-state work is then a handful of loads and
// FMAs (the first version of this kernel was issue-bound on address arithmetic, profiles/r1_v3*).
__device__ const double c_one = 1.0;
__device__ const double c_zero = 0.0;

struct BitDesc {                       // rate(bit, p) = p1[(p >> sh1) & m1] * p2[(p >> sh2) & m2]
    const double* p1;
    const double* p2;
    uint32_t m1, m2;
    uint32_t sh1, sh2;
};
struct SpaceCtx {
    BitDesc bit[MAXG];
    const double* dA;                  // diag(s) = dA[s & mA] + dB[(s >> KA) & mB]
    const double* dB;
    uint32_t mA, mB;
    int K, KA;
};

__device__ __forceinline__ void ctx_build(SpaceCtx& c, const SpaceDev& sp, const double* __restrict__ S, int t)
{
    const int KA = sp.KA, KB = sp.KB, K = KA + KB;
    if (t < K) {
        const int g = t >= KA, b = g ? t - KA : t;
        const int KG = g ? KB : KA, K1 = g ? sp.splitB : sp.splitA;
        const uint32_t sh = g ? KA : 0;
        const int ev = g ? sp.evB[b] : sp.evA[b];
        const double* base = S + (g ? sp.tabB : sp.tabA);
        BitDesc d;
        if (K1 == 0) {
            d.p1 = base + ((uint64_t)ev << KG); d.sh1 = sh; d.m1 = (1u << KG) - 1u;
            d.p2 = &c_one; d.sh2 = 0; d.m2 = 0;
        } else {
            const int K2 = KG - K1;
            d.p1 = base + ((uint64_t)ev << K1); d.sh1 = sh; d.m1 = (1u << K1) - 1u;
            d.p2 = base + ((uint64_t)NR << K1) + ((uint64_t)ev << K2); d.sh2 = sh + K1; d.m2 = (1u << K2) - 1u;
        }
        c.bit[t] = d;
    }
    if (t == 0) {
        auto drow = [&](int g) -> const double* {
            const int KG = g ? KB : KA, K1 = g ? sp.splitB : sp.splitA;
            const double* base = S + (g ? sp.tabB : sp.tabA);
            return K1 == 0 ? base + ((uint64_t)ROW_D << KG) : base + ((uint64_t)NR << K1) + ((uint64_t)NR << (KG - K1));
        };
        c.dA = drow(0); c.mA = (1u << KA) - 1u;
        if (sp.kind == K_JOINT) { c.dB = drow(1); c.mB = (1u << KB) - 1u; }
        else { c.dB = &c_zero; c.mB = 0; }
        c.K = K; c.KA = KA;
    }
}

__device__ __forceinline__ double ctx_rate(const SpaceCtx& c, int a, uint32_t p)
{
    const BitDesc& d = c.bit[a];
    return d.p1[(p >> d.sh1) & d.m1] * d.p2[(p >> d.sh2) & d.m2];
}

// One block of 32 consecutive states (lane = low 5 bits).  Edges that add a bit >= 5 read values of blocks
// finished earlier (same lane, coalesced); the 5 lane bits are resolved inside the warp by popcount
// sub-levels with shuffles.  Replaces the (k+1)-sweep Jacobi iteration of likelihood.py:231-262 /
// vanilla.py:269-305 by one exact substitution.
template <bool ADJ>
__device__ __forceinline__ void solve_block(const SpaceDev& sp, const SpaceDev* __restrict__ spaces, const SpaceCtx& c,
                                            double* __restrict__ S, uint32_t hi, int lane)
{
    const int K = c.K;
    const uint32_t s = (hi << 5) | (uint32_t)lane;
    const bool valid = K >= 5 || (uint32_t)lane < (1u << K);
    double* v = S + (ADJ ? sp.x_off : sp.y_off);
    const int nl = K < 5 ? K : 5;
    double acc = 0.0, inv = 0.0;
    double rl[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (valid) {
        inv = 1.0 / (c.dA[s & c.mA] + c.dB[(s >> c.KA) & c.mB]);
        acc = ADJ ? rhs_adj(sp, spaces, S, s) : rhs_fwd(sp, spaces, S, s);
        // bits >= 5: FWD visits the set bits of hi, ADJ the unset ones
        uint32_t m = ADJ ? (~hi & ((K > 5 ? (1u << (K - 5)) : 1u) - 1u)) : hi;
        while (m) {
            const int a = __ffs(m) + 4;
            m &= m - 1;
            const uint32_t bit = 1u << a;
            if (!ADJ) { const uint32_t p = s ^ bit; acc = fma(ctx_rate(c, a, p), v[p], acc); }
            else      { acc = fma(ctx_rate(c, a, s), v[s | bit], acc); }
        }
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            if (a < nl) {
                const uint32_t bit = 1u << a;
                if (!ADJ) { if (s & bit) rl[a] = ctx_rate(c, a, s ^ bit); }
                else      { if (!(s & bit)) rl[a] = ctx_rate(c, a, s); }
            }
        }
    }
    // lane bits: at sub-level l the lanes with popcount l publish their final value; every lane adds what its
    // neighbours publish (rl is zero for non-edges), so each edge is used exactly once.
    const int pl = __popc(lane);
    double val = 0.0;
    if (!ADJ) {
        for (int l = 0; l <= nl; ++l) {
            if (pl == l) val = acc * inv;
            if (l < nl) {
                const double pub = (pl == l) ? val : 0.0;
#pragma unroll
                for (int a = 0; a < 5; ++a)
                    if (a < nl) acc = fma(rl[a], __shfl_xor_sync(0xffffffffu, pub, 1 << a), acc);
            }
        }
    } else {
        for (int l = nl; l >= 0; --l) {
            if (pl == l) val = acc * inv;
            if (l > 0) {
                const double pub = (pl == l) ? val : 0.0;
#pragma unroll
                for (int a = 0; a < 5; ++a)
                    if (a < nl) acc = fma(rl[a], __shfl_xor_sync(0xffffffffu, pub, 1 << a), acc);
            }
        }
    }
    if (valid) v[s] = val;
}

// small tier: one warp owns a whole space and walks its blocks in index order (a valid
// topological order of the lattice: every predecessor has a smaller index).
template <bool ADJ>
__global__ void __launch_bounds__(256)
k_solve_small(const SpaceDev* __restrict__ spaces, const uint32_t* __restrict__ list,
              uint32_t count, double* __restrict__ S)
{
    __shared__ SpaceCtx ctx[8];
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    if (w >= count) return;
    const SpaceDev& sp = spaces[list[w]];
    ctx_build(ctx[wl], sp, S, lane);
    __syncwarp();
    const int K = sp.KA + sp.KB;
    const uint32_t nblk = K > 5 ? (1u << (K - 5)) : 1u;
    if (!ADJ) {
        for (uint32_t hi = 0; hi < nblk; ++hi) { solve_block<false>(sp, spaces, ctx[wl], S, hi, lane); __syncwarp(); }
    } else {
        for (uint32_t hi = nblk; hi-- > 0;) { solve_block<true>(sp, spaces, ctx[wl], S, hi, lane); __syncwarp(); }
    }
}

// ------------------------------------------------------------------------------------------
// Four states per lane.  A warp owns 128 consecutive states: bits 0,1 live inside the lane (registers),
// bits 2..6 across the lanes, bits >= 7 select the block.  The index arithmetic of an edge bit is done once
// per four states and the table rows / vectors are read with 16-byte loads (still fully coalesced: a warp reads
// 1 KB of consecutive addresses).  Needs K >= 7; smaller spaces use solve_block.
__device__ __forceinline__ void factor4(const double* __restrict__ p, uint32_t sh, uint32_t m, uint32_t p0, double (&f)[4])
{
    if (sh >= 2 || m == 0) {                                   // constant over the four states
        const double c = p[(p0 >> sh) & m];
        f[0] = c; f[1] = c; f[2] = c; f[3] = c;
    } else if (sh == 0 && m >= 3u) {                           // contiguous
        const double2* q = reinterpret_cast<const double2*>(p + (p0 & m));
        const double2 a = q[0], b = q[1];
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) f[t] = p[((p0 + t) >> sh) & m];
    }
}
__device__ __forceinline__ void rate4(const BitDesc& d, uint32_t p0, double (&r)[4])
{
    factor4(d.p1, d.sh1, d.m1, p0, r);
    if (d.p2 != &c_one) {
        double g[4];
        factor4(d.p2, d.sh2, d.m2, p0, g);
#pragma unroll
        for (int t = 0; t < 4; ++t) r[t] *= g[t];
    }
}

// MODE 1: pair with two plain tables and KA >= 7 -> bits below KA read four consecutive table entries, bits of
//         group B one scalar.  MODE 2: product tables (single group): four consecutive entries of the low-part
//         table times one scalar of the high-part table.  MODE 0: anything else (generic indexing).
// The modes are separate instantiations selected once per CTA, which keeps the index-class branches of factor4
// (otherwise if-converted into one long predicated sequence, profiles/r1_v6*) out of the hot loop.
__device__ __forceinline__ int solve_mode(const SpaceDev& sp)
{
    if (sp.kind == K_JOINT) return (!sp.splitA && !sp.splitB && sp.KA >= 7) ? 1 : 0;
    return sp.splitA ? 2 : 0;
}

template <int MODE>
__device__ __forceinline__ void rate4m(const SpaceCtx& c, int a, uint32_t p0, double (&r)[4])
{
    const BitDesc& d = c.bit[a];
    if (MODE == 0) { rate4(d, p0, r); return; }
    if (MODE == 1 && a >= c.KA) {
        const double k = d.p1[p0 >> c.KA];
        r[0] = k; r[1] = k; r[2] = k; r[3] = k;
        return;
    }
    const double2* q = reinterpret_cast<const double2*>(d.p1 + (p0 & d.m1));
    const double2 u = q[0], w = q[1];
    if (MODE == 2) {
        const double k = d.p2[(p0 >> d.sh2) & d.m2];
        r[0] = u.x * k; r[1] = u.y * k; r[2] = w.x * k; r[3] = w.y * k;
    } else { r[0] = u.x; r[1] = u.y; r[2] = w.x; r[3] = w.y; }
}

template <bool ADJ, int MODE>
__device__ __forceinline__ void solve_block4(const SpaceDev& sp, const SpaceDev* __restrict__ spaces, const SpaceCtx& c,
                                             double* __restrict__ S, uint32_t hi, int lane)
{
    const int K = c.K;
    const uint32_t s0 = (hi << 7) | ((uint32_t)lane << 2);
    double* v = S + (ADJ ? sp.x_off : sp.y_off);
    double acc[4], inv[4], r[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t] = ADJ ? rhs_adj(sp, spaces, S, s0 + t) : rhs_fwd(sp, spaces, S, s0 + t);
    // bits >= 7: FWD visits the set bits of hi, ADJ the unset ones; BATCH bits per round, loads before FMAs
    uint32_t m = ADJ ? (~hi & ((1u << (K - 7)) - 1u)) : hi;
    constexpr int BATCH = 2;
    while (m) {
        double rr[BATCH][4];
        double2 va[BATCH], vb[BATCH];
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            // an exhausted slot re-reads the block's own (finished or not) location with rate 0
            const bool on = m != 0;
            const int a = on ? __ffs(m) + 6 : 7;
            m &= m - 1;
            const uint32_t bit = 1u << a;
            const uint32_t p0 = ADJ ? s0 : (s0 ^ bit);
            const double2* qv = reinterpret_cast<const double2*>(v + (on ? (ADJ ? (s0 | bit) : p0) : s0));
            rate4m<MODE>(c, a, on ? p0 : s0, rr[q]);
            va[q] = qv[0]; vb[q] = qv[1];
            if (!on) { rr[q][0] = 0.0; rr[q][1] = 0.0; rr[q][2] = 0.0; rr[q][3] = 0.0; va[q] = make_double2(0.0, 0.0); vb[q] = va[q]; }
        }
#pragma unroll
        for (int q = 0; q < BATCH; ++q) {
            acc[0] = fma(rr[q][0], va[q].x, acc[0]); acc[1] = fma(rr[q][1], va[q].y, acc[1]);
            acc[2] = fma(rr[q][2], vb[q].x, acc[2]); acc[3] = fma(rr[q][3], vb[q].y, acc[3]);
        }
    }
    // diagonal
    {
        double dA[4], dB[4];
        if (MODE == 0) {
            factor4(c.dA, 0, c.mA, s0, dA);
            factor4(c.dB, (uint32_t)c.KA, c.mB, s0, dB);
        } else {
            const double2* q = reinterpret_cast<const double2*>(c.dA + (s0 & c.mA));
            const double2 u = q[0], w = q[1];
            dA[0] = u.x; dA[1] = u.y; dA[2] = w.x; dA[3] = w.y;
            const double k = MODE == 1 ? c.dB[s0 >> c.KA] : 0.0;
            dB[0] = k; dB[1] = k; dB[2] = k; dB[3] = k;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) inv[t] = 1.0 / (dA[t] + dB[t]);
    }
    // edges inside the lane: bit 0 from t = 0 and t = 2, bit 1 from t = 0 and t = 1
    rate4m<MODE>(c, 0, s0, r);
    const double e0_0 = r[0], e0_2 = r[2];
    rate4m<MODE>(c, 1, s0, r);
    const double e1_0 = r[0], e1_1 = r[1];
    // edges across lanes (bits 2..6): rate at the state that lacks the bit
    double rl[5][4];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const uint32_t bit = 4u << a;
        const bool has = (s0 & bit) != 0;
        if (ADJ ? !has : has) rate4m<MODE>(c, a + 2, ADJ ? s0 : (s0 ^ bit), rl[a]);
        else { rl[a][0] = 0.0; rl[a][1] = 0.0; rl[a][2] = 0.0; rl[a][3] = 0.0; }
    }
    const int pl = __popc(lane);
    double val[4] = {0.0, 0.0, 0.0, 0.0};
    if (!ADJ) {
        for (int l = 0; l <= 5; ++l) {
            const bool mine = pl == l;
            if (mine) {
                val[0] = acc[0] * inv[0];
                val[1] = fma(e0_0, val[0], acc[1]) * inv[1];
                val[2] = fma(e1_0, val[0], acc[2]) * inv[2];
                val[3] = fma(e0_2, val[2], fma(e1_1, val[1], acc[3])) * inv[3];
            }
            if (l < 5) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double pub = mine ? val[t] : 0.0;
#pragma unroll
                    for (int a = 0; a < 5; ++a) acc[t] = fma(rl[a][t], __shfl_xor_sync(0xffffffffu, pub, 1 << a), acc[t]);
                }
            }
        }
    } else {
        for (int l = 5; l >= 0; --l) {
            const bool mine = pl == l;
            if (mine) {
                val[3] = acc[3] * inv[3];
                val[2] = fma(e0_2, val[3], acc[2]) * inv[2];
                val[1] = fma(e1_1, val[3], acc[1]) * inv[1];
                val[0] = fma(e0_0, val[1], fma(e1_0, val[2], acc[0])) * inv[0];
            }
            if (l > 0) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double pub = mine ? val[t] : 0.0;
#pragma unroll
                    for (int a = 0; a < 5; ++a) acc[t] = fma(rl[a][t], __shfl_xor_sync(0xffffffffu, pub, 1 << a), acc[t]);
                }
            }
        }
    }
    double2* o = reinterpret_cast<double2*>(v + s0);
    o[0] = make_double2(val[0], val[1]);
    o[1] = make_double2(val[2], val[3]);
}

template <bool ADJ>
__global__ void __launch_bounds__(256)
k_solve_small4(const SpaceDev* __restrict__ spaces, const uint32_t* __restrict__ list, uint32_t count, double* __restrict__ S)
{
    __shared__ SpaceCtx ctx[8];
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    if (w >= count) return;
    const SpaceDev& sp = spaces[list[w]];
    ctx_build(ctx[wl], sp, S, lane);
    __syncwarp();
    const uint32_t nblk = 1u << (sp.KA + sp.KB - 7);
    const int mode = solve_mode(sp);
    if (!ADJ) {
        for (uint32_t hi = 0; hi < nblk; ++hi) {
            if (mode == 1) solve_block4<false, 1>(sp, spaces, ctx[wl], S, hi, lane);
            else if (mode == 2) solve_block4<false, 2>(sp, spaces, ctx[wl], S, hi, lane);
            else solve_block4<false, 0>(sp, spaces, ctx[wl], S, hi, lane);
            __syncwarp();
        }
    } else {
        for (uint32_t hi = nblk; hi-- > 0;) {
            if (mode == 1) solve_block4<true, 1>(sp, spaces, ctx[wl], S, hi, lane);
            else if (mode == 2) solve_block4<true, 2>(sp, spaces, ctx[wl], S, hi, lane);
            else solve_block4<true, 0>(sp, spaces, ctx[wl], S, hi, lane);
            __syncwarp();
        }
    }
}

template <bool ADJ>
__global__ void __launch_bounds__(256)
k_solve_big4(const SpaceDev* __restrict__ spaces, const Item* __restrict__ segs, const uint32_t* __restrict__ hs, double* __restrict__ S)
{
    __shared__ SpaceCtx ctx;
    const Item sg = segs[blockIdx.x];
    const SpaceDev& sp = spaces[sg.space];
    ctx_build(ctx, sp, S, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int mode = solve_mode(sp);
    if (mode == 1)      for (uint32_t r = w; r < sg.b; r += nw) solve_block4<ADJ, 1>(sp, spaces, ctx, S, hs[sg.a + r], lane);
    else if (mode == 2) for (uint32_t r = w; r < sg.b; r += nw) solve_block4<ADJ, 2>(sp, spaces, ctx, S, hs[sg.a + r], lane);
    else                for (uint32_t r = w; r < sg.b; r += nw) solve_block4<ADJ, 0>(sp, spaces, ctx, S, hs[sg.a + r], lane);
}

// ------------------------------------------------------------------------------------------
// Tiled solve (big tier, K >= BIGK).  The lattice is viewed as rows x columns:
//   pair            : columns = group A (KC = KA bits), rows = group B (KR = KB bits); an A-edge's rate is a column
//                     vector T_A[ev][uA], a B-edge's rate one scalar T_B[ev][uB] per row
//   product single  : columns = the K1 low bits, rows = the K2 high bits; every rate is a column factor T1[ev][lo]
//                     times a row factor T2[ev][hi]
// A warp solves 8 independent rows x 16 consecutive columns: lane = (row group, 4 columns); column bits 0,1 live in
// the lane's registers, bits 2,3 in the four lanes of a row group (three sub-levels, 3 shuffles per state), all
// higher bits read 128-byte lines of blocks finished by earlier launches.  The 8 rows of a tile have the same
// popcount lB and share the column block cA (popcount lA), so every loop of the warp has a uniform trip count and
// the column-rate loads of the 8 row groups hit the same line.  One launch per level lA + lB.
constexpr int TILES_PER_CTA = 32;
#ifndef TILE_CTAS
#define TILE_CTAS 4            // resident CTAs per SM of k_solve_tile (64 registers; 3 measured 1.2 % slower)
#endif
#ifndef ADJB_CTAS
#define ADJB_CTAS 3
#endif
#ifndef TILE_NBA
#define TILE_NBA 2            // column-bit edges in flight per round
#endif
#ifndef TILE_NBB
#define TILE_NBB 3            // row-bit edges in flight per round
#endif

struct TileCtx {
    const double* colA[MAXG];          // column bit q: column factor of its rate (T_A[ev] or T1[ev])
    const double* rowA[MAXG];          // product only: row factor T2[ev] of column bit q
    const double* rowB[MAXG];          // row bit b: row factor (T_B[ev] or T2[ev])
    const double* colB[MAXG];          // product only: column factor T1[ev] of row bit b
    const double* dA;                  // pair: A part of the diagonal; product: the full diagonal vector
    const double* dB;                  // pair: B part of the diagonal
    int KC, KR;
};
// wide pair (TM_WIDE): the row index is (uB << K2A | high part of uA); which part of it a factor table is indexed by
struct TileCtxW : TileCtx {
    uint32_t mH;                       // mask of the high part of uA inside the row index
    uint32_t mAfull;                   // 2^KA - 1
    int K2A;
    uint8_t rsh[MAXG];                 // row bit b: index of rowB[b] = (row >> rsh[b]) & rmask[b]
    uint32_t rmask[MAXG];
    uint8_t cfone[MAXG];               // row bit b has no column factor (an MT event: scalar rate per row)
};

// Table modes of the tile kernels.  TM_PAIR: a pair whose groups have at most MAXT bits (one plain table per group, an A
// edge's rate is a column vector, a B edge's rate one scalar per row).  TM_PROD: a product-form single-tumour space
// (rate = column factor T1[ev][lo] x row factor T2[ev][hi], full diagonal vector).  TM_WIDE: a pair whose PT group has more
// than MAXT bits, e.g. 18 PT events against 5 MT events: its A rates are stored in product form too, so the lattice is cut
// like a product space -- columns = the low K1A bits of uA, rows = (uB, high part of uA) -- which gives every level enough
// rows to fill the 8-row tiles however few MT events there are; MT edges are row edges with a scalar rate and no column
// factor, the diagonal is dA[uA] (full vector) + dB[uB].  Before, these pairs ran on the generic kernel: 60 such pairs were
// 31 % of the n = 20 / 10 000 step, 28 were 7 % of the n = 25 / 100 000 step (scripts/marginal_generic.py).
constexpr int TM_PAIR = 0, TM_PROD = 1, TM_WIDE = 2;


__device__ __forceinline__ void tile_ctx_build(TileCtx& c, const SpaceDev& sp, const double* __restrict__ S, int t)
{
    if (sp.kind == K_JOINT) {
        const int KA = sp.KA, KB = sp.KB;
        if (t < KA) { c.colA[t] = S + sp.tabA + ((uint64_t)sp.evA[t] << KA); c.rowA[t] = &c_one; }
        if (t < KB) { c.rowB[t] = S + sp.tabB + ((uint64_t)sp.evB[t] << KB); c.colB[t] = &c_one; }
        if (t == 0) {
            c.dA = S + sp.tabA + ((uint64_t)ROW_D << KA);
            c.dB = S + sp.tabB + ((uint64_t)ROW_D << KB);
            c.KC = KA; c.KR = KB;
        }
    } else {
        const int K1 = sp.splitA, K2 = sp.KA - K1;
        const double* T1 = S + sp.tabA;
        const double* T2 = T1 + ((uint64_t)NR << K1);
        if (t < K1) { c.colA[t] = T1 + ((uint64_t)sp.evA[t] << K1); c.rowA[t] = T2 + ((uint64_t)sp.evA[t] << K2); }
        if (t < K2) { c.rowB[t] = T2 + ((uint64_t)sp.evA[K1 + t] << K2); c.colB[t] = T1 + ((uint64_t)sp.evA[K1 + t] << K1); }
        if (t == 0) { c.dA = T2 + ((uint64_t)NR << K2); c.dB = &c_zero; c.KC = K1; c.KR = K2; }
    }
}

__device__ __forceinline__ void tile_ctx_build_wide(TileCtxW& c, const SpaceDev& sp, const double* __restrict__ S, int t)
{
    {
        const int KA = sp.KA, KB = sp.KB, K1 = sp.splitA, K2 = KA - K1;
        const double* T1 = S + sp.tabA;
        const double* T2 = T1 + ((uint64_t)NR << K1);
        if (t < K1) { c.colA[t] = T1 + ((uint64_t)sp.evA[t] << K1); c.rowA[t] = T2 + ((uint64_t)sp.evA[t] << K2); }
        if (t < K2) {
            c.rowB[t] = T2 + ((uint64_t)sp.evA[K1 + t] << K2); c.colB[t] = T1 + ((uint64_t)sp.evA[K1 + t] << K1);
            c.rsh[t] = 0; c.rmask[t] = (1u << K2) - 1u; c.cfone[t] = 0;
        }
        if (t < KB) {
            c.rowB[K2 + t] = S + sp.tabB + ((uint64_t)sp.evB[t] << KB); c.colB[K2 + t] = &c_one;
            c.rsh[K2 + t] = (uint8_t)K2; c.rmask[K2 + t] = (1u << KB) - 1u; c.cfone[K2 + t] = 1;
        }
        if (t == 0) {
            c.dA = T2 + ((uint64_t)NR << K2);                 // full vector of the A part of the diagonal (Side::special)
            c.dB = S + sp.tabB + ((uint64_t)ROW_D << KB);
            c.KC = K1; c.KR = KB + K2; c.K2A = K2; c.mH = (1u << K2) - 1u; c.mAfull = (1u << KA) - 1u;
        }
    }
}

// 32-byte global load / store (LDG.E.256 / STG.E.256 on sm_100a): four lanes cover one 128-byte line with a single
// request, which halves the L1 wavefronts of the tile kernels against two 16-byte loads per lane.
__device__ __forceinline__ void ld4(const double* __restrict__ p, double (&f)[4])
{
    asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(f[0]), "=d"(f[1]), "=d"(f[2]), "=d"(f[3]) : "l"(p));
}
__device__ __forceinline__ void st4(double* __restrict__ p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" :: "d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
}

// right-hand side of the four states of a lane: non-zero on few states only (except the second phase's start vector).
// WIDEJ (TM_WIDE): the tile's columns are only the low part of group A, (uB, uA) come from the state index with jKA / jKB.
template <bool ADJ, bool WIDEJ = false>
__device__ __forceinline__ void tile_rhs(const SpaceDev& sp, const SpaceDev* __restrict__ spaces, const double* __restrict__ S,
                                         int KC, int KR, uint32_t row, uint32_t lo0, double (&acc)[4], int jKA = 0, int jKB = 0)
{
    const uint32_t s0 = (row << KC) | lo0;
    {
        const uint32_t NC = 1u << KC, NRW = 1u << KR;
        bool any;
        if (WIDEJ) {
            const uint32_t NA = 1u << jKA, NB = 1u << jKB;
            const uint32_t uA = s0 & (NA - 1u), uB = s0 >> jKA;
            if (!ADJ) any = uB < (1u << sp.nb) && ((uB ^ uA) & ~3u) == 0u;
            else any = (sp.has_pf && (uA | 3u) == NA - 1u) || (sp.has_mf && uB == NB - 1u);
        } else if (!ADJ) {
            if (sp.kind == K_JOINT) any = row < (1u << sp.nb) && ((row ^ lo0) & ~3u) == 0u;
            else if (sp.kind == K_PF || sp.kind == K_MF) any = true;
            else any = s0 == 0u;
        } else {
            if (sp.kind == K_JOINT) any = (sp.has_pf && (lo0 | 3u) == NC - 1u) || (sp.has_mf && row == NRW - 1u);
            else any = (s0 | 3u) == (NC << KR) - 1u;
        }
        if (any) {
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t] = ADJ ? rhs_adj(sp, spaces, S, s0 + t) : rhs_fwd(sp, spaces, S, s0 + t);
        }
    }
}

// edges on the row bits of a lane's four states (sources in global memory): acc += rate * v[other row]
template <bool ADJ, int MODE, class CTX>
__device__ __forceinline__ void tile_row_edges(const CTX& c, const double* __restrict__ v, uint32_t row, uint32_t lo0,
                                               double (&acc)[4])
{
    const int KC = c.KC, KR = c.KR;
    {
        constexpr int NB = TILE_NBB;
        uint32_t m = ADJ ? (~row & ((1u << KR) - 1u)) : row;
        while (m) {
            double y[NB][4], k[NB];
            int bq[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const bool on = m != 0u;
                const int b = on ? __ffs(m) - 1 : 0;
                m &= m - 1;
                bq[q] = b;
                const uint32_t orow = row ^ (1u << b);
                if (on) {
                    const uint32_t rr = ADJ ? row : orow;
                    if constexpr (MODE == TM_WIDE) k[q] = c.rowB[b][(rr >> c.rsh[b]) & c.rmask[b]];
                    else k[q] = c.rowB[b][rr];
                    ld4(v + (((uint64_t)orow << KC) | lo0), y[q]);
                } else {
                    k[q] = 0.0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) y[q][t] = 0.0;
                }
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                bool has_cf = MODE == TM_PROD;
                if constexpr (MODE == TM_WIDE) has_cf = !c.cfone[bq[q]];
                if (has_cf) {
                    double cf[4];
                    ld4(c.colB[bq[q]] + lo0, cf);
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[t] = fma(cf[t] * k[q], y[q][t], acc[t]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[t] = fma(k[q], y[q][t], acc[t]);
                }
            }
        }
    }
}

// diagonal, column bits 0,1 (inside the lane) and 2,3 (across the four lanes of a row group): acc -> val
template <bool ADJ, int MODE, class CTX>
__device__ __forceinline__ void tile_tail(const CTX& c, uint32_t row, uint32_t lo0, int lane, double (&acc)[4], double (&val)[4])
{
    const int KC = c.KC;
    const int lc = lane & 3;
    const uint32_t s0 = (row << KC) | lo0;
    // ---- diagonal ----
    double inv[4];
    {
        double d[4];
        if (MODE == TM_PROD) ld4(c.dA + s0, d);
        else {
            uint32_t ia = lo0, ib = row;
            if constexpr (MODE == TM_WIDE) { ia = s0 & c.mAfull; ib = row >> c.K2A; }
            ld4(c.dA + ia, d);
            const double k = c.dB[ib];
#pragma unroll
            for (int t = 0; t < 4; ++t) d[t] += k;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) inv[t] = 1.0 / d[t];
    }
    // ---- column bits 0,1 (inside the lane) and 2,3 (across the four lanes of the row group) ----
    double k0 = 1.0, k1 = 1.0, k2 = 1.0, k3 = 1.0;
    if (MODE != TM_PAIR) {
        uint32_t ri = row;
        if constexpr (MODE == TM_WIDE) ri = row & c.mH;
        k0 = c.rowA[0][ri]; k1 = c.rowA[1][ri]; k2 = c.rowA[2][ri]; k3 = c.rowA[3][ri];
    }
    double e0a, e0b, e1a, e1b;
    {
        double q0[4], q1[4];
        ld4(c.colA[0] + lo0, q0);
        ld4(c.colA[1] + lo0, q1);
        e0a = q0[0] * k0; e0b = q0[2] * k0;          // bit 0: 0 -> 1, 2 -> 3
        e1a = q1[0] * k1; e1b = q1[1] * k1;          // bit 1: 0 -> 2, 1 -> 3
    }
    double w0[4] = {0.0, 0.0, 0.0, 0.0}, w1a[4] = {0.0, 0.0, 0.0, 0.0}, w1b[4] = {0.0, 0.0, 0.0, 0.0};
    const int pl = __popc(lc);
    if (!ADJ) {
        // round 0: lanes 1, 2 receive from lane 0;  round 1: lane 3 receives from lanes 2 (bit 2) and 1 (bit 3)
        if (pl == 1) {
            const int q = lc == 2;
            ld4(c.colA[2 + q] + (lo0 ^ (4u << q)), w0);
            const double k = q ? k3 : k2;
#pragma unroll
            for (int t = 0; t < 4; ++t) w0[t] *= k;
        } else if (lc == 3) {
            ld4(c.colA[2] + (lo0 ^ 4u), w1a);
            ld4(c.colA[3] + (lo0 ^ 8u), w1b);
#pragma unroll
            for (int t = 0; t < 4; ++t) { w1a[t] *= k2; w1b[t] *= k3; }
        }
    } else {
        // round 0: lanes 1, 2 receive from lane 3;  round 1: lane 0 receives from lanes 1 (bit 2) and 2 (bit 3)
        if (pl == 1) {
            const int q = lc == 1;                   // the bit this lane lacks
            ld4(c.colA[2 + q] + lo0, w0);
            const double k = q ? k3 : k2;
#pragma unroll
            for (int t = 0; t < 4; ++t) w0[t] *= k;
        } else if (lc == 0) {
            ld4(c.colA[2] + lo0, w1a);
            ld4(c.colA[3] + lo0, w1b);
#pragma unroll
            for (int t = 0; t < 4; ++t) { w1a[t] *= k2; w1b[t] *= k3; }
        }
    }
    auto fin = [&]() {
        if (!ADJ) {
            val[0] = acc[0] * inv[0];
            val[1] = fma(e0a, val[0], acc[1]) * inv[1];
            val[2] = fma(e1a, val[0], acc[2]) * inv[2];
            val[3] = fma(e0b, val[2], fma(e1b, val[1], acc[3])) * inv[3];
        } else {
            val[3] = acc[3] * inv[3];
            val[2] = fma(e0b, val[3], acc[2]) * inv[2];
            val[1] = fma(e1b, val[3], acc[1]) * inv[1];
            val[0] = fma(e0a, val[1], fma(e1a, val[2], acc[0])) * inv[0];
        }
    };
    // Values of lanes that are not final yet are finite partial results and only ever meet a zero weight.
    fin();
    {
        const int src = ADJ ? (lane | 3) : (lane & ~3);
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] = fma(w0[t], __shfl_sync(0xffffffffu, val[t], src), acc[t]);
    }
    fin();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const double p = __shfl_xor_sync(0xffffffffu, val[t], 1);
        const double q = __shfl_xor_sync(0xffffffffu, val[t], 2);
        acc[t] = fma(w1b[t], q, fma(w1a[t], p, acc[t]));
    }
    fin();
}

template <bool ADJ, int MODE, class CTX>
__device__ __forceinline__ void solve_tile16(const SpaceDev& sp, const SpaceDev* __restrict__ spaces, const CTX& c,
                                             double* __restrict__ S, uint32_t cA, uint32_t row, bool valid, int lane)
{
    const int KC = c.KC, KR = c.KR;
    const int lc = lane & 3;
    const uint32_t lo0 = (cA << 4) | ((uint32_t)lc << 2);
    const uint32_t s0 = (row << KC) | lo0;
    double* v = S + (ADJ ? sp.x_off : sp.y_off);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (MODE == TM_WIDE) tile_rhs<ADJ, true>(sp, spaces, S, KC, KR, row, lo0, acc, KC + c.K2A, KR - c.K2A);
    else tile_rhs<ADJ>(sp, spaces, S, KC, KR, row, lo0, acc);
    // Edges are taken NB at a time, all loads of a round before its FMAs: a tile is a chain of dependent rounds and
    // thin levels have no other warps to hide the memory latency behind.
    // ---- column bits >= 4 (uniform over the warp): FWD visits the set bits of cA, ADJ the unset ones ----
    {
        constexpr int NB = TILE_NBA;
        uint32_t m = ADJ ? (~cA & ((1u << (KC - 4)) - 1u)) : cA;
        while (m) {
            double r[NB][4], y[NB][4], k[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const bool on = m != 0u;
                const int a = on ? __ffs(m) + 3 : 4;
                m &= m - 1;
                const uint32_t bit = 1u << a;
                k[q] = 1.0;
                if (on) {
                    ld4(c.colA[a] + (ADJ ? lo0 : (lo0 ^ bit)), r[q]);
                    ld4(v + (s0 ^ bit), y[q]);
                    if constexpr (MODE == TM_WIDE) k[q] = c.rowA[a][row & c.mH];
                    else if (MODE == TM_PROD) k[q] = c.rowA[a][row];
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) { r[q][t] = 0.0; y[q][t] = 0.0; }
                }
            }
#pragma unroll
            for (int q = 0; q < NB; ++q)
#pragma unroll
                for (int t = 0; t < 4; ++t) acc[t] = fma(MODE != TM_PAIR ? r[q][t] * k[q] : r[q][t], y[q][t], acc[t]);
        }
    }
    tile_row_edges<ADJ, MODE>(c, v, row, lo0, acc);
    double val[4];
    tile_tail<ADJ, MODE>(c, row, lo0, lane, acc, val);
    if (valid) st4(v + s0, val[0], val[1], val[2], val[3]);
}

// item: space, a = lA | lB << 8 | tiles << 16, b = first tile; tile t -> column block t / nBg, row group t % nBg
template <bool ADJ>
__global__ void __launch_bounds__(256, TILE_CTAS)
k_solve_tile(const SpaceDev* __restrict__ spaces, const Item* __restrict__ segs, const uint32_t* __restrict__ hs,
             const uint32_t* __restrict__ hsidx, double* __restrict__ S)
{
    __shared__ TileCtx ctx;
    const Item sg = segs[blockIdx.x];
    const SpaceDev& sp = spaces[sg.space];
    tile_ctx_build(ctx, sp, S, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lA = sg.a & 255u, lB = (sg.a >> 8) & 255u, cnt = sg.a >> 16;
    const uint32_t offA = hsidx[(ctx.KC - 4) * 32 + lA];
    const uint32_t offB = hsidx[ctx.KR * 32 + lB];
    const uint32_t nB = hsidx[ctx.KR * 32 + lB + 1] - offB;
    const uint32_t nBg = (nB + 7u) >> 3;
    const bool prod = sp.kind != K_JOINT;
    for (uint32_t q = w; q < cnt; q += 8) {
        const uint32_t t = sg.b + q;
        const uint32_t iA = t / nBg, jB = t - iA * nBg;
        const uint32_t ri = jB * 8u + (uint32_t)(lane >> 2);
        const bool valid = ri < nB;
        const uint32_t cA = hs[offA + iA];
        const uint32_t row = hs[offB + min(ri, nB - 1u)];
        if (prod) solve_tile16<ADJ, TM_PROD>(sp, spaces, ctx, S, cA, row, valid, lane);
        else      solve_tile16<ADJ, TM_PAIR>(sp, spaces, ctx, S, cA, row, valid, lane);
    }
}

// the same for wide pairs (TM_WIDE): a separate kernel, so that its extra index arithmetic stays out of the register
// budget of the two main paths
template <bool ADJ>
__global__ void __launch_bounds__(256, 3)
k_solve_tile_w(const SpaceDev* __restrict__ spaces, const Item* __restrict__ segs, const uint32_t* __restrict__ hs,
               const uint32_t* __restrict__ hsidx, double* __restrict__ S)
{
    __shared__ TileCtxW ctx;
    const Item sg = segs[blockIdx.x];
    const SpaceDev& sp = spaces[sg.space];
    tile_ctx_build_wide(ctx, sp, S, threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lA = sg.a & 255u, lB = (sg.a >> 8) & 255u, cnt = sg.a >> 16;
    const uint32_t offA = hsidx[(ctx.KC - 4) * 32 + lA];
    const uint32_t offB = hsidx[ctx.KR * 32 + lB];
    const uint32_t nB = hsidx[ctx.KR * 32 + lB + 1] - offB;
    const uint32_t nBg = (nB + 7u) >> 3;
    for (uint32_t q = w; q < cnt; q += 8) {
        const uint32_t t = sg.b + q;
        const uint32_t iA = t / nBg, jB = t - iA * nBg;
        const uint32_t ri = jB * 8u + (uint32_t)(lane >> 2);
        const bool valid = ri < nB;
        const uint32_t cA = hs[offA + iA];
        const uint32_t row = hs[offB + min(ri, nB - 1u)];
        solve_tile16<ADJ, TM_WIDE>(sp, spaces, ctx, S, cA, row, valid, lane);
    }
}

// ------------------------------------------------------------------------------------------
// Adjoint tile solve of a pair with the group-B marginal statistics fused in.  The adjoint pass already holds, for
// every state s and every row bit b not in s, the value x[s + b]; with y[s] (one more 32-byte load) the lane adds
//     stB[1+b][uB] += sum_uA y[s] x[s + b]      stB[0][uB] += sum_uA x[s] y[s]
// which saves the separate k_stats_b pass (one more full read of x and y plus KB/2 re-reads of x through L2).
// A CTA is G row groups x C column blocks (G C <= 32) of one (lA, lB) split: the per-tile sums are parked in shared
// memory, added over the CTA's column blocks in a fixed order and written to the partial table `slot` of the space
// (every row receives every slot exactly once, no atomics; k_stats_reduce adds the slots).
// item: a = lA | lB << 8, b = first row group, c = column chunk k | slot << 16
constexpr int ADJB_MAXE = 17;                       // g + up to 16 row bits (pairs with plain tables have KB <= MAXT)

__device__ __forceinline__ double group4_sum(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

__global__ void __launch_bounds__(256, ADJB_CTAS)
k_solve_tile_adjb(const SpaceDev* __restrict__ spaces, const Item* __restrict__ segs, const uint32_t* __restrict__ hs,
                  const uint32_t* __restrict__ hsidx, double* __restrict__ S)
{
    __shared__ TileCtx ctx;
    __shared__ double pb[32][8][ADJB_MAXE];
    const Item sg = segs[blockIdx.x];
    const SpaceDev& sp = spaces[sg.space];
    tile_ctx_build(ctx, sp, S, threadIdx.x);
    for (int t = threadIdx.x; t < 32 * 8 * ADJB_MAXE; t += blockDim.x) (&pb[0][0][0])[t] = 0.0;
    __syncthreads();
    const TileCtx& c = ctx;
    const int KC = c.KC, KR = c.KR;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lc = lane & 3, lg = lane >> 2;
    const uint32_t lA = sg.a & 255u, lB = (sg.a >> 8) & 255u;
    const uint32_t kch = sg.c & 0xffffu, slot = sg.c >> 16;
    const uint32_t offA = hsidx[(KC - 4) * 32 + lA], nA = hsidx[(KC - 4) * 32 + lA + 1] - offA;
    const uint32_t offB = hsidx[KR * 32 + lB], nB = hsidx[KR * 32 + lB + 1] - offB;
    const uint32_t nBg = (nB + 7u) >> 3;
    uint32_t C = 32;                                   // column blocks per row group in this CTA
    if (nA < 32u) { C = 1; while (C < nA) C <<= 1; }
    const uint32_t G = 32u / C;
    double* v = S + sp.x_off;
    const double* yv = S + sp.y_off;
    for (uint32_t q = w; q < 32u; q += 8) {
        const uint32_t g = q / C, ci = q - g * C;
        const uint32_t jB = sg.b + g, iA = kch * 32u + ci;
        if (jB >= nBg || iA >= nA) continue;           // uniform over the warp
        const uint32_t ri = jB * 8u + (uint32_t)lg;
        const bool valid = ri < nB;
        const uint32_t cA = hs[offA + iA];
        const uint32_t row = hs[offB + min(ri, nB - 1u)];
        const uint32_t lo0 = (cA << 4) | ((uint32_t)lc << 2);
        const uint32_t s0 = (row << KC) | lo0;
        double acc[4] = {0.0, 0.0, 0.0, 0.0}, y4[4];
        ld4(yv + s0, y4);
        tile_rhs<true>(sp, spaces, S, KC, KR, row, lo0, acc);
        // column bits >= 4
        {
            constexpr int NB = TILE_NBA;
            uint32_t m = ~cA & ((1u << (KC - 4)) - 1u);
            while (m) {
                double r[NB][4], y[NB][4];
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const bool on = m != 0u;
                    const int a = on ? __ffs(m) + 3 : 4;
                    m &= m - 1;
                    if (on) { ld4(c.colA[a] + lo0, r[e]); ld4(v + (s0 | (1u << a)), y[e]); }
                    else {
#pragma unroll
                        for (int t = 0; t < 4; ++t) { r[e][t] = 0.0; y[e][t] = 0.0; }
                    }
                }
#pragma unroll
                for (int e = 0; e < NB; ++e)
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[t] = fma(r[e][t], y[e][t], acc[t]);
            }
        }
        // row bits: solve edge and statistic from the same loaded values
        {
            constexpr int NB = TILE_NBB;
            uint32_t m = ~row & ((1u << KR) - 1u);
            while (m) {
                double y[NB][4], k[NB];
                int bq[NB];
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const bool on = m != 0u;
                    const int b = on ? __ffs(m) - 1 : 0;
                    m &= m - 1;
                    bq[e] = on ? b : -1;
                    if (on) {
                        k[e] = c.rowB[b][row];
                        ld4(v + (((uint64_t)(row | (1u << b)) << KC) | lo0), y[e]);
                    } else {
                        k[e] = 0.0;
#pragma unroll
                        for (int t = 0; t < 4; ++t) y[e][t] = 0.0;
                    }
                }
#pragma unroll
                for (int e = 0; e < NB; ++e) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[t] = fma(k[e], y[e][t], acc[t]);
                    const double d = group4_sum(fma(y4[3], y[e][3], fma(y4[2], y[e][2], fma(y4[1], y[e][1], y4[0] * y[e][0]))));
                    if (lc == 0 && bq[e] >= 0) pb[q][lg][1 + bq[e]] = d;
                }
            }
        }
        double val[4];
        tile_tail<true, TM_PAIR>(c, row, lo0, lane, acc, val);
        const double gsum = group4_sum(fma(y4[3], val[3], fma(y4[2], val[2], fma(y4[1], val[1], y4[0] * val[0]))));
        if (lc == 0) pb[q][lg][0] = gsum;
        if (valid) st4(v + s0, val[0], val[1], val[2], val[3]);
    }
    __syncthreads();
    // add the CTA's column blocks (fixed order) and write the partial table of this slot
    const uint32_t NBr = 1u << KR;
    double* out = S + sp.stPB + (uint64_t)slot * (KR + 1) * NBr;
    for (uint32_t t = threadIdx.x; t < G * 8u * (uint32_t)(KR + 1); t += blockDim.x) {
        const uint32_t e = t % (uint32_t)(KR + 1), rr = (t / (uint32_t)(KR + 1)) & 7u, g = t / ((uint32_t)(KR + 1) * 8u);
        const uint32_t jB = sg.b + g;
        const uint32_t ri = jB * 8u + rr;
        if (jB >= nBg || ri >= nB) continue;
        double s = 0.0;
        for (uint32_t ci = 0; ci < C; ++ci) s += pb[g * C + ci][rr][e];
        out[(uint64_t)e * NBr + hs[offB + ri]] = s;
    }
}

__global__ void k_logp(const SpaceDev* __restrict__ spaces, const uint32_t* __restrict__ list, uint32_t count,
                       const double* __restrict__ S, double* __restrict__ logp)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const SpaceDev& sp = spaces[list[t]];
    const uint32_t NA = 1u << sp.KA;
    double v;
    if (sp.kind == K_S1) v = log(S[sp.y_off + NA - 1u]);
    else if (sp.kind == K_S2) v = log(S[sp.y_off + NA - 1u] * side_of(sp, 0, S).special(ROW_DM, NA - 1u));
    else v = log(joint_score(sp, spaces, S));
    logp[sp.patient] = v;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// direct dependence of the second-phase start vector v = c * (slice of the joint y) on the diagnosis effects
// (likelihood.py:574-576, 617-619):  T = x2 . v  with  A2 p2 = v  and  A2^T x2 = e_last / score,
// hence  T = e_last^T p2 / score = score_side / score.  One thread per joint space.
__global__ void k_direct(const SpaceDev* __restrict__ spaces, const uint32_t* __restrict__ list, uint32_t count,
                         const double* __restrict__ S, double* __restrict__ tdir)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const SpaceDev& j = spaces[list[t]];
    const double sc = joint_score(j, spaces, S);
    double tp = 0.0, tm = 0.0;
    if (j.has_pf) { const SpaceDev& q = spaces[j.pf]; tp = S[q.y_off + ((1u << q.KA) - 1u)] / sc; }
    if (j.has_mf) { const SpaceDev& q = spaces[j.mf]; tm = S[q.y_off + ((1u << q.KA) - 1u)] / sc; }
    tdir[2 * t] = tp;
    tdir[2 * t + 1] = tm;
}

// diracc[side][e] += sum over the chunk's joint spaces of wgt * T_side * [e in mask_side or e == n]
__global__ void k_direct_acc(const SpaceDev* __restrict__ spaces, const uint32_t* __restrict__ list, uint32_t count,
                             const double* __restrict__ tdir, double w_other, double* __restrict__ diracc)
{
    const int e = blockIdx.x, side = blockIdx.y;
    __shared__ double red[256];
    double s = 0.0;
    for (uint32_t t = threadIdx.x; t < count; t += blockDim.x) {
        const SpaceDev& j = spaces[list[t]];
        const uint32_t mask = (side ? j.mtmask : j.ptmask) | (1u << (j.n_tot - 1));
        if ((mask >> e) & 1u) s += tdir[2 * t + side];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) diracc[side * NR + e] += w_other * red[0];
}

// ------------------------------------------------------------------------------------------
// Marginal statistics of a joint space:
//   stA[0][uA]   = sum_uB x y                     stB[0][uB]   = sum_uA x y
//   stA[1+a][uA] = sum_uB y[uB,uA] x[uB,uA|a]     stB[1+a][uB] = sum_uA y[uB,uA] x[uB|a,uA]
// They are all the joint pass has to deliver: every theta / d_p / d_m gradient entry is a
// contraction of these small tables with the group rate tables (k_finish).
// group-A statistics: a lane owns TWO consecutive uA (16-byte loads), a warp 64; bit 0 is resolved inside the
// lane, bits 1..5 with shuffles, higher bits by a loop over the bits that are free in this chunk.
__global__ void __launch_bounds__(256)
k_stats_a(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items, uint32_t count, double* __restrict__ S)
{
    const uint32_t wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wg >= count) return;
    const Item it = items[wg];                       // a = chunk of 64 uA, b = slice
    const SpaceDev& sp = spaces[it.space];
    const int KA = sp.KA;
    const uint32_t NA = 1u << KA, NB = 1u << sp.KB;
    const uint32_t per = (NB + sp.slices - 1) / sp.slices;
    const uint32_t b0 = it.b * per, b1 = min(NB, b0 + per);
    const double* y = S + sp.y_off;
    const double* x = S + sp.x_off;
    double* out = S + sp.stP + (uint64_t)it.b * (KA + 1) * NA;
    if (KA == 0) {                                   // a single uA: only g
        if (lane == 0) { double g = 0.0; for (uint32_t uB = b0; uB < b1; ++uB) g = fma(x[uB], y[uB], g); out[0] = g; }
        return;
    }
    const uint32_t uA = (it.a << 6) | (lane << 1);   // first of the two states of this lane
    const bool valid = uA < NA;
    double g0 = 0.0, g1 = 0.0, a0 = 0.0;             // bit 0: only the even state lacks it
    double l0[5] = {0, 0, 0, 0, 0}, l1[5] = {0, 0, 0, 0, 0};
    // Bits 6..11 are uniform over the warp's 64 sub-states.  The slab of rows is walked in sub-slabs of SUB rows: bits
    // 0..5 first, then one short loop per absent high bit over the SAME rows, whose y values are still in L1 (walking
    // the whole slab once per high bit re-read it from DRAM 3x, profiles/r1_v11 capture).
#ifndef STA_MAXHB
#define STA_MAXHB 6             // column bits 6..11 with register accumulators (A/B tested against 10)
#endif
#ifndef STA_SUB
#define STA_SUB 16
#endif
    constexpr int HB0 = 6, MAXHB = STA_MAXHB, SUB = STA_SUB;
    double h0[MAXHB], h1[MAXHB];
#pragma unroll
    for (int q = 0; q < MAXHB; ++q) { h0[q] = 0.0; h1[q] = 0.0; }
    uint32_t hmask = 0;
    if (valid)
        for (int a = HB0; a < KA && a < HB0 + MAXHB; ++a) if (!((uA >> a) & 1u)) hmask |= 1u << (a - HB0);
    for (uint32_t u0 = b0; u0 < b1; u0 += SUB) {
        const uint32_t u1 = min(b1, u0 + SUB);
#pragma unroll 4
        for (uint32_t uB = u0; uB < u1; ++uB) {
            const uint64_t s = ((uint64_t)uB << KA) | uA;
            double2 yv = make_double2(0.0, 0.0), xv = yv;
            if (valid) { yv = *reinterpret_cast<const double2*>(y + s); xv = *reinterpret_cast<const double2*>(x + s); }
            g0 = fma(xv.x, yv.x, g0); g1 = fma(xv.y, yv.y, g1);
            a0 = fma(yv.x, xv.y, a0);
#pragma unroll
            for (int a = 1; a <= 5; ++a) {
                const double px = __shfl_xor_sync(0xffffffffu, xv.x, 1 << (a - 1));
                const double py = __shfl_xor_sync(0xffffffffu, xv.y, 1 << (a - 1));
                if (a < KA && !((lane >> (a - 1)) & 1)) { l0[a - 1] = fma(yv.x, px, l0[a - 1]); l1[a - 1] = fma(yv.y, py, l1[a - 1]); }
            }
        }
#pragma unroll
        for (int q = 0; q < MAXHB; ++q) {
            if (!((hmask >> q) & 1u)) continue;
            const uint64_t bit = (uint64_t)64 << q;
            double e0 = 0.0, e1 = 0.0;
#pragma unroll 4
            for (uint32_t uB = u0; uB < u1; ++uB) {
                const uint64_t s = ((uint64_t)uB << KA) | uA;
                const double2 yv = *reinterpret_cast<const double2*>(y + s);
                const double2 xa = *reinterpret_cast<const double2*>(x + (s | bit));
                e0 = fma(yv.x, xa.x, e0); e1 = fma(yv.y, xa.y, e1);
            }
            h0[q] += e0; h1[q] += e1;
        }
    }
    if (valid) {
        out[uA] = g0; out[uA + 1] = g1;
        out[(uint64_t)NA + uA] = a0; out[(uint64_t)NA + uA + 1] = 0.0;
#pragma unroll
        for (int a = 1; a <= 5; ++a)
            if (a < KA) { out[(uint64_t)(1 + a) * NA + uA] = l0[a - 1]; out[(uint64_t)(1 + a) * NA + uA + 1] = l1[a - 1]; }
#pragma unroll
        for (int q = 0; q < MAXHB; ++q)
            if (HB0 + q < KA) { out[(uint64_t)(1 + HB0 + q) * NA + uA] = h0[q]; out[(uint64_t)(1 + HB0 + q) * NA + uA + 1] = h1[q]; }
    }
    // bits >= 6 + MAXHB: one sweep of the slab per bit
    for (int a = HB0 + MAXHB; a < KA; ++a) {
        if (!valid) break;
        double e0 = 0.0, e1 = 0.0;
        if (!((uA >> a) & 1u)) {
            const uint64_t bit = 1ull << a;
            for (uint32_t uB = b0; uB < b1; ++uB) {
                const uint64_t s = ((uint64_t)uB << KA) | uA;
                const double2 yv = *reinterpret_cast<const double2*>(y + s);
                const double2 xa = *reinterpret_cast<const double2*>(x + (s | bit));
                e0 = fma(yv.x, xa.x, e0); e1 = fma(yv.y, xa.y, e1);
            }
        }
        out[(uint64_t)(1 + a) * NA + uA] = e0; out[(uint64_t)(1 + a) * NA + uA + 1] = e1;
    }
}

__global__ void k_stats_reduce(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items,
                               double* __restrict__ S)
{
    const Item it = items[blockIdx.x];               // a = group, b = first element of a 1024-element block (units of 1024)
    const SpaceDev& sp = spaces[it.space];
    const int KG = it.a ? sp.KB : sp.KA;
    const uint64_t len = (uint64_t)(KG + 1) << KG;
    const uint64_t t = (uint64_t)it.b * 1024u + threadIdx.x;
    if (t >= len) return;
    const uint32_t ns = it.a ? sp.slicesB : sp.slices;
    const double* src = S + (it.a ? sp.stPB : sp.stP);
    double s = 0.0;
    uint32_t k = 0;
    for (; k + 8 <= ns; k += 8) {                            // eight loads in flight, added in slot order
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = src[(uint64_t)(k + q) * len + t];
#pragma unroll
        for (int q = 0; q < 8; ++q) s += v[q];
    }
    for (; k < ns; ++k) s += src[(uint64_t)k * len + t];
    S[(it.a ? sp.stB : sp.stA) + t] = s;
}

// group-B statistics: one warp per uB; dot products of the y row with the x rows of up to four free bits at a
// time, 16-byte loads.
__device__ __forceinline__ void row_dots4(const double* __restrict__ yr, const double* const (&xr)[4], uint32_t NA,
                                          uint32_t i0, uint32_t i1, int lane, double (&acc)[4])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = 0.0;
    if (NA >= 2) {
        for (uint32_t i = i0 + 2 * lane; i < i1; i += 64) {
            const double2 yv = *reinterpret_cast<const double2*>(yr + i);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 xv = *reinterpret_cast<const double2*>(xr[q] + i);
                acc[q] = fma(yv.x, xv.x, acc[q]); acc[q] = fma(yv.y, xv.y, acc[q]);
            }
        }
    } else if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = yr[0] * xr[q][0];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
}

__global__ void __launch_bounds__(256)
k_stats_b(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items, uint32_t count, double* __restrict__ S)
{
    const uint32_t wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wg >= count) return;
    const Item it = items[wg];                       // a = uB, b = slice of the uA range (one warp per item)
    const SpaceDev& sp = spaces[it.space];
    const int KA = sp.KA, KB = sp.KB;
    const uint32_t NA = 1u << KA, NB = 1u << KB;
    const uint32_t uB = it.a;
    const uint32_t per = ((NA + sp.slicesB - 1) / sp.slicesB + 63u) & ~63u;
    const uint32_t i0 = min(NA, it.b * per), i1 = min(NA, i0 + per);
    const double* yr = S + sp.y_off + ((uint64_t)uB << KA);
    const double* xb = S + sp.x_off;
    double* out = S + sp.stPB + (uint64_t)it.b * (KB + 1) * NB;
    // row 0 (g = sum x y) travels with the first group of free bits
    int rows[4] = {0, -1, -1, -1};
    int nrow = 1;
    for (int a = 0; a <= KB; ++a) {
        const bool last = (a == KB);
        if (!last) {
            if ((uB >> a) & 1u) { if (lane == 0) out[(uint64_t)(1 + a) * NB + uB] = 0.0; }
            else rows[nrow++] = 1 + a;
        }
        if (nrow == 4 || (last && nrow > 0)) {
            const double* xr[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q < nrow ? rows[q] : 0;
                xr[q] = xb + ((uint64_t)(r == 0 ? uB : (uB | (1u << (r - 1)))) << KA);
            }
            double acc[4];
            row_dots4(yr, xr, NA, i0, i1, lane, acc);
            if (lane == 0)
#pragma unroll
                for (int q = 0; q < 4; ++q) if (q < nrow) out[(uint64_t)rows[q] * NB + uB] = acc[q];
            nrow = 0;
        }
    }
}

// group-B statistics of pairs with at most STB_NARROW PT bits (small pairs, and pairs whose PT has almost no events): a row
// has at most 32 columns, a warp per row would idle most of its lanes -- one LANE per uB instead, the row sums run in the
// lane.  One partial table (slot 0).   item: a = first uB of 32
constexpr int STB_NARROW = 5;
__global__ void __launch_bounds__(256)
k_stats_b_narrow(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items, uint32_t count, double* __restrict__ S)
{
    const uint32_t wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wg >= count) return;
    const Item it = items[wg];
    const SpaceDev& sp = spaces[it.space];
    const int KA = sp.KA, KB = sp.KB;
    const uint32_t NA = 1u << KA, NB = 1u << KB;
    const uint32_t uB = it.a + (uint32_t)lane;
    if (uB >= NB) return;
    const double* yr = S + sp.y_off + ((uint64_t)uB << KA);
    const double* xb = S + sp.x_off;
    double* out = S + sp.stPB;
    {
        const double* xr = xb + ((uint64_t)uB << KA);
        double g = 0.0;
        for (uint32_t i = 0; i < NA; ++i) g = fma(yr[i], xr[i], g);
        out[uB] = g;
    }
    for (int b = 0; b < KB; ++b) {
        double s = 0.0;
        if (!((uB >> b) & 1u)) {
            const double* xr = xb + ((uint64_t)(uB | (1u << b)) << KA);
            for (uint32_t i = 0; i < NA; ++i) s = fma(yr[i], xr[i], s);
        }
        out[(uint64_t)(1 + b) * NB + uB] = s;
    }
}

// ------------------------------------------------------------------------------------------
// Single-tumour spaces in product form (K1 low bits "lo", K2 high bits "hi"; rate(r,u) = T1[r][lo] T2[r][hi]).
//
// Row metadata shared by the three kernels below: for table row r of a product-form space
//   kind 0: unused   1: event absent in this patient (or a diagnosis pseudo row)   2: event on lo bit `bit`
//   3: event on hi bit `bit`
struct PfRows { int8_t kind[NR], bit[NR], act[NR]; int n_act; };   // act: the rows in use, events of the patient first
__device__ __forceinline__ void pf_rows_build(PfRows& m, const SpaceDev& sp, int t)
{
    if (t >= NR) return;
    const int K1 = sp.splitA;
    const SetupSel sel = setup_sel(sp, 0);
    int kind = t < sel.nrows ? 1 : 0, bit = 0;
    if ((t == ROW_DP || t == ROW_DM) && sel.dg == 4) kind = 1;
    if (t < sel.nrows)
        for (int b = 0; b < sp.KA; ++b)
            if (sp.evA[b] == t) { kind = b < K1 ? 2 : 3; bit = b < K1 ? b : b - K1; }
    m.kind[t] = (int8_t)kind; m.bit[t] = (int8_t)bit;
    // compact list of the rows in use (called by the first warp of the CTA): rows of present events (one neighbour load per
    // state each) first, then absent events / pseudo rows, so that dealing the list round-robin balances the warps of k_pf
    const unsigned mp = __ballot_sync(0xffffffffu, kind >= 2), mo = __ballot_sync(0xffffffffu, kind == 1);
    const unsigned below = (1u << t) - 1u;
    if (kind >= 2) m.act[__popc(mp & below)] = (int8_t)t;
    else if (kind == 1) m.act[__popc(mp) + __popc(mo & below)] = (int8_t)t;
    if (t == 0) m.n_act = __popc(mp) + __popc(mo);
}

// Diagonal of a product-form space (replaces the per-state 26-term loop of the first version):
//   D[hi][lo] = d(u) + sum_{r not in u} T1[r][lo] T2[r][hi]
// is a rank-(nrows+2) product once the rows of present events are masked on their own bit; d(u) is 1
// (diagnosis_theta form, vanilla.py:269) or, for type 2, the two diagnosis-rate products held in rows 30/31.
// A small FP64 matrix product [hi x rows] [rows x lo]: the masked factor tiles of the CTA (32 hi x 128 lo) are staged in
// shared memory once, a thread owns 4 hi x 4 lo (16 accumulators, four 16-byte shared loads per 16 FMAs).
// The lo tile is reused for up to DG_HIB / DG_HI hi tiles.   item: a = lo block of 128, b = hi block of DG_HIB
constexpr int DG_HI = 32, DG_LO = 128, DG_HIB = 64;
__global__ void __launch_bounds__(256)
k_diag_prod(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items, double* __restrict__ S)
{
    __shared__ PfRows rows;
    __shared__ __align__(16) double fa[NR][DG_LO];       // T1[r][lo], zero where the row does not count
    __shared__ __align__(16) double fb[NR][DG_HI];       // T2[r][hi]
    const Item it = items[blockIdx.x];
    const SpaceDev& sp = spaces[it.space];
    pf_rows_build(rows, sp, threadIdx.x);
    __syncthreads();
    const int K1 = sp.splitA, K2 = sp.KA - K1;
    const uint32_t N1 = 1u << K1, N2 = 1u << K2;
    const double* T1 = S + sp.tabA;
    const double* T2 = T1 + ((uint64_t)NR << K1);
    double* vec = S + sp.tabA + ((uint64_t)NR << K1) + ((uint64_t)NR << K2);
    const bool s2 = setup_sel(sp, 0).dg == 4;
    const uint32_t lob = it.a << 7;
    // rows 0..28 are events; for type 2 the two diagnosis-rate products (rows 30 / 31) join the sum
    for (int t = threadIdx.x; t < NR * DG_LO; t += blockDim.x) {
        const int r = t >> 7;
        const uint32_t lo = lob + ((uint32_t)t & (DG_LO - 1));
        const int kind = rows.kind[r];
        double v = 0.0;
        const bool use = (r < ROW_D && kind != 0) || (s2 && r >= ROW_DP);
        if (use && lo < N1 && !(kind == 2 && r < ROW_D && ((lo >> rows.bit[r]) & 1u))) v = T1[((uint64_t)r << K1) + lo];
        fa[r][t & (DG_LO - 1)] = v;
    }
    const int tl = threadIdx.x & 31, th = threadIdx.x >> 5;            // 32 column groups x 8 row groups
    const uint32_t lo0 = lob + ((uint32_t)tl << 2);
    const uint64_t NG = (uint64_t)N1 << K2;
    for (uint32_t hib = it.b * DG_HIB; hib < min(N2, (it.b + 1u) * DG_HIB); hib += DG_HI) {
        __syncthreads();                                               // the previous hi tile is consumed; (first) fa is in place
        for (int t = threadIdx.x; t < NR * DG_HI; t += blockDim.x) {
            const int r = t >> 5;
            const uint32_t hi = hib + ((uint32_t)t & (DG_HI - 1));
            const int kind = rows.kind[r];
            double v = 0.0;
            const bool use = (r < ROW_D && kind != 0) || (s2 && r >= ROW_DP);
            if (use && hi < N2 && !(kind == 3 && r < ROW_D && ((hi >> rows.bit[r]) & 1u))) v = T2[((uint64_t)r << K2) + hi];
            fb[r][t & (DG_HI - 1)] = v;
        }
        __syncthreads();
        const uint32_t hi0 = hib + ((uint32_t)th << 2);
        double d[4][4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
            for (int t = 0; t < 4; ++t) d[h][t] = s2 ? 0.0 : 1.0;
#pragma unroll 4
        for (int r = 0; r < ROW_D; ++r) {
            const double2 b01 = *reinterpret_cast<const double2*>(&fa[r][tl << 2]), b23 = *reinterpret_cast<const double2*>(&fa[r][(tl << 2) + 2]);
            const double2 a01 = *reinterpret_cast<const double2*>(&fb[r][th << 2]), a23 = *reinterpret_cast<const double2*>(&fb[r][(th << 2) + 2]);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int h = 0; h < 4; ++h)
#pragma unroll
                for (int t = 0; t < 4; ++t) d[h][t] = fma(a[h], b[t], d[h][t]);
        }
        if (lo0 >= N1) continue;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            if (hi0 + h >= N2) break;
            const uint64_t u0 = ((uint64_t)(hi0 + h) << K1) | lo0;
            if (s2) {
                double vp[4], vm[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    vp[t] = fb[ROW_DP][(th << 2) + h] * fa[ROW_DP][(tl << 2) + t];
                    vm[t] = fb[ROW_DM][(th << 2) + h] * fa[ROW_DM][(tl << 2) + t];
                    d[h][t] += vp[t] + vm[t];
                }
                st4(vec + NG + u0, vp[0], vp[1], vp[2], vp[3]);
                st4(vec + 2 * NG + u0, vm[0], vm[1], vm[2], vm[3]);
            }
            st4(vec + u0, d[h][0], d[h][1], d[h][2], d[h][3]);
        }
    }
}

// Weighted marginals of a product-form space.  With x the adjoint and y the forward vector, row r of the gradient
// needs  sum_u T1[r][lo] T2[r][hi] E_r(u) [b in u]  with
//     E_r(u) = -x(u) y(u)                          event r absent in this patient (and the diagnosis pseudo rows)
//            = [a not in u] y(u) (x(u + a) - x(u))  event r sits on bit a
// (vanilla.py:328-393 spends n_tot x_partial_Q_y passes on this).  Summing out hi first and lo first gives
//     H1[r][lo] = sum_hi T2[r][hi] E_r(hi,lo)        H2[r][hi] = sum_lo T1[r][lo] E_r(hi,lo)
// and k_finish folds T1 H1 over the lo bits, T2 H2 over the hi bits.  A lane owns four consecutive lo (32-byte
// loads), a warp PF_RW table rows, the PF_WARPS warps of a CTA the 32 rows.
// Output (stP): slices x (NR + KA) x N1 partial H1 tables, then (NR + KA) x N2 for H2 (rows >= NR unused).
__device__ __forceinline__ void pf_E(int kind, int bit, uint32_t hi, uint32_t lo0, int K1, const double* __restrict__ x,
                                     const double (&xv)[4], const double (&yv)[4], const double (&ng)[4], double (&e)[4], bool& zero)
{
    zero = false;
    if (kind == 1) {
#pragma unroll
        for (int t = 0; t < 4; ++t) e[t] = ng[t];
        return;
    }
    if (kind == 2) {
        if (bit == 0) {
            e[0] = yv[0] * (xv[1] - xv[0]); e[1] = 0.0; e[2] = yv[2] * (xv[3] - xv[2]); e[3] = 0.0;
        } else if (bit == 1) {
            e[0] = yv[0] * (xv[2] - xv[0]); e[1] = yv[1] * (xv[3] - xv[1]); e[2] = 0.0; e[3] = 0.0;
        } else {
            if ((lo0 >> bit) & 1u) { zero = true; return; }
            double xa[4];
            ld4(x + (((uint64_t)hi << K1) | (lo0 | (1u << bit))), xa);
#pragma unroll
            for (int t = 0; t < 4; ++t) e[t] = yv[t] * (xa[t] - xv[t]);
        }
        return;
    }
    if ((hi >> bit) & 1u) { zero = true; return; }
    double xa[4];
    ld4(x + (((uint64_t)(hi | (1u << bit)) << K1) | lo0), xa);
#pragma unroll
    for (int t = 0; t < 4; ++t) e[t] = yv[t] * (xa[t] - xv[t]);
}

constexpr int PF_RW = 4, PF_WARPS = NR / PF_RW, PF_HBATCH = 4;   // table rows per warp, warps per CTA, hi per reduction round
// Both weighted marginals in ONE pass over the lattice (round 2; before, k_pf_lo and k_pf_hi each re-read x, y and every
// lattice neighbour).  CTA = a block of 128 lo x a slice of hi; lane = four consecutive lo, warp = PF_RW table rows:
//   H1: the lane keeps acc[row][lo] over the hi of the slice -> partial table of the slice (as before)
//   H2: per hi the lane's  sum_t T1[row][lo+t] E(hi, lo+t)  is parked in shared memory, PF_HBATCH hi at a time, and
//       summed over the 32 lanes by one lane per (row, hi) in a fixed order -> partial table of this lo block
// k_finish adds the slices / the lo blocks.   item: a = block of 128 lo, b = hi slice
// Output (stP): slices x (NR + KA) x N1 partial H1 tables, then max(1, N1/128) x (NR + KA) x N2 partial H2 tables.
#ifndef PF_CTAS
#define PF_CTAS 2
#endif
__global__ void __launch_bounds__(32 * PF_WARPS, PF_CTAS)
k_pf(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items, double* __restrict__ S)
{
    __shared__ PfRows rows;
    __shared__ double red[PF_WARPS][PF_RW * PF_HBATCH][33];
    const Item it = items[blockIdx.x];
    const SpaceDev& sp = spaces[it.space];
    pf_rows_build(rows, sp, threadIdx.x);
    __syncthreads();
    const int KA = sp.KA, K1 = sp.splitA, K2 = KA - K1;
    const uint32_t N1 = 1u << K1, N2 = 1u << K2;
    const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const uint32_t lo0 = (it.a << 7) | ((uint32_t)lane << 2);
    const bool live = lo0 < N1;                           // narrow low parts leave lanes without columns
    const uint32_t per = (N2 + sp.slices - 1) / sp.slices;
    const uint32_t h0 = it.b * per, h1 = min(N2, h0 + per);
    const double* x = S + sp.x_off;
    const double* y = S + sp.y_off;
    const double* T1 = S + sp.tabA;
    const double* T2 = T1 + ((uint64_t)NR << K1);
    int kind[PF_RW], bit[PF_RW], row[PF_RW];              // the rows in use are dealt round-robin to the warps
    double acc[PF_RW][4], tvr[PF_RW][4];
#pragma unroll
    for (int j = 0; j < PF_RW; ++j) {
        const int slot = rg + PF_WARPS * j;
        row[j] = slot < rows.n_act ? rows.act[slot] : 0;
        kind[j] = (live && slot < rows.n_act) ? rows.kind[row[j]] : 0;
        bit[j] = rows.bit[row[j]];
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[j][t] = 0.0; tvr[j][t] = 0.0; }
        if (kind[j] != 0) ld4(T1 + ((uint64_t)row[j] << K1) + lo0, tvr[j]);
    }
    const bool any_row = (kind[0] | kind[1] | kind[2] | kind[3]) != 0;     // warps without a live row skip the loads
    double* out2 = S + sp.stP + (uint64_t)sp.slices * (NR + KA) * N1 + (uint64_t)it.a * (NR + KA) * N2;
    for (uint32_t hb = h0; hb < h1; hb += PF_HBATCH) {
#pragma unroll
        for (int hh = 0; hh < PF_HBATCH; ++hh) {
            const uint32_t hi = hb + hh;
            double p[PF_RW];
#pragma unroll
            for (int j = 0; j < PF_RW; ++j) p[j] = 0.0;
            if (any_row && hi < h1) {
                const uint64_t u0 = ((uint64_t)hi << K1) | lo0;
                double xv[4], yv[4], ng[4];
                ld4(x + u0, xv);
                ld4(y + u0, yv);
#pragma unroll
                for (int t = 0; t < 4; ++t) ng[t] = -(xv[t] * yv[t]);
#pragma unroll
                for (int j = 0; j < PF_RW; ++j) {
                    if (kind[j] == 0) continue;
                    double e[4];
                    bool zero;
                    pf_E(kind[j], bit[j], hi, lo0, K1, x, xv, yv, ng, e, zero);
                    if (zero) continue;
                    const double tw = T2[((uint64_t)row[j] << K2) + hi];
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[j][t] = fma(tw, e[t], acc[j][t]);
                    p[j] = fma(tvr[j][3], e[3], fma(tvr[j][2], e[2], fma(tvr[j][1], e[1], tvr[j][0] * e[0])));
                }
            }
#pragma unroll
            for (int j = 0; j < PF_RW; ++j) red[rg][j * PF_HBATCH + hh][lane] = p[j];
        }
        __syncwarp();
        if (lane < PF_RW * PF_HBATCH) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < 32; ++l) s += red[rg][lane][l];
            const int j = lane / PF_HBATCH, hh = lane % PF_HBATCH;
            if (hb + hh < h1 && rg + PF_WARPS * j < rows.n_act) out2[(uint64_t)rows.act[rg + PF_WARPS * j] * N2 + hb + hh] = s;
        }
        __syncwarp();
    }
    if (!live) return;
    double* out = S + sp.stP + (uint64_t)it.b * (NR + KA) * N1;
#pragma unroll
    for (int j = 0; j < PF_RW; ++j)
        if (rg + PF_WARPS * j < rows.n_act) st4(out + (uint64_t)row[j] * N1 + lo0, acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
}
// ------------------------------------------------------------------------------------------
// Gradient contraction.  For event row i and sub-state u of a group (i not in u)
//     w_i(u) = T[i][u] * E_i(u),   E_i(u) = sum_other y (x[u + i] - x[u])   (i is a bit of the group)
//                                          = - sum_other x y                (i absent in this patient)
// and  dL/dlogW[i][ev(b)] += w_i(u) for every bit b of u,  dL/dlogW[i][i] += w_i(u)
// (likelihood.py:125-201 and vanilla.py:328-393 do this with one shuffle pass per (i, j)).
// A warp takes FIN_U sub-states of one group, 32 at a time:
//   phase 1, lane = sub-state: w_r(u) for every row r with coalesced table / statistics reads -> shared tile
//   phase 2, lane = row: the 32 values of the row are folded into per-bit sums; the membership of the five low
//            bits is known at compile time, the higher bits are uniform over the tile.
// The per-item results go into one shared accumulator per CTA in a fixed warp order (no atomics): repeated
// evaluations are bit-identical.
constexpr int NACC = 3;                               // effective-parameter spaces: theta, theta_pt/d_p, theta/d_m
constexpr int FIN_WARPS = 4;
template <int MB>
__global__ void __launch_bounds__(FIN_WARPS * 32)
k_finish(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items, uint32_t count,
         const double* __restrict__ S, double w_type0, double w_other, double* __restrict__ partial)
{
    extern __shared__ double sm[];                    // [NACC][NR][NR] accumulator, then FIN_WARPS tiles of 32 x 33
    double* G = sm;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* tile = sm + NACC * NR * NR + w * (NR * 33);
    for (int t = threadIdx.x; t < NACC * NR * NR; t += blockDim.x) G[t] = 0.0;
    __syncthreads();
    const uint32_t per_cta = (count + gridDim.x - 1) / gridDim.x;
    const uint32_t c0 = blockIdx.x * per_cta, c1 = min(count, c0 + per_cta);
    for (uint32_t base = c0; base < c1; base += FIN_WARPS) {
        const uint32_t k = base + w;
        const bool have = k < c1;
        double tot = 0.0, ac[MB];
#pragma unroll
        for (int b = 0; b < MB; ++b) ac[b] = 0.0;
        int KG = 0, accid = 0, n = 0, pmode = 0;
        bool always_n = false, pseudo_tot = false, is_row = false, is_pseudo = false;
        const uint8_t* ev = nullptr;
        double wgt = 0.0;
        if (have) {
            const Item it = items[k];                 // a = group (0/1) or 2/3 = product low/high part, b = first sub-state
            const SpaceDev& sp = spaces[it.space];
            pmode = it.a >= 2 ? (int)it.a - 1 : 0;
            const int g = pmode ? 0 : (int)it.a;
            const bool joint = sp.kind == K_JOINT;
            const int K1 = sp.splitA;
            KG = pmode == 1 ? K1 : pmode == 2 ? sp.KA - K1 : (g ? sp.KB : sp.KA);
            const uint32_t NG = 1u << KG;
            ev = (g ? sp.evB : sp.evA) + (pmode == 2 ? K1 : 0);
            const Side sd = side_of(sp, g, S);
            const double* ptab = pmode == 2 ? sd.t + ((uint64_t)NR << K1) : sd.t;     // T1 or T2
            const double* pH = S + sp.stP + (pmode == 2 ? (uint64_t)sp.slices * (NR + sp.KA) * (1u << K1) : 0);
            const int psl = pmode == 1 ? (int)sp.slices : pmode == 2 ? (int)max(1u, (1u << K1) >> 7) : 1;   // hi slices / lo blocks of k_pf
            const uint64_t pslice = (uint64_t)(NR + sp.KA) * NG;
            const double* st = S + (g ? sp.stB : sp.stA);
            const double* y = S + sp.y_off;
            const double* x = S + sp.x_off;
            const int n_tot = sp.n_tot;
            n = n_tot - 1;
            int nrows = n;
            switch (sp.kind) {
                case K_PRE:   nrows = n_tot; accid = 0; break;
                case K_JOINT: nrows = n; accid = 0; always_n = (g == 1); pseudo_tot = true; break;
                case K_PF:    nrows = n; accid = 2; always_n = true; break;
                case K_MF:    nrows = n; accid = 1; always_n = true; break;
                case K_S1:    nrows = n_tot; accid = 1; break;
                default:      nrows = n_tot; accid = 0; break;
            }
            const bool has_pseudo = (sp.kind == K_PRE || sp.kind == K_JOINT || sp.kind == K_S2);
            is_row = lane < nrows;
            is_pseudo = (lane == ROW_DP || lane == ROW_DM) && has_pseudo;
            wgt = sp.cls ? w_other : w_type0;
            // bit of row `lane` inside the group (table mode) / inside the whole space (product mode)
            int abit = -1;
            if (pmode) { for (int b = 0; b < sp.KA; ++b) if (sp.evA[b] == lane) abit = b; }
            else       { for (int b = 0; b < KG; ++b) if (ev[b] == lane) abit = b; }
            const bool pre_seed = sp.kind == K_PRE;
            const SpaceDev& jsp = spaces[pre_seed ? sp.joint : it.space];
            const uint32_t u_end = min(NG, it.b + (it.c ? it.c : (uint32_t)FIN_U));
            for (uint32_t u0 = it.b; u0 < u_end; u0 += 32) {
                const uint32_t u = u0 + lane;
                const bool uv = u < u_end;
                // ---- phase 1: lane = sub-state ----
                double gg = 0.0, xv = 0.0, yv = 0.0;
                if (uv && !pmode) { if (joint) gg = st[u]; else { xv = x[u]; yv = y[u]; gg = xv * yv; } }
                for (int r = 0; r < NR; ++r) {
                    const int rb = __shfl_sync(0xffffffffu, abit, r);
                    const bool rrow = r < nrows, rps = (r == ROW_DP || r == ROW_DM) && has_pseudo;
                    double wv = 0.0;
                    if (uv && (rrow || rps)) {
                        if (pmode) {                              // weighted marginals from k_pf_lo / k_pf_hi
                            double hsum = 0.0;
                            const double* ph = pH + (uint64_t)r * NG + u;
                            int q = 0;
                            for (; q + 4 <= psl; q += 4) {            // four loads in flight, added in slice order
                                const double h0 = ph[q * pslice], h1 = ph[(q + 1) * pslice], h2 = ph[(q + 2) * pslice], h3 = ph[(q + 3) * pslice];
                                hsum += h0; hsum += h1; hsum += h2; hsum += h3;
                            }
                            for (; q < psl; ++q) hsum += ph[q * pslice];
                            wv = ptab[(uint64_t)r * NG + u] * hsum;
                        } else {
                            const double R = rrow ? sd.rate(r, u) : sd.special(r, u);
                            double E = -gg;
                            if (rrow && rb >= 0) {
                                if ((u >> rb) & 1u) E = 0.0;
                                else if (joint) E = st[(uint64_t)(1 + rb) * NG + u] - gg;
                                else E = yv * (x[u | (1u << rb)] - xv);
                            }
                            wv = R * E;
                            // the seeding edge of the pre-seeding lattice ends in the joint lattice
                            if (pre_seed && r == n) wv += R * yv * S[jsp.x_off + (((uint64_t)u << jsp.KA) | u)];
                        }
                    }
                    tile[r * 33 + lane] = wv;
                }
                __syncwarp();
                // ---- phase 2: lane = row ----
                double tsum = 0.0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double wv = tile[lane * 33 + j];
                    tsum += wv;
#pragma unroll
                    for (int b = 0; b < 5; ++b) if ((j >> b) & 1) ac[b] += wv;
                }
                tot += tsum;
#pragma unroll
                for (int b = 5; b < MB; ++b) if (b < KG && ((u0 >> b) & 1u)) ac[b] += tsum;
                __syncwarp();
            }
        }
        // ---- flush into the CTA accumulator, warp after warp ----
        for (int q = 0; q < FIN_WARPS; ++q) {
            if (q == w && have && (is_row || is_pseudo)) {
                double* row = G + ((size_t)accid * NR + lane) * NR;
#pragma unroll
                for (int b = 0; b < MB; ++b) if (b < KG) row[ev[b]] += wgt * ac[b];
                if (pmode != 2) {                               // totals are counted once (low part)
                    if (is_row) { row[lane] += wgt * tot; if (always_n) row[n] += wgt * tot; }
                    else if (pseudo_tot) row[n] += wgt * tot;
                }
            }
            __syncthreads();
        }
    }
    double* out = partial + (size_t)blockIdx.x * NACC * NR * NR;
    for (int t = threadIdx.x; t < NACC * NR * NR; t += blockDim.x) out[t] += G[t];     // same CTA index, chunk after chunk
}

// first stage of the partials reduction: slice y of the CTA range, fixed order inside a slice
__global__ void k_reduce_partials(const double* __restrict__ partial, int n_cta, double* __restrict__ out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = (n_cta + gridDim.y - 1) / gridDim.y;
    const int c0 = blockIdx.y * per, c1 = min(n_cta, c0 + per);
    double s = 0.0;
    for (int c = c0; c < c1; ++c) s += partial[(size_t)c * NACC * NR * NR + e];
    out[(size_t)blockIdx.y * NACC * NR * NR + e] = s;
}

// ------------------------------------------------------------------------------------------
// Reduce the per-CTA partials, map the three effective-parameter gradients back to
// (log_theta, log_d_p, log_d_m) and add the weighted log-likelihood sum.
//   theta_pt/d_p space (likelihood.py:457-461, 605-609): theta_ij += G1 except (i<n, j=n);  d_p[j] -= sum_{i!=j} G1[i][j]
//   theta/d_m space    (likelihood.py:565-566):          theta_ij += G2;                    d_m[j] -= sum_{i!=j} G2[i][j]
__global__ void k_final(const double* __restrict__ partial, int n_cta, const double* __restrict__ diracc, int n_dir,
                        const double* __restrict__ logp, const uint8_t* __restrict__ cls, int64_t n_dat,
                        const double* __restrict__ cnt_dm2, double w_type0, double w_other, int n_tot,
                        int want_grad, double* __restrict__ out)
{
    __shared__ double G[NACC][NR][NR];
    __shared__ double red[1024];
    const int n = n_tot - 1;
    if (want_grad) {
        for (int t = threadIdx.x; t < NACC * NR * NR; t += blockDim.x) {
            double s = 0.0;
            for (int c = 0; c < n_cta; ++c) s += partial[(size_t)c * NACC * NR * NR + t];
            (&G[0][0][0])[t] = s;
        }
    }
    double s = 0.0;
    for (int64_t p = threadIdx.x; p < n_dat; p += blockDim.x) {
        const uint8_t c = cls[p];
        if (c != 255) s += (c ? w_other : w_type0) * logp[p];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
    if (!want_grad) return;
    double* gth = out + 1;
    double* gdp = gth + n_tot * n_tot;
    double* gdm = gdp + n_tot;
    for (int t = threadIdx.x; t < n_tot * n_tot; t += blockDim.x) {
        const int i = t / n_tot, j = t % n_tot;
        double v = G[0][i][j] + G[2][i][j];
        if (!(i < n && j == n)) v += G[1][i][j];
        gth[t] = v;
    }
    for (int j = threadIdx.x; j < n_tot; j += blockDim.x) {
        double a = G[0][ROW_DP][j], b = G[0][ROW_DM][j] + w_other * cnt_dm2[j];
        for (int q = 0; q < n_dir; ++q) { a += diracc[q * 2 * NR + j]; b += diracc[q * 2 * NR + NR + j]; }
        for (int i = 0; i < n_tot; ++i) if (i != j) { a -= G[1][i][j]; b -= G[2][i][j]; }
        gdp[j] = a; gdm[j] = b;
    }
}

__global__ void k_fp64_peak(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

}  // namespace mmh
