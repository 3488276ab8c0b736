// Row-block resolvent solve of a pair lattice (big tier): forward substitution, adjoint substitution and the
// group-B marginal statistics with the column lattice kept in shared memory.
//
// A pair's operator is a Kronecker SUM: an edge on a row bit (group B, MT events) has ONE scalar rate per
// row, T_B[ev][uB], whatever the column.  So a whole row of the solution depends on earlier rows only through
//     rhs[uB, :] = b[uB, :] + sum_{b in uB} T_B[ev_b][uB - b] * y[uB - b, :]
// a sum of scaled, contiguous, finished rows -- a streaming read with perfectly coalesced 32-byte loads -- followed by
// a triangular solve over the column lattice alone (group A, shifted diagonal dA[uA] + dB[uB]).  A CTA works on blocks
// of 2^12 states = R rows x 2^KI columns (KI = min(KA, 12)), up to RB_MAXBLK blocks one after the other:
//   phase 1  thread = four consecutive columns: right-hand side and every OUTER edge (row bits; for KA > 12 also the
//            column bits >= 12, whose rates are vectors read from the plain table), result parked in shared memory
//   phase 2  the column lattice is solved inside shared memory, one __syncthreads per popcount level of the (KI-2)-bit
//            group index; thread = one group of four states (column bits 0,1 in registers); every operand of this
//            phase is in shared memory (values, product-form rate tables t1 (bits 0..5) x t2 (bits 6..11), the level
//            lists) except the diagonal vector dA (one 32-byte load per group, issued one level ahead)
//   phase 3  coalesced write of the block; the adjoint pass adds sum_uA x y here
// Launch levels run over the OUTER bits only (KB + max(0, KA-12) instead of KA-4+KB), and the L2 line reads per state
// fall from (K-4)/2 to (K-12)/2.  The adjoint pass reads, for every row bit b not in uB, the row x[uB + b, :] anyway:
// with the own row of y the dot products are the group-B statistics stB[1+b][uB] = sum_uA y[uB,uA] x[uB+b,uA]
// (k_solve_tile_adjb does the same per 16-column tile; here a CTA owns complete rows, so no partial tables per column
// chunk are needed unless KA > 12).
// Replaces the Jacobi sweeps of likelihood.py:231-262 over kronvec.py:259-539 for pairs with plain tables.
#pragma once

namespace mmh {

constexpr int RB_KI   = 12;                 // inner column bits
constexpr int RB_N    = 1 << RB_KI;         // states per block (32 KB)
constexpr int RB_T    = 256;                // threads per CTA
constexpr int RB_MINKA = 8;                 // at most 16 rows per block
constexpr int RB_MAXR = 1 << (RB_KI - RB_MINKA);
constexpr int RB_MAXE = RB_MAXR * 16;       // row-edge slots per block (rows x edges of a row)
constexpr int RB_MAXBLK = 4;                // blocks per CTA on fat levels
constexpr int RB_TABR = RB_KI * 64 + RB_KI * 16;   // doubles of the per-space table (k_rb_tables)
constexpr uint32_t RB_NONE = 0xffffffffu;

// Shared-memory layout.  A group of four consecutive states lives in two 16-byte halves, (v0,v1) in plane A and (v2,v3)
// in plane B: a quarter-warp's 16-byte accesses are free of bank conflicts when its eight group indices differ modulo 8,
// which the level lists (rb_level_lists) arrange; a 32-byte slot per group would collide two-way whatever the order.
struct RbSmem {
    double2  blkA[RB_N / 4], blkB[RB_N / 4];
    double2  t1A[RB_KI][16], t1B[RB_KI][16]; // rate factor over column bits 0..5 (with the base rate) = T_A[ev][l], split like blk
    double   t2[RB_KI][64];                 // rate factor over column bits 6..11
    double   ek[RB_MAXE];                   // outer row edges: scalar rate
    double   part[32][18];                  // adjoint: partial dot products per 128-state slab: [0] x.y, [1+q] edge q
    double   rdB[RB_MAXR];                  // dB[uB] of every row of the block
    double   fac[RB_KI];                    // KA > 12: rate factor of the block's outer column bits
    uint32_t es[RB_MAXE];                   //                  source row
    uint32_t ro[RB_MAXR];                   // outer index (uB << KOc | cH) of every row, RB_NONE = unused
    uint32_t rne[RB_MAXR];                  // number of outer row edges
    uint16_t lv[1 << (RB_KI - 2)];          // popcount-sorted group indices
    uint16_t lvoff[RB_KI];
};

// Per-space factors of the inner rate tables that do not exist in the plain table: t2[a][h] = product over the column
// bits 6..KI-1 set in h, fac[a][c] = product over the outer column bits (>= 12) set in c.
__global__ void k_rb_tables(const SpaceDev* __restrict__ spaces, const uint32_t* __restrict__ list,
                            const EvalPar* __restrict__ P, double* __restrict__ S)
{
    const SpaceDev& sp = spaces[list[blockIdx.x]];
    const int KA = sp.KA, KI = KA < RB_KI ? KA : RB_KI, KOc = KA - KI;
    double* T2 = S + sp.tabR;
    double* FAC = T2 + RB_KI * 64;
    for (int t = threadIdx.x; t < RB_KI * 64; t += blockDim.x) {
        const int a = t >> 6;
        const uint32_t h = (uint32_t)t & 63u;
        double r = 1.0;
        if (a < KI) {
            const int ev = sp.evA[a];
            for (int j = 6; j < KI; ++j)
                if ((h >> (j - 6)) & 1u) { const int e = sp.evA[j]; if (e != ev) r *= P->W[0][ev][e]; }
        }
        T2[t] = r;
    }
    for (int t = threadIdx.x; t < RB_KI * 16; t += blockDim.x) {
        const int a = t >> 4;
        const uint32_t c = (uint32_t)t & 15u;
        double r = 1.0;
        if (a < KI)
            for (int j = 0; j < KOc; ++j)
                if ((c >> j) & 1u) r *= P->W[0][sp.evA[a]][sp.evA[RB_KI + j]];
        FAC[t] = r;
    }
}

__device__ __forceinline__ void lds4(const double2* A, const double2* B, uint32_t g, double (&f)[4])
{
    const double2 a = A[g], b = B[g];
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void sts4(double2* A, double2* B, uint32_t g, const double (&f)[4])
{
    A[g] = make_double2(f[0], f[1]);
    B[g] = make_double2(f[2], f[3]);
}
__device__ __forceinline__ double warp_sum_rb(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// 1/d for d > 0 in the normal range (diagonal entries are sums of positive rates): MUFU.RCP64H seed (about 20 bits)
// and two Newton steps, 5 instructions instead of the ~25 of an IEEE division with its slow path; relative error below 1e-15
__device__ __forceinline__ double rcp_pos(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ double dot4(const double (&a)[4], const double (&b)[4])
{
    return fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])));
}

#ifdef RB_TIMING
__device__ unsigned long long rb_timing[16];
#define RB_TICK(i) do { __syncthreads(); if (tid == 0) { const long long t_ = clock64(); atomicAdd(&rb_timing[i], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
#else
#define RB_TICK(i) do {} while (0)
#endif

// what a phase-2 unit needs besides the values in shared memory; prepared one level ahead
struct RbUnit {
    uint32_t sidx, gc;
    double d4[4], db;
    bool on;
};

#ifndef RB_CTAS_ADJ
#define RB_CTAS_ADJ 3
#endif
#define RB_CTAS_OF(adj) ((adj) ? RB_CTAS_ADJ : 4)
// item: space, a = outer level | rows << 8, b = first position inside the level; the rows are taken R at a time
template <bool ADJ>
__global__ void __launch_bounds__(RB_T, RB_CTAS_OF(ADJ))
k_solve_rb(const SpaceDev* __restrict__ spaces, const Item* __restrict__ items, const uint32_t* __restrict__ hs,
           const uint32_t* __restrict__ hsidx, const uint16_t* __restrict__ rblv, double* __restrict__ S)
{
    extern __shared__ __align__(16) unsigned char rb_raw[];
    RbSmem& sm = *reinterpret_cast<RbSmem*>(rb_raw);
    const Item it = items[blockIdx.x];
    const SpaceDev& sp = spaces[it.space];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int KA = sp.KA, KB = sp.KB;
    const int KI = KA < RB_KI ? KA : RB_KI;
    const int KOc = KA - KI, KO = KB + KOc;
    const int lR = RB_KI - KI;                               // log2 of the rows per block
    const uint32_t R = 1u << lR;
    const uint32_t lvl = it.a & 255u, rows_all = it.a >> 8;
    const uint32_t NI = 1u << KI, mC = (1u << KOc) - 1u;
    const uint32_t NB = 1u << KB;
    const int KG = KI - 2;                                   // bits of the group index inside a row
    double* v = S + (ADJ ? sp.x_off : sp.y_off);
    const double* yv = S + sp.y_off;
    const double* tabA = S + sp.tabA;
    const double* tabB = S + sp.tabB;
    const double* tabR = S + sp.tabR;
    const double* dA = tabA + ((uint64_t)ROW_D << KA);
    const uint32_t* lvl_rows = hs + hsidx[KO * 32 + lvl] + it.b;

#ifdef RB_TIMING
    long long t_prev = clock64();
    if (tid == 0) atomicAdd(&rb_timing[8 + (ADJ ? 1 : 0)], 1ull);
#endif
    // ---- per-CTA tables (plain copies) ----------------------------------------------------------------------------
    for (uint32_t t = tid; t < (1u << KG); t += RB_T) sm.lv[t] = rblv[(KG << (RB_KI - 2)) + t];
    if (tid <= KG + 1) sm.lvoff[tid] = (uint16_t)(hsidx[KG * 32 + tid] - hsidx[KG * 32]);
    for (int t = tid; t < KI * 64; t += RB_T) {
        const int a = t >> 6;
        const uint32_t l = (uint32_t)t & 63u;
        const double r = tabA[((uint64_t)sp.evA[a] << KA) + l];         // NI >= 256 > 63
        double2* q = (l & 2u) ? &sm.t1B[a][l >> 2] : &sm.t1A[a][l >> 2];
        if (l & 1u) q->y = r; else q->x = r;
        sm.t2[a][l] = tabR[t];
    }

#pragma unroll 1
    for (uint32_t row0 = 0; row0 < rows_all; row0 += R) {
        const uint32_t cnt = min(R, rows_all - row0);
        __syncthreads();                                     // the previous block is written out; (first) the tables are in place
        RB_TICK(0);
        // ---- block context: rows, their diagonal parts and their outer row edges ---------------------------------
        {
            const uint32_t r = (uint32_t)tid >> 4;
            const int b = tid & 15;
            if (r < R) {
                uint32_t o = RB_NONE;
                if (r < cnt) o = lvl_rows[row0 + r];
                const uint32_t uB = o >> KOc;
                const uint32_t rel = ADJ ? (~uB & (NB - 1u)) : uB;
                if (b == 0) {
                    sm.ro[r] = o;
                    sm.rne[r] = o == RB_NONE ? 0u : (uint32_t)__popc(rel);
                    sm.rdB[r] = o == RB_NONE ? 0.0 : tabB[((uint64_t)ROW_D << KB) + uB];
                }
                if (o != RB_NONE && b < KB && ((rel >> b) & 1u)) {
                    const uint32_t q = (uint32_t)__popc(rel & ((1u << b) - 1u));
                    const uint32_t orow = uB ^ (1u << b);
                    sm.es[r * 16u + q] = orow;
                    sm.ek[r * 16u + q] = tabB[((uint64_t)sp.evB[b] << KB) + (ADJ ? uB : orow)];
                }
            }
            if (KOc && tid >= RB_T - RB_KI) {                // blocks with outer column bits hold one row
                const int a = tid - (RB_T - RB_KI);
                sm.fac[a] = tabR[RB_KI * 64 + a * 16 + (lvl_rows[row0] & mC)];
            }
        }
        __syncthreads();
        RB_TICK(1);

        // ---- phase 1: right-hand side and outer edges, G groups of four states per thread and round ------------------
        constexpr int G = ADJ ? 1 : 2, NBE = ADJ ? 3 : 2;    // the adjoint also holds the own row of y
#pragma unroll 1
        for (int p = 0; p < RB_N / (4 * G * RB_T); ++p) {
            uint32_t idx[G], uA0[G], ne[G], e0[G];
            uint64_t s0[G];
            bool on[G];
            double acc[G][4], y4[G][4];
            uint32_t nemax = 0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                idx[g] = ((uint32_t)(G * p + g) * RB_T + (uint32_t)tid) << 2;
                const uint32_t r = idx[g] >> KI, lo = idx[g] & (NI - 1u);
                const uint32_t o = sm.ro[r];
                on[g] = o != RB_NONE;                        // uniform over the warp (KI >= 8)
                const uint32_t uB = o >> KOc, cH = o & mC;
                uA0[g] = (cH << RB_KI) | lo;
                s0[g] = ((uint64_t)uB << KA) | uA0[g];
                ne[g] = sm.rne[r]; e0[g] = r * 16u;
                nemax = max(nemax, ne[g]);
#pragma unroll
                for (int t = 0; t < 4; ++t) { acc[g][t] = 0.0; y4[g][t] = 0.0; }
                if (on[g]) {
                    if (ADJ) ld4(yv + s0[g], y4[g]);
                    tile_rhs<ADJ>(sp, spaces, S, KA, KB, uB, uA0[g], acc[g]);
                }
            }
            for (uint32_t q0 = 0; q0 < nemax; q0 += NBE) {
                double ys[G][NBE][4], k[G][NBE];
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int e = 0; e < NBE; ++e) {
                        if (q0 + e < ne[g]) {
                            k[g][e] = sm.ek[e0[g] + q0 + e];
                            ld4(v + (((uint64_t)sm.es[e0[g] + q0 + e] << KA) | uA0[g]), ys[g][e]);
                        } else {
                            k[g][e] = 0.0;
#pragma unroll
                            for (int t = 0; t < 4; ++t) ys[g][e][t] = 0.0;
                        }
                    }
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int e = 0; e < NBE; ++e) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) acc[g][t] = fma(k[g][e], ys[g][e][t], acc[g][t]);
                        if (ADJ) {
                            const double d = warp_sum_rb(dot4(y4[g], ys[g][e]));
                            if (lane == 0 && q0 + e < ne[g]) sm.part[(G * p + g) * (RB_T / 32) + w][1 + q0 + e] = d;
                        }
                    }
            }
            // outer column bits (KA > 12): vector rates from the plain table
            if (KOc) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (!on[g]) continue;
                    const uint32_t cH = uA0[g] >> RB_KI;
                    uint32_t m = ADJ ? (~cH & mC) : cH;
                    while (m) {
                        const int a = RB_KI + __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t bit = 1u << a;
                        double rv[4], ys[4];
                        ld4(tabA + ((uint64_t)sp.evA[a] << KA) + (ADJ ? uA0[g] : (uA0[g] ^ bit)), rv);
                        ld4(v + (s0[g] ^ bit), ys);
#pragma unroll
                        for (int t = 0; t < 4; ++t) acc[g][t] = fma(rv[t], ys[t], acc[g][t]);
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) sts4(sm.blkA, sm.blkB, idx[g] >> 2, acc[g]);
        }

        // ---- phase 2: column lattice inside shared memory ------------------------------------------------------------
        // unit u of a level = (row u / nl, entry u % nl of the level's list); the first unit of a thread is prepared
        // (indices, diagonal load) before the barrier that ends the previous level
        auto prepare = [&](int l, uint32_t u, RbUnit& q) {
            const uint32_t off = sm.lvoff[l], nl = (uint32_t)sm.lvoff[l + 1] - off;
            q.on = false;
            if (u >= (nl << lR)) return;
            const uint32_t r = lR ? u / nl : 0u, i = u - r * nl;     // consecutive lanes = consecutive list entries of a row
            const uint32_t o = sm.ro[r];
            if (o == RB_NONE) return;
            q.on = true;
            q.gc = sm.lv[off + i];
            q.sidx = (r << KG) | q.gc;                               // group index inside the block
            q.db = sm.rdB[r];
            ld4(dA + (((o & mC) << RB_KI) | (q.gc << 2)), q.d4);
        };
        auto solve = [&](const RbUnit& q) {
            const uint32_t gc = q.gc, sidx = q.sidx;
            double acc[4];
            lds4(sm.blkA, sm.blkB, sidx, acc);
            uint32_t m = ADJ ? (~gc & ((1u << KG) - 1u)) : gc;
            while (m) {                                                  // two edges per round, loads first
                double y4[2][4], r4[2][4], k[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool live = m != 0u;
                    const int b = live ? __ffs(m) - 1 : 0;
                    m &= m - 1;
                    const uint32_t us = ADJ ? gc : (gc ^ (1u << b));     // group whose rates apply
                    lds4(sm.blkA, sm.blkB, sidx ^ (1u << b), y4[e]);
                    lds4(sm.t1A[b + 2], sm.t1B[b + 2], us & 15u, r4[e]);
                    k[e] = live ? sm.t2[b + 2][us >> 4] : 0.0;
                    if (KOc) k[e] *= sm.fac[b + 2];
                }
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[t] = fma(r4[e][t] * k[e], y4[e][t], acc[t]);
            }
            double k0 = sm.t2[0][gc >> 4], k1 = sm.t2[1][gc >> 4];
            if (KOc) { k0 *= sm.fac[0]; k1 *= sm.fac[1]; }
            const double2 q1 = sm.t1A[1][gc & 15u];
            const double e0a = sm.t1A[0][gc & 15u].x * k0, e0b = sm.t1B[0][gc & 15u].x * k0;   // bit 0: 0 -> 1, 2 -> 3
            const double e1a = q1.x * k1, e1b = q1.y * k1;                                      // bit 1: 0 -> 2, 1 -> 3
            double inv[4], val[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) inv[t] = rcp_pos(q.d4[t] + q.db);
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
            sts4(sm.blkA, sm.blkB, sidx, val);
        };
        RbUnit nxt;
        prepare(ADJ ? KG : 0, (uint32_t)tid, nxt);
        __syncthreads();
        RB_TICK(2);
#pragma unroll 1
        for (int step = 0; step <= KG; ++step) {
            const int l = ADJ ? KG - step : step;
            const RbUnit cur = nxt;
#ifdef RB_TIMING
            const long long tl0 = clock64();
#endif
            if (step < KG) prepare(ADJ ? l - 1 : l + 1, (uint32_t)tid, nxt);   // its diagonal load flies during this level
#ifdef RB_TIMING
            const long long tl1 = clock64();
#endif
            if (cur.on) solve(cur);
#ifdef RB_TIMING
            const long long tl2 = clock64();
            if (tid == 0) { atomicAdd(&rb_timing[12], (unsigned long long)(tl1 - tl0)); atomicAdd(&rb_timing[13], (unsigned long long)(tl2 - tl1)); }
            if (tid == 255) { atomicAdd(&rb_timing[14], (unsigned long long)(tl1 - tl0)); atomicAdd(&rb_timing[15], (unsigned long long)(tl2 - tl1)); }
#endif
            const uint32_t units = ((uint32_t)sm.lvoff[l + 1] - sm.lvoff[l]) << lR;
            for (uint32_t u = tid + RB_T; u < units; u += RB_T) {      // levels with more units than threads
                RbUnit q;
                prepare(l, u, q);
                if (q.on) solve(q);
            }
            __syncthreads();
        }

        RB_TICK(3);
        // ---- phase 3: coalesced write of the block; adjoint: sum_uA x y ----------------------------------------------
#pragma unroll 1
        for (int p = 0; p < RB_N / (4 * RB_T); ++p) {
            const uint32_t idx = ((uint32_t)p * RB_T + (uint32_t)tid) << 2;
            const uint32_t r = idx >> KI, lo = idx & (NI - 1u);
            const uint32_t o = sm.ro[r];
            if (o == RB_NONE) continue;
            const uint64_t s0 = ((uint64_t)(o >> KOc) << KA) | ((o & mC) << RB_KI) | lo;
            double val[4];
            lds4(sm.blkA, sm.blkB, idx >> 2, val);
            st4(v + s0, val[0], val[1], val[2], val[3]);
            if (ADJ) {
                double y4[4];
                ld4(yv + s0, y4);
                const double g = warp_sum_rb(dot4(y4, val));
                if (lane == 0) sm.part[p * (RB_T / 32) + w][0] = g;
            }
        }
        RB_TICK(4);
        if (ADJ) {
            __syncthreads();
            // group-B statistics of the block's rows: slabs of a row are added in a fixed order; one partial table per
            // value of the outer column bits (k_stats_reduce adds them)
            const int spr = KI - 7;                                  // log2 slabs per row
            for (uint32_t t = tid; t < cnt * (uint32_t)(KB + 1); t += RB_T) {
                const uint32_t r = t / (uint32_t)(KB + 1), e = t - r * (uint32_t)(KB + 1);
                const uint32_t o = sm.ro[r];
                const uint32_t uB = o >> KOc;
                double s = 0.0;
                int col = -1;
                if (e == 0) col = 0;
                else if (!((uB >> (e - 1)) & 1u)) col = 1 + __popc(~uB & ((1u << (e - 1)) - 1u));
                if (col >= 0)
                    for (uint32_t j = r << spr; j < ((r + 1u) << spr); ++j) s += sm.part[j][col];
                S[sp.stPB + (uint64_t)(o & mC) * (uint64_t)(KB + 1) * NB + (uint64_t)e * NB + uB] = s;
            }
        }
        RB_TICK(5);
#ifdef RB_TIMING
        if (tid == 0) atomicAdd(&rb_timing[10 + (ADJ ? 1 : 0)], 1ull);
#endif
    }
}

}  // namespace mmh
