// Blocked substitution solve of the big lattices (K >= 13 bits): warp-private shared-memory blocks, skewed wavefront,
// outer-bit sources streamed in by bulk asynchronous copies (TMA, cp.async.bulk + mbarrier).
//
// Replaces the (k+1)-sweep Jacobi iteration of the reference (metmhn/jx/likelihood.py:231-262 `R_i_inv_vec`,
// metmhn/jx/vanilla.py:269-305 `R_inv_vec`) and the per-event Kronecker shuffles behind it (kronvec.py:259-539) by ONE
// exact forward (or adjoint) substitution on the subset lattice, blocked for the memory system of a B200:
//
//   * the lattice index s has K bits; the low eight are "column" bits, four more bits are "sequence" bits, the K-12
//     others are "outer" bits.  A ROW is the 256 consecutive states of one value of the sequence + outer bits (2 KB),
//     a BLOCK the 16 rows of one value of the outer bits.  One WARP owns one block at a time and keeps its 4096
//     solved values in its private 32 KB of shared memory;
//   * every edge on an OUTER bit reads the value of the same position in a block finished by an earlier launch
//     (launches go level by level over the popcount of the outer bits only: K-11 launches instead of K-3), i.e.
//     (K-12)/2 global reads per state instead of (K-4)/2 -- the reads that bound the previous tile kernel
//     (profiles/r1_final_solve_tile_ncu_full.txt: 76 B of L2 traffic per state against 8 algorithmic).  Those source
//     rows are contiguous 2 KB pieces: one elected lane streams them into a small shared-memory ring with
//     cp.async.bulk (completion on an mbarrier), several rows ahead of their use, so the L2 / HBM latency is not on
//     the warp's critical path and no registers are spent on staging;
//   * inside the block nothing leaves the SM.  A lane owns eight states of a row (column bits 0, 6, 7 = register
//     bits; column bits 1..5 = the lane number, so that a 16-byte access of the warp covers 512 contiguous bytes in
//     shared AND global memory): sequence-bit edges read rows the lane itself finished earlier, lane-bit edges the
//     same row of a neighbouring lane, register-bit edges are register arithmetic;
//   * each iteration has two phases.  OUTER (all lanes on row t): right-hand side + the outer-bit edges of row t,
//     from the ring, accumulated into the row's shared-memory slot.  INNER, a SKEWED WAVEFRONT: lane l finishes row
//     t - popcount(l) (adjoint: mirrored).  A lane-bit neighbour l^a has one bit less, so it finished the same row one
//     iteration earlier: every lane is busy at every iteration (except 5 fill / drain iterations per block), there is
//     no __syncthreads and no shuffle, only __syncwarp.
//
// The per-lane functions are written against a small abstract description of the space (BlkCtx: per-bit rate
// descriptors, diagonal, right-hand side functor) and are plain C++: tests/host/blk_host_test.cpp compiles the very same
// functions for the host and emulates a warp lane by lane against a sequential substitution.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define MMH_HD __host__ __device__ __forceinline__
#else
#define MMH_HD inline
#endif

namespace mmh {

constexpr int BLK_MAXBITS = 26;
constexpr int BLK_CB = 8;                       // column bits
constexpr int BLK_SB = 3;                       // sequence bits
constexpr int BLK_Q = 1 << BLK_SB;              // rows of a block
constexpr int BLK_ROW = 1 << BLK_CB;            // doubles per row
constexpr int BLK_DOUBLES = BLK_Q * BLK_ROW;    // 4096 doubles = 32 KB per warp
constexpr int BLK_ITERS = BLK_Q + 5;            // skew: 5 fill / drain iterations
constexpr int BLK_NS = 4;                       // ring slots (one source row = 2 KB each)
constexpr uint32_t BLK_REGMASK = 0xC1u;         // register bits of a lane's eight states: positions 0, 6, 7
constexpr int BLK_MAXC = 15;                    // column-profile vectors of a space (256 doubles each, shared by the CTA)
constexpr int BLK_MAXKO = 23 - BLK_CB - BLK_SB; // outer bits (K <= 23)
// per-row scalars of a block: [0..7] column-bit edges, then the sequence-bit edges, the diagonal part, the outer edges
constexpr int BLK_SC_SEQ = BLK_CB, BLK_SC_D2 = BLK_CB + BLK_SB, BLK_SC_OUT = BLK_CB + BLK_SB + 1;
constexpr int BLK_SCW = BLK_SC_OUT + BLK_MAXKO;
constexpr int BLK_SC_DOUBLES = BLK_Q * BLK_SCW;

// Edge u -> u | (1 << t) of bit t (t not in u):  rate_t(u) = P[u & mP] * Q[(u >> shQ) & mQ]   (null pointer = 1).
// P is read two consecutive entries at a time (mP >= 1), Q must not depend on bit 0 (shQ >= 1).
struct BlkBit {
    const double* P;
    const double* Q;
    uint32_t mP, mQ, shQ;
    int32_t cidx;                               // its column-profile vector in the CTA's table, -1 = the rate does not depend on the column bits
};

struct BlkCtx {
    BlkBit bit[BLK_MAXBITS];                    // by bit position
    const double* d1;                           // diag(u) = d1[u & m1] + (d2 ? d2[(u >> sh2) & m2] : 0)
    const double* d2;
    uint32_t m1, m2, sh2;
    int K, KO;
    int nC;                                     // column-profile vectors in use
    int d1row;                                  // 1: d1 depends on the row of the block (read per row), 0: eight values per lane and block
    int d2mode;                                 // 0: no d2, 1: one scalar per row (sh2 >= 8), 2: depends on column bits (read per piece)
    int simple;                                 // 1: no sequence bit has a column profile, d1 per block, d2 per row (a pair whose columns are
                                                //    group-A bits and whose sequence bits are group-B bits): the fast instantiation
    uint8_t seq[BLK_SB];                        // positions of the sequence bits, ascending
    uint8_t out[BLK_MAXBITS];                   // positions of the outer bits, ascending
    uint32_t seqm[BLK_Q];                       // index offset of row q of a block (its sequence bits)
    uint8_t cbit[BLK_MAXC];                     // bit position of every column-profile vector
};

// ---- small helpers ------------------------------------------------------------------------------------------
MMH_HD int blk_popc(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
MMH_HD int blk_ffs(uint32_t v)                  // index of the lowest set bit (v != 0)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// state j (0..7) of a lane relative to its base state: bit 0 of j = column bit 0, bit 1 = column bit 6, bit 2 = column bit 7
MMH_HD uint32_t blk_joff(int j) { return (uint32_t)(j & 1) | ((uint32_t)(j >> 1) << 6); }

struct blk_d2 { double x, y; };
MMH_HD void blk_ldg2(const double* __restrict__ p, double& a, double& b)      // 16 bytes, 16-byte aligned, global
{
#if defined(__CUDA_ARCH__)
    asm("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
#else
    a = p[0]; b = p[1];
#endif
}
// eight values of a lane from a row in global memory (piece pc at +64 doubles): four coalesced 16-byte accesses
MMH_HD void blk_ldg8(const double* __restrict__ p, double (&v)[8])
{
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) blk_ldg2(p + pc * 64, v[2 * pc], v[2 * pc + 1]);
}
MMH_HD void blk_stg8(double* __restrict__ p, const double (&v)[8])
{
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
#if defined(__CUDA_ARCH__)
        asm volatile("st.global.v2.f64 [%2], {%0,%1};" :: "d"(v[2 * pc]), "d"(v[2 * pc + 1]), "l"(p + pc * 64) : "memory");
#else
        p[pc * 64] = v[2 * pc]; p[pc * 64 + 1] = v[2 * pc + 1];
#endif
    }
}
// the same from / to a row in shared memory (natural layout: double index = column); p = row + lane * 2
MMH_HD void blk_lds8(const double* p, double (&v)[8])
{
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
#if defined(__CUDA_ARCH__)
        const double2 t = *reinterpret_cast<const double2*>(p + pc * 64);
        v[2 * pc] = t.x; v[2 * pc + 1] = t.y;
#else
        v[2 * pc] = p[pc * 64]; v[2 * pc + 1] = p[pc * 64 + 1];
#endif
    }
}
MMH_HD void blk_sts8(double* p, const double (&v)[8])
{
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<double2*>(p + pc * 64) = make_double2(v[2 * pc], v[2 * pc + 1]);
#else
        p[pc * 64] = v[2 * pc]; p[pc * 64 + 1] = v[2 * pc + 1];
#endif
    }
}

// global offset of row q of a block (its sequence bits) and of the outer coordinate o
MMH_HD uint32_t blk_seq_mask(const BlkCtx& c, uint32_t q) { return c.seqm[q]; }
MMH_HD uint32_t blk_seq_mask_slow(const BlkCtx& c, uint32_t q)
{
    uint32_t m = 0;
    for (int i = 0; i < BLK_SB; ++i) m |= ((q >> i) & 1u) << c.seq[i];
    return m;
}
// rate of the edge of bit b that starts at state u
MMH_HD double blk_rate1(const BlkBit& b, uint32_t u)
{
    double r = 1.0;
    if (b.P) r = b.P[u & b.mP];
    if (b.Q) r *= b.Q[(u >> b.shQ) & b.mQ];
    return r;
}
MMH_HD uint32_t blk_outer_mask(const BlkCtx& c, uint32_t o)
{
    uint32_t m = 0;
    for (int i = 0; i < c.KO; ++i) m |= ((o >> i) & 1u) << c.out[i];
    return m;
}

// Rates of the edges of bit `b` that START at the lane's eight states u0 + blk_joff(j) (u0 even).
MMH_HD void blk_rate8(const BlkBit& b, uint32_t u0, double (&r)[8])
{
    if (b.shQ >= BLK_CB || !b.Q) {              // the scalar factor is the same for the four pieces
        const double q = b.Q ? b.Q[(u0 >> b.shQ) & b.mQ] : 1.0;
        if (b.P) {
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                blk_ldg2(b.P + ((u0 + pc * 64) & b.mP), r[2 * pc], r[2 * pc + 1]);
                r[2 * pc] *= q; r[2 * pc + 1] *= q;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = q;
        }
    } else {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
            const uint32_t up = u0 + pc * 64;
            const double q = b.Q[(up >> b.shQ) & b.mQ];
            if (b.P) {
                blk_ldg2(b.P + (up & b.mP), r[2 * pc], r[2 * pc + 1]);
                r[2 * pc] *= q; r[2 * pc + 1] *= q;
            } else { r[2 * pc] = q; r[2 * pc + 1] = q; }
        }
    }
}

// eight reciprocals with one division (products of positive, well-scaled diagonal entries)
MMH_HD void blk_inv8(const double (&d)[8], double (&inv)[8])
{
    const double p01 = d[0] * d[1], p23 = d[2] * d[3], p45 = d[4] * d[5], p67 = d[6] * d[7];
    const double p0123 = p01 * p23, p4567 = p45 * p67;
    const double r = 1.0 / (p0123 * p4567);
    const double r0123 = r * p4567, r4567 = r * p0123;
    const double r01 = r0123 * p23, r23 = r0123 * p01, r45 = r4567 * p67, r67 = r4567 * p45;
    inv[0] = r01 * d[1]; inv[1] = r01 * d[0]; inv[2] = r23 * d[3]; inv[3] = r23 * d[2];
    inv[4] = r45 * d[5]; inv[5] = r45 * d[4]; inv[6] = r67 * d[7]; inv[7] = r67 * d[6];
}

// ---- per-lane constants of a space and per-lane state of a block -----------------------------------------
// Column-bit edges: rate_c(u) = profile_c(columns of u) * scalar_c(u with the column bits cleared), profile_c(col) =
// rate_c(col) / rate_c(0).  The profile values of the lane's eight states are constants of the SPACE (registers, set
// once per CTA); the scalar is a per-row entry of the block's table.  Forward: edges ENDING in the lane's states;
// adjoint: edges STARTING there.
struct BlkLane {
    double Rr[3][4];                            // register bits (positions 0, 6, 7): edge k of bit b, in the order of the source states lacking b
    double RL[5][8];                            // lane bits (positions 1..5); 0 where the lane has no such edge
    uint32_t base;                              // per block: outer bits of the block | lane << 1
    double d1v[8];                              // per block: d1 part of the diagonal of the lane's states (when the same for every row)
};

template <bool ADJ>
MMH_HD void blk_lane_consts(const BlkCtx& c, int lane, BlkLane& L)
{
    const uint32_t u0 = (uint32_t)lane << 1;
    {
        double r[8];
        double n = 1.0 / blk_rate1(c.bit[0], 0u);
        blk_rate8(c.bit[0], u0, r);
        L.Rr[0][0] = r[0] * n; L.Rr[0][1] = r[2] * n; L.Rr[0][2] = r[4] * n; L.Rr[0][3] = r[6] * n;      // j: 0->1 2->3 4->5 6->7
        n = 1.0 / blk_rate1(c.bit[6], 0u);
        blk_rate8(c.bit[6], u0, r);
        L.Rr[1][0] = r[0] * n; L.Rr[1][1] = r[1] * n; L.Rr[1][2] = r[4] * n; L.Rr[1][3] = r[5] * n;      // j: 0->2 1->3 4->6 5->7
        n = 1.0 / blk_rate1(c.bit[7], 0u);
        blk_rate8(c.bit[7], u0, r);
        L.Rr[2][0] = r[0] * n; L.Rr[2][1] = r[1] * n; L.Rr[2][2] = r[2] * n; L.Rr[2][3] = r[3] * n;      // j: 0->4 1->5 2->6 3->7
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const uint32_t bit = 2u << a;
        const bool has = (u0 & bit) != 0u;
        const bool edge = ADJ ? !has : has;      // forward: the edge comes from l ^ a (which lacks the bit); adjoint: it goes there
        if (edge) {
            const double n = 1.0 / blk_rate1(c.bit[1 + a], 0u);
            blk_rate8(c.bit[1 + a], u0 & ~bit, L.RL[a]);
#pragma unroll
            for (int j = 0; j < 8; ++j) L.RL[a][j] *= n;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) L.RL[a][j] = 0.0;
        }
    }
}

MMH_HD void blk_lane_block(const BlkCtx& c, uint32_t outer_mask, int lane, BlkLane& L)
{
    const uint32_t u0 = outer_mask | ((uint32_t)lane << 1);
    L.base = u0;
    if (!c.d1row) {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) blk_ldg2(c.d1 + ((u0 + pc * 64) & c.m1), L.d1v[2 * pc], L.d1v[2 * pc + 1]);
    }
}

// OUTER phase, one edge: acc += rate * (source row in the ring slot), rate = column profile (shared memory, or 1) times
// the scalar of (row, edge) from the block's table.
MMH_HD void blk_outer_edge(int lane, double sc, const double* cv, const double* slot, double (&acc)[8])
{
    double y[8];
    blk_lds8(slot + lane * 2, y);
    if (cv) {
        double r[8];
        blk_lds8(cv + lane * 2, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(r[j] * sc, y[j], acc[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(sc, y[j], acc[j]);
    }
}

// INNER phase of one lane at iteration t: finish row t - popcount(lane) (adjoint: mirrored).  `sm_rd` / `sm_wr` are the
// block's shared memory (the same pointer on the device; the host emulation reads from a snapshot of the iteration start).
// Straight-line code: an edge that does not exist for this lane / row has rate 0 and reads a finite value (the block's
// memory never holds anything but zeros, partial sums and solved values).
template <bool ADJ, bool SIMPLE>
MMH_HD void blk_inner(const BlkCtx& c, const BlkLane& L, int lane, int t, double* __restrict__ v,
                      const double* sm_rd, double* sm_wr, const double* sc, const double* ctab)
{
    const int pl = blk_popc((uint32_t)lane);
    const int qi = t - (ADJ ? 5 - pl : pl);
    if (qi < 0 || qi >= BLK_Q) return;
    const int q = ADJ ? BLK_Q - 1 - qi : qi;
    const uint32_t s0 = L.base | blk_seq_mask(c, (uint32_t)q);
    const double* scq = sc + q * BLK_SCW;
    // the diagonal first: when it has to come from global memory its latency hides behind the edges below
    double d[8];
    if (!SIMPLE && c.d1row) {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) blk_ldg2(c.d1 + ((s0 + pc * 64) & c.m1), d[2 * pc], d[2 * pc + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = L.d1v[j];
    }
    double acc[8];
    blk_lds8(sm_rd + q * BLK_ROW + lane * 2, acc);           // right-hand side + outer edges (OUTER phase, own columns)
    // ---- sequence bits: rows of this block the lane finished at earlier iterations ----
#pragma unroll
    for (int i = 0; i < BLK_SB; ++i) {
        const int ci = SIMPLE ? -1 : c.bit[c.seq[i]].cidx;
        blk_outer_edge(lane, scq[BLK_SC_SEQ + i], ci >= 0 ? ctab + ci * BLK_ROW : nullptr, sm_rd + (q ^ (1 << i)) * BLK_ROW, acc);
    }
    // ---- lane bits: the same row of the neighbouring lanes (finished one iteration earlier) ----
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        double y[8];
        blk_lds8(sm_rd + q * BLK_ROW + (lane ^ (1 << a)) * 2, y);
        const double k = scq[1 + a];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(L.RL[a][j], y[j] * k, acc[j]);
    }
    // ---- diagonal ----
    double inv[8];
    {
        if (SIMPLE || c.d2mode == 1) {
            const double k = scq[BLK_SC_D2];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] += k;
        } else if (c.d2mode == 2) {
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                const double k = c.d2[((s0 + pc * 64) >> c.sh2) & c.m2];
                d[2 * pc] += k; d[2 * pc + 1] += k;
            }
        }
        blk_inv8(d, inv);
    }
    // ---- register bits (positions 0, 6, 7 <-> bits 0, 1, 2 of j) ----
    const double k0 = scq[0], k1 = scq[6], k2 = scq[7];
    double e0[4], e1[4], e2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { e0[k] = L.Rr[0][k] * k0; e1[k] = L.Rr[1][k] * k1; e2[k] = L.Rr[2][k] * k2; }
    double y[8];
    if (!ADJ) {
        y[0] = acc[0] * inv[0];
        y[1] = fma(e0[0], y[0], acc[1]) * inv[1];
        y[2] = fma(e1[0], y[0], acc[2]) * inv[2];
        y[4] = fma(e2[0], y[0], acc[4]) * inv[4];
        y[3] = fma(e0[1], y[2], fma(e1[1], y[1], acc[3])) * inv[3];
        y[5] = fma(e0[2], y[4], fma(e2[1], y[1], acc[5])) * inv[5];
        y[6] = fma(e1[2], y[4], fma(e2[2], y[2], acc[6])) * inv[6];
        y[7] = fma(e0[3], y[6], fma(e1[3], y[5], fma(e2[3], y[3], acc[7]))) * inv[7];
    } else {
        y[7] = acc[7] * inv[7];
        y[6] = fma(e0[3], y[7], acc[6]) * inv[6];
        y[5] = fma(e1[3], y[7], acc[5]) * inv[5];
        y[3] = fma(e2[3], y[7], acc[3]) * inv[3];
        y[4] = fma(e0[2], y[5], fma(e1[2], y[6], acc[4])) * inv[4];
        y[2] = fma(e0[1], y[3], fma(e2[2], y[6], acc[2])) * inv[2];
        y[1] = fma(e1[1], y[3], fma(e2[1], y[5], acc[1])) * inv[1];
        y[0] = fma(e0[0], y[1], fma(e1[0], y[2], fma(e2[0], y[4], acc[0]))) * inv[0];
    }
    blk_sts8(sm_wr + q * BLK_ROW + lane * 2, y);
    blk_stg8(v + s0, y);
}

// Sequence bits = the first four positions of pref, pref+1, ..., K-1, 8, 9, ... (a pair whose column bits all belong to
// group A prefers group-B bits: then no column-bit rate depends on the row of the block); outer bits = the rest.
MMH_HD void blk_ctx_layout(BlkCtx& c, int K, int pref)
{
    c.K = K; c.KO = K - BLK_CB - BLK_SB;
    uint32_t used = 0;
    int p = (pref >= BLK_CB && pref < K) ? pref : BLK_CB;
    for (int i = 0; i < BLK_SB; ++i) {
        used |= 1u << p;
        p = (p + 1 < K) ? p + 1 : BLK_CB;
    }
    int ns = 0, no = 0;
    for (int t = BLK_CB; t < K; ++t) {
        if ((used >> t) & 1u) c.seq[ns++] = (uint8_t)t;
        else c.out[no++] = (uint8_t)t;
    }
    for (int q = 0; q < BLK_Q; ++q) c.seqm[q] = blk_seq_mask_slow(c, (uint32_t)q);
    // column-profile vectors: every non-column bit whose rate depends on the column bits
    c.nC = 0;
    for (int t = 0; t < K; ++t) {
        BlkBit& b = c.bit[t];
        const bool dep = t >= BLK_CB && ((b.P != nullptr && (b.mP & (uint32_t)(BLK_ROW - 1)) != 0u) || (b.Q != nullptr && b.shQ < (uint32_t)BLK_CB));
        b.cidx = -1;
        if (dep) { if (c.nC < BLK_MAXC) { b.cidx = c.nC; c.cbit[c.nC] = (uint8_t)t; } ++c.nC; }
    }
    c.d1row = (c.m1 & c.seqm[BLK_Q - 1]) != 0u;
    c.d2mode = !c.d2 ? 0 : (c.sh2 >= (uint32_t)BLK_CB ? 1 : 2);
    bool seqc = false;
    for (int i = 0; i < BLK_SB; ++i) seqc = seqc || c.bit[c.seq[i]].cidx >= 0;
    c.simple = (!seqc && !c.d1row && c.d2mode == 1) ? 1 : 0;
}

// number of column-profile vectors a space with these descriptors needs (host side planning uses the same rule)
// entry c of the profile of bit t: rate_t(c) / rate_t(0) for the 256 column patterns c
MMH_HD void blk_ctab_entry(const BlkCtx& c, int t, int col, double* ctab)
{
    const BlkBit& b = c.bit[t];
    ctab[b.cidx * BLK_ROW + col] = blk_rate1(b, (uint32_t)col) / blk_rate1(b, 0u);
}

// The source rows the OUTER phase of a block consumes, in order: rows in processing order, for each row the outer bits
// of `omask` (a mask over c.out[] indices) in ascending order.  chunk index -> element offset of the 256-double row.
struct BlkPlan {
    uint32_t base;                              // outer bits of the block (no lane part)
    int nE;                                     // outer edges per row
    uint8_t epos[BLK_MAXBITS];                  // their bit positions
    int8_t ecidx[BLK_MAXBITS];                  // their column-profile vectors (-1 = none)
};
MMH_HD void blk_plan(const BlkCtx& c, uint32_t o, bool adj, BlkPlan& p)
{
    p.base = blk_outer_mask(c, o);
    uint32_t m = adj ? (~o & ((1u << c.KO) - 1u)) : o;
    p.nE = 0;
    while (m) {
        const int t = c.out[blk_ffs(m)];
        p.epos[p.nE] = (uint8_t)t;
        p.ecidx[p.nE] = (int8_t)c.bit[t].cidx;
        ++p.nE;
        m &= m - 1;
    }
}
// element offset of the source row of chunk (row_it, k): rows in processing order, edges in ascending bit order
template <bool ADJ>
MMH_HD uint32_t blk_chunk_row(const BlkCtx& c, const BlkPlan& p, uint32_t row_it, int k)
{
    const uint32_t q = ADJ ? (uint32_t)(BLK_Q - 1) - row_it : row_it;
    return (p.base | c.seqm[q]) ^ (1u << p.epos[k]);
}

// Entry `idx` of the block's table of per-row scalars (idx = q * (13 + nE) + k): the rate of (row q, edge k) at the row's
// base state with the column bits cleared -- the column profile supplies the rest -- or the d2 part of the diagonal.
// Split in two so that the device can issue the loads of the NEXT block's table one iteration before it stores them.
struct BlkScReq { const double* a; const double* b; int dst; };       // value = (a ? *a : 1) * (b ? *b : 1), dst < 0: nothing to do
template <bool ADJ>
MMH_HD BlkScReq blk_sc_request(const BlkCtx& c, const BlkPlan& p, int idx)
{
    const int per = BLK_SC_OUT + p.nE;
    BlkScReq r{nullptr, nullptr, -1};
    if (idx >= BLK_Q * per) return r;
    const int q = idx / per, k = idx - q * per;
    const uint32_t s = p.base | c.seqm[q];
    r.dst = q * BLK_SCW + k;
    int t;
    uint32_t u = s;
    if (k < BLK_CB) t = k;                                             // column bit: scalar part of its rate in this row
    else if (k < BLK_SC_D2) {
        const int i = k - BLK_SC_SEQ;
        const bool set = (q >> i) & 1;
        t = c.seq[i];
        if (ADJ ? set : !set) { r.dst = -2 - r.dst; return r; }        // no such edge in this row: the entry is 0
        if (!ADJ) u = s ^ (1u << t);
    } else if (k == BLK_SC_D2) {
        if (c.d2mode == 1) r.a = c.d2 + ((s >> c.sh2) & c.m2);
        else r.dst = -2 - r.dst;
        return r;
    } else {
        t = p.epos[k - BLK_SC_OUT];
        if (!ADJ) u = s ^ (1u << t);
    }
    const BlkBit& b = c.bit[t];
    if (b.P) r.a = b.P + (u & b.mP);
    if (b.Q) r.b = b.Q + ((u >> b.shQ) & b.mQ);
    return r;
}
MMH_HD void blk_sc_store(const BlkScReq& r, double va, double vb, double* sc)
{
    if (r.dst >= 0) sc[r.dst] = va * vb;
    else if (r.dst < -1) sc[-2 - r.dst] = 0.0;
}
template <bool ADJ>
MMH_HD void blk_sc_entry(const BlkCtx& c, const BlkPlan& p, int idx, double* sc)
{
    const BlkScReq r = blk_sc_request<ADJ>(c, p, idx);
    blk_sc_store(r, r.a ? *r.a : 1.0, r.b ? *r.b : 1.0, sc);
}

}  // namespace mmh
