// Blocked substitution solve of the big lattices (K >= 13 bits): warp-private shared-memory blocks, skewed wavefront.
//
// Replaces the (k+1)-sweep Jacobi iteration of the reference (metmhn/jx/likelihood.py:231-262 `R_i_inv_vec`,
// metmhn/jx/vanilla.py:269-305 `R_inv_vec`) and the per-event Kronecker shuffles behind it (kronvec.py:259-539) by ONE
// exact forward (or adjoint) substitution on the subset lattice, blocked for the memory system of a B200:
//
//   * the lattice index s has K bits; the low eight are "column" bits (bits 0-2 = eight states in the registers of a
//     lane, bits 3-7 = the 32 lanes of a warp), four more bits are "sequence" bits, the K-12 others are "outer" bits;
//   * a BLOCK is the 2^12-state sub-lattice spanned by the column and sequence bits for one value of the outer bits.
//     One WARP owns one block at a time and keeps its 4096 solved values in its private 32 KB of shared memory;
//   * every edge on an OUTER bit reads the value of the same position in a block finished by an earlier launch
//     (launches go level by level over the popcount of the outer bits only: K-11 launches instead of K-3), i.e.
//     (K-12)/2 global reads per state instead of (K-4)/2 -- the reads that bound the previous tile kernel
//     (profiles/r1_final_solve_tile_ncu_full.txt: 76 B of L2 traffic per state against 8 algorithmic);
//   * inside the block nothing leaves the SM: sequence-bit edges read rows the lane itself finished earlier,
//     lane-bit edges read the row of a neighbouring lane, register-bit edges are register arithmetic;
//   * the 16 rows of a block are processed as a SKEWED WAVEFRONT: at step t lane l works on row t - popcount(l)
//     (adjoint: mirrored).  A lane-bit neighbour l^a has one bit less, so it finished the same row one step earlier:
//     every lane is busy at every step (except 5 fill / drain steps per block), there is no __syncthreads and no
//     shuffle, only a __syncwarp per step.
//
// The kernel body is written against a small abstract description of the space (BlkCtx: per-bit rate descriptors,
// diagonal, right-hand side functor) and is plain C++: tests/host/blk_host_test.cpp compiles the very same functions
// for the host and emulates a warp lane by lane against a sequential substitution.
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
constexpr int BLK_CB = 8;                       // column bits (3 register bits + 5 lane bits)
constexpr int BLK_SB = 4;                       // sequence bits
constexpr int BLK_Q = 1 << BLK_SB;              // rows of a block
constexpr int BLK_ROW = 1 << BLK_CB;            // doubles per row
constexpr int BLK_DOUBLES = BLK_Q * BLK_ROW;    // 4096 doubles = 32 KB per warp
constexpr int BLK_STEPS = BLK_Q + 5;            // skew: 5 fill / drain steps

// Edge u -> u | (1 << t) of bit t (t not in u):  rate_t(u) = P[u & mP] * Q[(u >> shQ) & mQ]   (null pointer = 1).
// P is read eight consecutive entries at a time (the eight register states of a lane), so mP >= 7; Q must not depend
// on the three register bits (shQ >= 3).
struct BlkBit {
    const double* P;
    const double* Q;
    uint32_t mP, mQ, shQ, pad;
};

struct BlkCtx {
    BlkBit bit[BLK_MAXBITS];                    // by bit position
    const double* d1;                           // diag(u) = d1[u & m1] + (d2 ? d2[(u >> sh2) & m2] : 0)
    const double* d2;
    uint32_t m1, m2, sh2;
    int K, KO;
    uint32_t seqdep;                            // column bits whose rate depends on the sequence bits (mask over bits 0..7)
    uint8_t seq[BLK_SB];                        // positions of the sequence bits, ascending
    uint8_t out[BLK_MAXBITS];                   // positions of the outer bits, ascending
    double cs[BLK_Q][BLK_CB + 1];               // rate_t(u | seq(q)) / rate_t(u) for column bit t (+1: padding against bank conflicts)
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

MMH_HD void blk_ld8(const double* __restrict__ p, double (&v)[8])      // 64 bytes, 32-byte aligned
{
#if defined(__CUDA_ARCH__)
    asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7]) : "l"(p + 4));
#else
    for (int j = 0; j < 8; ++j) v[j] = p[j];
#endif
}
MMH_HD void blk_st8(double* __restrict__ p, const double (&v)[8])
{
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" :: "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "l"(p) : "memory");
    asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" :: "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]), "l"(p + 4) : "memory");
#else
    for (int j = 0; j < 8; ++j) p[j] = v[j];
#endif
}

// Shared-memory layout of a block: row q, then four 16-byte pieces, each piece lane-contiguous:
//   double index = q * 256 + piece * 64 + lane * 2 + e,  register state j = 2 * piece + e
// so that a 16-byte access of a warp covers 512 contiguous bytes whatever lane permutation (l ^ a) it uses.
MMH_HD void blk_lds8(const double* __restrict__ sm, int q, int lane, double (&v)[8])
{
    const double* p = sm + q * BLK_ROW + lane * 2;
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) { v[2 * pc] = p[pc * 64]; v[2 * pc + 1] = p[pc * 64 + 1]; }
}
MMH_HD void blk_sts8(double* __restrict__ sm, int q, int lane, const double (&v)[8])
{
    double* p = sm + q * BLK_ROW + lane * 2;
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) { p[pc * 64] = v[2 * pc]; p[pc * 64 + 1] = v[2 * pc + 1]; }
}

// global offset of row q of a block (its sequence bits) and of the outer coordinate o
MMH_HD uint32_t blk_seq_mask(const BlkCtx& c, uint32_t q)
{
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < BLK_SB; ++i) m |= ((q >> i) & 1u) << c.seq[i];
    return m;
}
MMH_HD uint32_t blk_outer_mask(const BlkCtx& c, uint32_t o)
{
    uint32_t m = 0;
    for (int i = 0; i < c.KO; ++i) m |= ((o >> i) & 1u) << c.out[i];
    return m;
}

// eight rates of the edges of bit `b` that START at the states u0 .. u0+7 (u0 a multiple of 8, bit b clear in u0)
MMH_HD void blk_rate8(const BlkBit& b, uint32_t u0, double (&r)[8])
{
    const double q = b.Q ? b.Q[(u0 >> b.shQ) & b.mQ] : 1.0;
    if (b.P) {
        blk_ld8(b.P + (u0 & b.mP), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] *= q;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = q;
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

// ---- per-lane state of one block ---------------------------------------------------------------------------
// Rates of the column-bit edges of the lane's eight states at row 0 of the block (the sequence bits only scale them:
// BlkCtx::cs).  Forward: edges ENDING in the lane's states; adjoint: edges STARTING there.
struct BlkLane {
    uint32_t base;                              // outer bits of the block | lane << 3
    double Rr[3][4];                            // register bits: edge k of bit b, in the order of the source states lacking b
    double RL[5][8];                            // lane bits 3..7 (0 where the lane has no such edge)
};

template <bool ADJ>
MMH_HD void blk_lane_setup(const BlkCtx& c, uint32_t outer_mask, int lane, BlkLane& L)
{
    // forward rows start at q = 0 (no sequence bit), adjoint rows at q = Q-1; the rates are taken at q = 0 in both
    // cases and scaled by cs[q][t] per row when the space needs it
    const uint32_t u0 = outer_mask | ((uint32_t)lane << 3);
    L.base = u0;
    {
        double r[8];
        blk_rate8(c.bit[0], u0, r);
        L.Rr[0][0] = r[0]; L.Rr[0][1] = r[2]; L.Rr[0][2] = r[4]; L.Rr[0][3] = r[6];      // 0->1 2->3 4->5 6->7
        blk_rate8(c.bit[1], u0, r);
        L.Rr[1][0] = r[0]; L.Rr[1][1] = r[1]; L.Rr[1][2] = r[4]; L.Rr[1][3] = r[5];      // 0->2 1->3 4->6 5->7
        blk_rate8(c.bit[2], u0, r);
        L.Rr[2][0] = r[0]; L.Rr[2][1] = r[1]; L.Rr[2][2] = r[2]; L.Rr[2][3] = r[3];      // 0->4 1->5 2->6 3->7
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const uint32_t bit = 8u << a;
        const bool has = (u0 & bit) != 0u;
        const bool edge = ADJ ? !has : has;      // forward: the edge comes from l ^ a (which lacks the bit); adjoint: it goes there
        if (edge) blk_rate8(c.bit[3 + a], u0 & ~bit, L.RL[a]);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) L.RL[a][j] = 0.0;
        }
    }
}

// One step of one lane.  `sm_rd` / `sm_wr` are the block's shared memory (the same pointer on the device; the host
// emulation reads from a snapshot taken at the start of the step).  `omask` = outer coordinate bits of the block whose
// edges this pass follows (forward: the set outer bits, adjoint: the clear ones), as a mask over c.out[] indices.
template <bool ADJ, class RHS>
MMH_HD void blk_lane_step(const BlkCtx& c, const BlkLane& L, int lane, int t, uint32_t omask, double* __restrict__ v,
                          const double* sm_rd, double* sm_wr, const RHS& rhs)
{
    const int pl = blk_popc((uint32_t)lane);
    const int qi = t - (ADJ ? 5 - pl : pl);
    if (qi < 0 || qi >= BLK_Q) return;
    const int q = ADJ ? BLK_Q - 1 - qi : qi;
    const uint32_t s0 = L.base | blk_seq_mask(c, (uint32_t)q);
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    rhs(s0, acc);
    // ---- outer bits: values of blocks finished by earlier launches (global memory) ----
    {
        uint32_t m = omask;
        while (m) {
            const int i0 = blk_ffs(m);
            m &= m - 1;
            const int i1 = m ? blk_ffs(m) : -1;
            if (m) m &= m - 1;
            const uint32_t b0 = 1u << c.out[i0];
            const uint32_t b1 = i1 >= 0 ? 1u << c.out[i1] : 0u;
            double y0[8], y1[8], r0[8], r1[8];
            blk_ld8(v + (s0 ^ b0), y0);
            if (i1 >= 0) blk_ld8(v + (s0 ^ b1), y1);
            blk_rate8(c.bit[c.out[i0]], ADJ ? s0 : (s0 ^ b0), r0);
            if (i1 >= 0) blk_rate8(c.bit[c.out[i1]], ADJ ? s0 : (s0 ^ b1), r1);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fma(r0[j], y0[j], acc[j]);
            if (i1 >= 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fma(r1[j], y1[j], acc[j]);
            }
        }
    }
    // ---- sequence bits: rows of this block the lane finished at earlier steps ----
#pragma unroll
    for (int i = 0; i < BLK_SB; ++i) {
        const bool set = (q >> i) & 1;
        if (ADJ ? !set : set) {
            const uint32_t bit = 1u << c.seq[i];
            double y[8], r[8];
            blk_lds8(sm_rd, q ^ (1 << i), lane, y);
            blk_rate8(c.bit[c.seq[i]], ADJ ? s0 : (s0 ^ bit), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fma(r[j], y[j], acc[j]);
        }
    }
    // ---- lane bits: the same row of the neighbouring lanes (finished one step earlier) ----
    const bool dep = c.seqdep != 0u;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        double y[8];
        const bool has = (lane >> a) & 1;
        if (ADJ ? !has : has) {
            blk_lds8(sm_rd, q, lane ^ (1 << a), y);
            if (dep) {
                const double k = c.cs[q][3 + a];
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] *= k;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fma(L.RL[a][j], y[j], acc[j]);
        }
    }
    // ---- diagonal ----
    double inv[8];
    {
        double d[8];
        blk_ld8(c.d1 + (s0 & c.m1), d);
        if (c.d2) {
            const double k = c.d2[(s0 >> c.sh2) & c.m2];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] += k;
        }
        blk_inv8(d, inv);
    }
    // ---- register bits ----
    double k0 = 1.0, k1 = 1.0, k2 = 1.0;
    if (dep) { k0 = c.cs[q][0]; k1 = c.cs[q][1]; k2 = c.cs[q][2]; }
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
    blk_sts8(sm_wr, q, lane, y);
    blk_st8(v + s0, y);
}

// rates of the column bits relative to row 0 of a block (see BlkCtx::cs); call after bit[], seq[] are set
MMH_HD void blk_ctx_cs_row(BlkCtx& c, int q)
{
    const uint32_t sm = blk_seq_mask(c, (uint32_t)q);
    for (int t = 0; t < BLK_CB; ++t) {
        const BlkBit& b = c.bit[t];
        double num = 1.0, den = 1.0;
        if (b.P) { num *= b.P[sm & b.mP]; den *= b.P[0]; }
        if (b.Q) { num *= b.Q[(sm >> b.shQ) & b.mQ]; den *= b.Q[0]; }
        c.cs[q][t] = num / den;
    }
    c.cs[q][BLK_CB] = 1.0;
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
}

// cs table + seqdep mask; call once bit[], seq[] are final (any thread / the host)
MMH_HD void blk_ctx_finish(BlkCtx& c)
{
    c.seqdep = 0;
    for (int q = 0; q < BLK_Q; ++q) {
        blk_ctx_cs_row(c, q);
        for (int t = 0; t < BLK_CB; ++t) if (c.cs[q][t] != 1.0) c.seqdep |= 1u << t;
    }
}

}  // namespace mmh
