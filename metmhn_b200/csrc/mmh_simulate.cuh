// GPU Gillespie sampler of the metMHN process (the step BEFORE the likelihood path: synthetic cohorts for recall studies and
// the frequency known-answer test).  Same process as the reference's `metmhn/simulations.py:8-77` (`single_traject`):
// primary tumour (PT) and metastasis (MT) evolve in lock-step until the seeding event -- every PT-side event, the seeding
// itself and a PT diagnosis are copied to the MT side -- and independently afterwards; a trajectory stops when the PT is
// diagnosed before seeding or when both tumours are diagnosed.  One thread per trajectory, genotypes as bit masks in
// registers, rates as products of exp(theta) entries, Philox4x32-10 counter-based random numbers keyed by
// (seed, trajectory): the result does not depend on the launch configuration.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mmh {

struct SimPar {                      // n_tot <= 29
    double W[32][32];                // exp(theta_ij), i != j   (row i = affected event, column j = present event)
    double Wpt[32][32];              // the same with the seeding column of the genomic events set to 1 (simulations.py:62-64)
    double base[32];                 // exp(theta_ii)
    double dp[32], dm[32];           // exp(log_d_p), exp(log_d_m)
};

__global__ void k_sim_prep(const double* __restrict__ params, int n_tot, SimPar* __restrict__ P)
{
    const int n = n_tot - 1;
    const double* th = params;
    const double* ldp = params + n_tot * n_tot;
    const double* ldm = ldp + n_tot;
    for (int t = threadIdx.x; t < 32 * 32; t += blockDim.x) {
        const int i = t >> 5, j = t & 31;
        double w = 1.0, wpt = 1.0;
        if (i < n_tot && j < n_tot && i != j) {
            w = exp(th[i * n_tot + j]);
            wpt = (j == n && i < n) ? 1.0 : w;
        }
        P->W[i][j] = w; P->Wpt[i][j] = wpt;
    }
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
        P->base[i] = i < n_tot ? exp(th[i * n_tot + i]) : 0.0;
        P->dp[i] = i < n_tot ? exp(ldp[i]) : 1.0;
        P->dm[i] = i < n_tot ? exp(ldm[i]) : 1.0;
    }
}

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ void philox4x32(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// geno: n_sim rows of 2n+1 int8 [PT_0, MT_0, ..., PT_{n-1}, MT_{n-1}, seeding]; order: 1 = PT diagnosed first, 2 = MT first,
// 0 = never seeded
__global__ void __launch_bounds__(128)
k_simulate(const SimPar* __restrict__ P, int n_tot, int64_t n_sim, uint64_t seed, int8_t* __restrict__ geno, int8_t* __restrict__ order)
{
    __shared__ SimPar sp;
    for (int t = threadIdx.x; t < (int)(sizeof(SimPar) / sizeof(double)); t += blockDim.x)
        reinterpret_cast<double*>(&sp)[t] = reinterpret_cast<const double*>(P)[t];
    __syncthreads();
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_sim) return;
    const int n = n_tot - 1;
    const uint32_t sbit = 1u << n;
    uint32_t pt = 0, mt = 0;                       // genotypes incl. the seeding bit (bit n)
    bool obs_pt = false, obs_mt = false;
    int ord = 0;
    for (uint32_t step = 0; step < 4u * 32u; ++step) {          // at most 2 n_tot + 2 events happen
        const bool seeded = (pt & sbit) != 0u;
        if ((obs_pt && obs_mt) || (obs_pt && !seeded)) break;
        // ---- rates: PT events 0..n (n = seeding), PT diagnosis, MT events, MT diagnosis ----
        double r_pt[32], r_mt[32];
        double r_po = 0.0, r_mo = 0.0, tot = 0.0;
        const bool pt_live = !obs_pt, mt_live = seeded && !obs_mt;
#pragma unroll 1
        for (int i = 0; i < n_tot; ++i) {
            double a = 0.0, b = 0.0;
            if (pt_live && !((pt >> i) & 1u)) {
                a = sp.base[i];
                for (uint32_t m = pt; m; m &= m - 1) a *= sp.Wpt[i][__ffs(m) - 1];
            }
            if (mt_live && !((mt >> i) & 1u)) {
                b = sp.base[i];
                for (uint32_t m = mt; m; m &= m - 1) b *= sp.W[i][__ffs(m) - 1];
            }
            r_pt[i] = a; r_mt[i] = b;
            tot += a + b;
        }
        if (pt_live) { r_po = 1.0; for (uint32_t m = pt; m; m &= m - 1) r_po *= sp.dp[__ffs(m) - 1]; }
        if (mt_live) { r_mo = 1.0; for (uint32_t m = mt; m; m &= m - 1) r_mo *= sp.dm[__ffs(m) - 1]; }
        tot += r_po + r_mo;
        // ---- one uniform number per step ----
        uint32_t c[4] = {step, 0u, (uint32_t)id, (uint32_t)((uint64_t)id >> 32)};
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u = ((double)(((uint64_t)c[0] << 21) ^ (uint64_t)(c[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0) * tot;
        // ---- inverse CDF in the order PT events, PT diagnosis, MT events, MT diagnosis ----
        int ev = -1;
        double cum = 0.0;
#pragma unroll 1
        for (int i = 0; i < n_tot && ev < 0; ++i) { cum += r_pt[i]; if (u < cum) ev = i; }
        if (ev < 0) { cum += r_po; if (u < cum) ev = n_tot; }
#pragma unroll 1
        for (int i = 0; i < n_tot && ev < 0; ++i) { cum += r_mt[i]; if (u < cum) ev = n_tot + 1 + i; }
        if (ev < 0) ev = r_mo > 0.0 ? 2 * n_tot + 1 : -2;
        if (ev == -2) {                              // rounding at the upper end: take the last event with a positive rate
            for (int i = n_tot - 1; i >= 0 && ev < 0; --i) if (r_mt[i] > 0.0) ev = n_tot + 1 + i;
            if (ev < 0 && r_po > 0.0) ev = n_tot;
            for (int i = n_tot - 1; i >= 0 && ev < 0; --i) if (r_pt[i] > 0.0) ev = i;
            if (ev < 0) break;
        }
        // ---- apply (simulations.py:52-55: before seeding a PT-side event hits both copies) ----
        if (ev < n_tot) {
            pt |= 1u << ev;
            if (!seeded) mt |= 1u << ev;
        } else if (ev == n_tot) {
            if (seeded && !obs_mt) ord = 1;
            obs_pt = true;
            if (!seeded) obs_mt = true;
        } else if (ev <= 2 * n_tot) {
            mt |= 1u << (ev - n_tot - 1);
        } else {
            if (!obs_pt) ord = 2;
            obs_mt = true;
        }
    }
    int8_t* g = geno + id * (2 * n + 1);
    for (int e = 0; e < n; ++e) { g[2 * e] = (int8_t)((pt >> e) & 1u); g[2 * e + 1] = (int8_t)((mt >> e) & 1u); }
    g[2 * n] = (int8_t)((pt >> n) & 1u);
    order[id] = (int8_t)ord;
}

}  // namespace mmh
