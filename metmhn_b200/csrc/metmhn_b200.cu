// Host side of libmetmhn_b200.so: dataset preprocessing (done once, the dataset is constant over the
// L-BFGS iterations, regularized_optimization.py:328), scratch/chunk planning, the per-evaluation launch
// sequence and the C-ABI of include/metmhn_b200.h.  No CPU fallback: every compute path is a CUDA kernel.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "../../include/metmhn_b200.h"
#include "mmh_device.cuh"
#include "mmh_rowblock.cuh"
#include "mmh_simulate.cuh"
#include "mmh_lbfgs.hpp"

using namespace mmh;

static constexpr int RED_SLICES = 32;            // first stage of the gradient-partials reduction
static constexpr size_t FIN_SMEM = (size_t)(NACC * NR * NR + FIN_WARPS * NR * 33) * sizeof(double);

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? MMH_ENOMEM : MMH_ECUDA,             \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

namespace {

struct Range { uint64_t off = 0; uint32_t cnt = 0; };

struct ChunkPlan {
    uint64_t space0 = 0;
    uint32_t nspaces = 0;
    Range pre4, main_small4, sec_small4;         // small-tier spaces with K >= 7 (four states per lane)
    Range rb_list;                               // pairs of the row-block kernel
    Range setup, setup_wide, diag, pre, main_small, sec_small, logp, joints, st_a, st_ar, st_b, st_bn, pf_lo, fin;
    bool wide = false;                           // some group has more than MAXT bits
    std::vector<Range> main_lv, sec_lv;          // big-tier segments per popcount level (generic kernel)
    std::vector<Range> main_lvt, sec_lvt;        // big-tier tiles per level (tiled kernel)
    std::vector<Range> main_lvt_adj, main_lvt_adjb;  // adjoint pass of the main tiled spaces: plain / with fused B statistics
    std::vector<Range> main_rb;                  // row-block kernel: blocks per OUTER level (both passes)
    std::vector<Range> main_lvw;                 // pairs with a wide PT group: tiles per level (both passes, k_solve_tile_w)
    uint64_t scratch = 0;                        // doubles
};

struct PatientPlan {
    std::vector<SpaceDev> sp;                    // links are local indices
    uint64_t scratch = 0;
};

}  // namespace

struct mmh_handle {
    int device = 0, n = 0, n_tot = 0;
    int64_t n_dat = 0, n_em = 0;
    std::vector<ChunkPlan> chunks;
    SpaceDev* d_spaces = nullptr;
    uint32_t* d_lists = nullptr;
    Item* d_items = nullptr;
    uint32_t* d_hs = nullptr;
    uint32_t* d_hsidx = nullptr;                 // [bits][level] -> first entry of that popcount level in d_hs
    uint16_t* d_rblv = nullptr;                  // row-block kernel: level lists in a bank-conflict-free order
    int nsm = 0;
    uint8_t* d_cls = nullptr;
    double* d_cnt = nullptr;
    EvalPar* d_par = nullptr;
    double *d_params = nullptr, *d_scratch = nullptr, *d_logp = nullptr, *d_partial = nullptr;
    double *d_partial2 = nullptr, *d_diracc = nullptr, *d_tdir = nullptr, *d_out = nullptr, *h_out = nullptr;
    // Chunks are independent, so they are spread round-robin over NS side streams, each with its own scratch
    // buffer and gradient partials: the thin popcount levels and the tails of one chunk overlap with the work
    // of the others.  Chunk -> stream is static and k_final adds the slots in a fixed order: results stay
    // bit-identical from call to call.
    static constexpr int NS = 32;                // most side streams; `ns` of them are used (MMH_STREAMS, default 6)
    int ns = 6;
    cudaStream_t stream = nullptr;               // main stream: parameters, k_prep, k_final, copies
    cudaStream_t side[NS] = {};
    cudaEvent_t ev_side[NS] = {}, ev_prep = nullptr;
    double* d_scratch_s[NS] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int fin_ctas = 0;
    int profile = 0;
    int use_graph = 1;                           // MMH_GRAPH=0 disables CUDA-graph replay of the launch sequence
    struct GraphEntry { const double* params; double w0, w1; int want_grad; cudaGraphExec_t exec; int64_t launches; };
    std::vector<GraphEntry> graphs;
    std::vector<cudaEvent_t> evpool;
    uint32_t max_joints = 0;
    mmh_stats_t st{};
    std::vector<int8_t> rows;                    // copy of the data matrix (n_dat x (2n+3)), for the per-row test hook
    void* comm = nullptr;                        // ncclComm_t once mmh_comm_init ran: every evaluation ends with an all-reduce
    bool comm_owned = true;
};

// MMH_ROWBLK=1 sends the pairs to the shared-memory row-block kernel instead of the tile kernels (measured slower, kept
// for experiments and covered by the parity tests); read once per process, so the plan and the launches always agree
static bool use_rb()
{
    static const bool on = [] { const char* e = std::getenv("MMH_ROWBLK"); return e && std::atoi(e) != 0; }();
    return on;
}
// pairs solved by the row-block kernel (k_solve_rb, mmh_rowblock.cuh): plain tables, at least 7 PT bits
static bool rb_space(const SpaceDev& s)
{
    return use_rb() && s.kind == K_JOINT && !s.splitA && !s.splitB && s.KA >= RB_MINKA && (int)s.KA + (int)s.KB >= BIGK;
}
// pairs whose adjoint solve also produces the group-B statistics (k_solve_rb / k_solve_tile_adjb); needs splitA/splitB
static bool fused_b(const SpaceDev& s)
{
    return s.kind == K_JOINT && !s.splitA && !s.splitB && s.KA >= 4 && (int)s.KA + (int)s.KB >= BIGK;
}
// number of partial tables: per column-block level lA one per chunk of 32 column blocks; base[lA] = first slot
static uint32_t adjb_slots(int kbA, uint32_t* base)
{
    uint32_t n = 0;
    double c = 1.0;                                     // C(kbA, lA)
    for (int lA = 0; lA <= kbA; ++lA) {
        if (base) base[lA] = n;
        const uint64_t nA = (uint64_t)(c + 0.5);
        n += (uint32_t)std::max<uint64_t>(1, (nA + 31) / 32);
        c = c * (kbA - lA) / (lA + 1);
    }
    return n;
}

// pairs whose group-B statistics are taken one lane per row (k_stats_b_narrow)
static bool narrow_b(const SpaceDev& s) { return s.kind == K_JOINT && !fused_b(s) && s.KA <= STB_NARROW; }

static uint64_t space_scratch(SpaceDev& s, uint64_t off)
{
    const uint64_t NA = 1ull << s.KA, NB = 1ull << s.KB, N = NA * NB;
    auto take = [&](uint64_t len) { uint64_t o = off; off += (len + 3) & ~3ull; return o; };
    // A joint side's table (NR x 2^KG) is small next to the 2^(KA+KB) vectors; a single-group space's full table
    // would be NR times its vectors, so those switch to the product form early.
    const int max_full = (s.kind == K_JOINT) ? MAXT : 8;
    const bool has_diag_tables = (s.kind == K_PRE || s.kind == K_JOINT || s.kind == K_S2);
    auto table = [&](int KG, uint8_t& split) {
        if (KG <= max_full) { split = 0; return take((uint64_t)NR << KG); }
        split = (uint8_t)((KG + 1) / 2);             // rate(i,u) = T1[i][u_lo] * T2[i][u_hi] + full special-row vectors
        return take(((uint64_t)NR << split) + ((uint64_t)NR << (KG - split)) + ((has_diag_tables ? 3ull : 1ull) << KG));
    };
    s.splitA = s.splitB = 0;
    s.tabA = table(s.KA, s.splitA);
    s.tabB = s.kind == K_JOINT ? table(s.KB, s.splitB) : 0;
    s.y_off = take(N);
    s.x_off = take(N);
    s.stA = s.stB = s.stP = s.stPB = s.tabR = 0;
    s.slices = 1; s.slicesB = 1;
    if (s.kind == K_JOINT) {
        // group-A statistics sum over uB, group-B statistics over uA: slice the summed range so that lopsided pairs
        // (one tumour with many events, the other with few) still expose enough parallel work
        // (the partial tables are capped at 2M doubles: more slices only where the table is small)
        const uint64_t capA = std::max<uint64_t>(8, (2ull << 20) / ((s.KA + 1) * NA));
        const uint64_t capB = std::max<uint64_t>(8, (2ull << 20) / ((s.KB + 1) * NB));
        s.slices = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(256, capA), std::max<uint64_t>(1, NB / 128));
        s.slicesB = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(256, capB), std::max<uint64_t>(1, NA / 2048));
        if (fused_b(s)) s.slicesB = adjb_slots(s.KA - 4, nullptr);     // one partial table per (lA, column chunk)
        if (rb_space(s)) s.slicesB = 1u << std::max(0, (int)s.KA - RB_KI);   // one per value of the outer column bits
        if (narrow_b(s)) s.slicesB = 1;
        s.stA = take((s.KA + 1) * NA);
        s.stB = take((s.KB + 1) * NB);
        s.stP = take((uint64_t)s.slices * (s.KA + 1) * NA);
        s.stPB = take((uint64_t)s.slicesB * (s.KB + 1) * NB);
        if (rb_space(s)) s.tabR = take(RB_TABR);
    } else if (s.kind != K_PRE && s.splitA) {
        // product-form gradient: weighted marginals over the low / high part (k_pfin_lo / k_pfin_hi)
        const uint64_t N1 = 1ull << s.splitA, N2 = 1ull << (s.KA - s.splitA);
        s.slices = (uint32_t)std::min<uint64_t>(16, std::max<uint64_t>(1, N2 / 8));
        s.stP = take((uint64_t)s.slices * (NR + s.KA) * N1 + std::max<uint64_t>(1, N1 >> 7) * (NR + s.KA) * N2);   // k_pf partials
    }
    return off;
}

static int build_patient(const int8_t* row, int n, int32_t pid, PatientPlan& pp, mmh_stats_t& st,
                         std::vector<double>& cnt_dm2, uint8_t& cls)
{
    const int n_tot = n + 1;
    const int typ = row[2 * n + 2];
    pp.sp.clear();
    cls = 255;
    if (typ < 0 || typ > 3) return MMH_OK;       // rows of an unknown type contribute nothing (regularized_optimization.py:187-254)
    for (int c = 0; c < 2 * n + 1; ++c)
        if (row[c] != 0 && row[c] != 1) return fail(MMH_EINVAL, "genotype entries must be 0 or 1");
    auto base = [&](Kind k) {
        SpaceDev s{};
        s.kind = k; s.n_tot = (uint8_t)n_tot; s.patient = pid; s.cls = (typ == 0) ? 0 : 1;
        s.joint = s.pre = s.pf = s.mf = -1;
        return s;
    };
    if (typ == 0 || typ == 1) {
        SpaceDev s = base(K_S1);
        int k = 0;
        for (int j = 0; j < n_tot; ++j)
            if (row[2 * j]) { if (k >= MMH_MAX_BITS) return fail(MMH_ETOOLARGE, "unpaired patient exceeds the supported lattice size"); s.evA[k++] = (uint8_t)j; }
        s.KA = (uint8_t)k;
        st.k_hist[typ][k]++;
        st.states_value_grad += std::ldexp(1.0, k);
        st.alg_flops += std::ldexp(1.0, k) * (2.0 * k + 4.0 * n_tot);
        pp.sp.push_back(s);
        cls = (uint8_t)s.cls;
    } else if (typ == 2) {
        SpaceDev s = base(K_S2);
        int k = 0;
        for (int j = 0; j < n; ++j)
            if (row[2 * j + 1]) { if (k >= MMH_MAX_BITS - 1) return fail(MMH_ETOOLARGE, "unpaired patient exceeds the supported lattice size"); s.evA[k++] = (uint8_t)j; cnt_dm2[j] += 1.0; }
        s.evA[k++] = (uint8_t)n;
        cnt_dm2[n] += 1.0;
        s.KA = (uint8_t)k;
        st.k_hist[2][k]++;
        st.states_value_grad += std::ldexp(1.0, k);
        st.alg_flops += std::ldexp(1.0, k) * (2.0 * k + 4.0 * n_tot);
        pp.sp.push_back(s);
        cls = 1;
    } else if (typ == 3) {
        if (row[2 * n] != 1) return fail(MMH_EINVAL, "paired row (type 3) without the seeding bit");
        int order = row[2 * n + 1];
        SpaceDev j = base(K_JOINT), pre = base(K_PRE);
        std::vector<int> both, ponly, monly;
        for (int e = 0; e < n; ++e) {
            if (row[2 * e] && row[2 * e + 1]) both.push_back(e);
            else if (row[2 * e]) ponly.push_back(e);
            else if (row[2 * e + 1]) monly.push_back(e);
        }
        const int nb = (int)both.size(), ka = nb + (int)ponly.size(), kb = nb + (int)monly.size();
        if (ka + kb > MMH_MAX_BITS)
            return fail(MMH_ETOOLARGE, "paired patient exceeds the supported lattice size");
        for (int b = 0; b < nb; ++b) { j.evA[b] = j.evB[b] = pre.evA[b] = (uint8_t)both[b]; }
        for (size_t b = 0; b < ponly.size(); ++b) j.evA[nb + b] = (uint8_t)ponly[b];
        for (size_t b = 0; b < monly.size(); ++b) j.evB[nb + b] = (uint8_t)monly[b];
        for (int e = 0; e < n; ++e) { if (row[2 * e]) j.ptmask |= 1u << e; if (row[2 * e + 1]) j.mtmask |= 1u << e; }
        j.KA = (uint8_t)ka; j.KB = (uint8_t)kb; j.nb = (uint8_t)nb;
        pre.KA = (uint8_t)nb; pre.nb = (uint8_t)nb;
        j.has_pf = (order == 0 || order == 1);
        j.has_mf = (order != 1);
        // local indices: 0 = pre, 1 = joint, then pf / mf
        pre.joint = 1; j.pre = 0;
        pp.sp.push_back(pre);
        pp.sp.push_back(j);
        const int kj = ka + kb + 1;
        st.k_hist[3][kj]++;
        double neff = std::ldexp(1.0, ka + kb);
        st.alg_flops += neff * (2.0 * (ka + kb) + 8.0 * n_tot);
        if (j.has_pf) {
            SpaceDev s = base(K_PF);
            s.KA = (uint8_t)kb; std::memcpy(s.evA, j.evB, MAXG); s.joint = 1;
            pp.sp[1].pf = (int32_t)pp.sp.size();
            pp.sp.push_back(s);
            neff += std::ldexp(1.0, kb);
            st.alg_flops += std::ldexp(1.0, kb) * (2.0 * kb + 4.0 * n_tot);
        }
        if (j.has_mf) {
            SpaceDev s = base(K_MF);
            s.KA = (uint8_t)ka; std::memcpy(s.evA, j.evA, MAXG); s.joint = 1;
            pp.sp[1].mf = (int32_t)pp.sp.size();
            pp.sp.push_back(s);
            neff += std::ldexp(1.0, ka);
            st.alg_flops += std::ldexp(1.0, ka) * (2.0 * ka + 4.0 * n_tot);
        }
        st.states_value_grad += neff;
        cls = 1;
    }
    uint64_t off = 0;
    for (auto& s : pp.sp) off = space_scratch(s, off);
    pp.scratch = off;
    return MMH_OK;
}

static int create_impl(mmh_handle* h, int n_mut, const int8_t* dat, int64_t n_dat, int64_t row_stride,
                       int device, int64_t chunk_bytes);

extern "C" int mmh_create(mmh_handle** out, int n_mut, const int8_t* dat, int64_t n_dat, int64_t row_stride,
                          int device, int64_t chunk_bytes)
{
    if (!out || !dat || n_mut < 1 || n_mut > MMH_MAX_MUT || n_dat < 0 || row_stride < 2 * n_mut + 3)
        return fail(MMH_EINVAL, "mmh_create: bad arguments (n_mut must be 1..28, row_stride >= 2n+3)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(MMH_ECUDA, "mmh_create: no usable CUDA device (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    mmh_handle* h = new mmh_handle();
    const int rc = create_impl(h, n_mut, dat, n_dat, row_stride, device, chunk_bytes);
    if (rc != MMH_OK) { const std::string keep = g_err; mmh_destroy(h); g_err = keep; return rc; }   // nothing leaks on a failed create
    *out = h;
    return MMH_OK;
}

static int create_impl(mmh_handle* h, int n_mut, const int8_t* dat, int64_t n_dat, int64_t row_stride,
                       int device, int64_t chunk_bytes)
{
    h->device = device; h->n = n_mut; h->n_tot = n_mut + 1; h->n_dat = n_dat;
    h->st.n_dat = n_dat;
    const int n = n_mut;

    // ---- parse rows into per-patient space plans ------------------------------------------------------
    std::vector<PatientPlan> pats((size_t)n_dat);
    std::vector<uint8_t> cls((size_t)std::max<int64_t>(n_dat, 1), 255);
    std::vector<double> cnt_dm2(NR, 0.0);
    for (int64_t p = 0; p < n_dat; ++p) {
        const int8_t* row = dat + p * row_stride;
        h->n_em += row[2 * n];
        int rc = build_patient(row, n, (int32_t)p, pats[(size_t)p], h->st, cnt_dm2, cls[(size_t)p]);
        if (rc != MMH_OK) return rc;
    }
    h->st.n_em = h->n_em;
    h->st.alg_bytes = 32.0 * h->st.states_value_grad;
    h->rows.resize((size_t)n_dat * (2 * n + 3));
    for (int64_t p = 0; p < n_dat; ++p) std::memcpy(h->rows.data() + (size_t)p * (2 * n + 3), dat + p * row_stride, (size_t)(2 * n + 3));

    // ---- pack patients into scratch chunks (largest first) --------------------------------------------
    std::vector<int64_t> order((size_t)n_dat);
    for (int64_t p = 0; p < n_dat; ++p) order[(size_t)p] = p;
    std::stable_sort(order.begin(), order.end(),
                     [&](int64_t a, int64_t b) { return pats[(size_t)a].scratch > pats[(size_t)b].scratch; });
    // default chunk budget: 1 GiB, but small datasets are still cut into about two chunks per side stream so that
    // independent chunks can overlap (a single chunk would serialise every thin popcount level)
    if (const char* e = std::getenv("MMH_STREAMS")) h->ns = std::max(1, std::min((int)mmh_handle::NS, std::atoi(e)));
    uint64_t total_scratch = 0;
    for (const auto& pp : pats) total_scratch += pp.scratch;
    uint64_t budget = (uint64_t)((int64_t)1 << 30) / 8;
    budget = std::min<uint64_t>(budget, std::max<uint64_t>(((uint64_t)8 << 20) / 8, total_scratch / (2 * (uint64_t)h->ns)));
    if (chunk_bytes > 0) budget = (uint64_t)chunk_bytes / 8;
    std::vector<SpaceDev> spaces;
    std::vector<uint32_t> lists;
    std::vector<Item> items;
    std::vector<std::vector<uint32_t>> hs_tab(MMH_MAX_BITS);     // popcount-sorted block indices per K-5
    std::vector<std::vector<uint32_t>> hs_lvl(MMH_MAX_BITS);
    std::vector<uint64_t> hs_off(MMH_MAX_BITS, 0);
    std::vector<uint32_t> hs_all;
    auto need_hs = [&](int kh) {
        if (!hs_tab[kh].empty()) return;
        const uint32_t n_blk = 1u << kh;
        std::vector<uint32_t> cntl(kh + 2, 0);
        for (uint32_t v = 0; v < n_blk; ++v) cntl[__builtin_popcount(v) + 1]++;
        for (int l = 0; l <= kh; ++l) cntl[l + 1] += cntl[l];
        hs_lvl[kh] = cntl;
        std::vector<uint32_t> pos(cntl.begin(), cntl.end() - 1);
        hs_tab[kh].resize(n_blk);
        for (uint32_t v = 0; v < n_blk; ++v) hs_tab[kh][pos[__builtin_popcount(v)]++] = v;
        hs_off[kh] = hs_all.size();
        hs_all.insert(hs_all.end(), hs_tab[kh].begin(), hs_tab[kh].end());
    };

    size_t cur = 0;
    uint64_t max_scratch = 0;
    while (cur < order.size()) {
        ChunkPlan ck;
        ck.space0 = spaces.size();
        uint64_t used = 0;
        while (cur < order.size()) {
            const PatientPlan& pp = pats[(size_t)order[cur]];
            if (pp.sp.empty()) { ++cur; continue; }
            if (used > 0 && used + pp.scratch > budget) break;
            const uint32_t base_idx = (uint32_t)(spaces.size() - ck.space0);
            for (SpaceDev s : pp.sp) {
                s.y_off += used; s.x_off += used; s.tabA += used;
                if (s.kind == K_JOINT) { s.tabB += used; s.stA += used; s.stB += used; s.stP += used; s.stPB += used; if (s.tabR) s.tabR += used; }
                else if (s.stP) s.stP += used;                  // product-form single-tumour spaces
                if (s.joint >= 0) s.joint += base_idx;
                if (s.pre >= 0) s.pre += base_idx;
                if (s.pf >= 0) s.pf += base_idx;
                if (s.mf >= 0) s.mf += base_idx;
                spaces.push_back(s);
            }
            used += pp.scratch;
            ++cur;
        }
        ck.nspaces = (uint32_t)(spaces.size() - ck.space0);
        ck.scratch = used;
        max_scratch = std::max(max_scratch, used);
        if (ck.nspaces == 0) continue;
        // ---- work lists of this chunk -------------------------------------------------------------------
        const SpaceDev* sp = spaces.data() + ck.space0;
        auto list_of = [&](auto pred) {
            Range r; r.off = lists.size();
            for (uint32_t i = 0; i < ck.nspaces; ++i) if (pred(sp[i])) lists.push_back(i);
            r.cnt = (uint32_t)(lists.size() - r.off);
            return r;
        };
        auto is_main = [](const SpaceDev& s) { return s.kind == K_JOINT || s.kind == K_S1 || s.kind == K_S2; };
        auto is_sec = [](const SpaceDev& s) { return s.kind == K_PF || s.kind == K_MF; };
        auto bits = [](const SpaceDev& s) { return (int)s.KA + (int)s.KB; };
        ck.pre = list_of([&](const SpaceDev& s) { return s.kind == K_PRE && bits(s) < 7; });
        ck.main_small = list_of([&](const SpaceDev& s) { return is_main(s) && bits(s) < 7; });
        ck.sec_small = list_of([&](const SpaceDev& s) { return is_sec(s) && bits(s) < 7; });
        ck.pre4 = list_of([&](const SpaceDev& s) { return s.kind == K_PRE && bits(s) >= 7; });
        ck.main_small4 = list_of([&](const SpaceDev& s) { return is_main(s) && bits(s) >= 7 && bits(s) < BIGK; });
        ck.sec_small4 = list_of([&](const SpaceDev& s) { return is_sec(s) && bits(s) >= 7 && bits(s) < BIGK; });
        ck.logp = list_of([&](const SpaceDev& s) { return is_main(s); });
        ck.joints = list_of([&](const SpaceDev& s) { return s.kind == K_JOINT; });
        ck.rb_list = list_of([&](const SpaceDev& s) { return rb_space(s); });
        h->max_joints = std::max(h->max_joints, ck.joints.cnt);
        ck.setup.off = items.size();
        auto setup_items = [&](uint32_t i, uint32_t g, int KG, int K1) {
            auto blocks = [&](uint32_t part, int bits) { for (uint32_t u = 0; u < (1u << bits); u += 256) items.push_back({i, g | (part << 1), u}); };
            if (!K1) blocks(0, KG); else { blocks(0, K1); blocks(1, KG - K1); }
            if (KG > MAXT) ck.wide = true;
        };
        for (uint32_t i = 0; i < ck.nspaces; ++i) {
            setup_items(i, 0, sp[i].KA, sp[i].splitA);
            if (sp[i].kind == K_JOINT) setup_items(i, 1, sp[i].KB, sp[i].splitB);
        }
        ck.setup.cnt = (uint32_t)(items.size() - ck.setup.off);
        auto is_prod = [](const SpaceDev& s) { return s.kind != K_JOINT && s.kind != K_PRE && s.splitA; };
        ck.setup_wide.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i) {
            if (is_prod(sp[i])) continue;
            if (sp[i].splitA) for (uint32_t u = 0; u < (1u << sp[i].KA); u += 1024) items.push_back({i, 0u, u});
            if (sp[i].splitB) for (uint32_t u = 0; u < (1u << sp[i].KB); u += 1024) items.push_back({i, 1u, u});
        }
        ck.setup_wide.cnt = (uint32_t)(items.size() - ck.setup_wide.off);
        ck.diag.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i)
            if (is_prod(sp[i])) {
                const uint32_t nlo = std::max<uint32_t>(1u, (1u << sp[i].splitA) >> 7);
                const uint32_t nhi = std::max<uint32_t>(1u, (1u << (sp[i].KA - sp[i].splitA)) / DG_HIB);
                for (uint32_t hb = 0; hb < nhi; ++hb)
                    for (uint32_t lb = 0; lb < nlo; ++lb) items.push_back({i, lb, hb});
            }
        ck.diag.cnt = (uint32_t)(items.size() - ck.diag.off);
        // spaces the tile solve kernels take (k_solve_tile / k_solve_tile_adjb / k_solve_tile_w)
        auto tiled = [&](const SpaceDev& s) {
            if (bits(s) < BIGK || s.kind == K_PRE || rb_space(s)) return false;
            static const bool wide_on = [] { const char* e = std::getenv("MMH_WIDE_TILE"); return !e || std::atoi(e) != 0; }();
            if (s.kind == K_JOINT) return !s.splitB && (s.splitA ? (wide_on && s.splitA >= 4) : s.KA >= 4);
            return s.splitA >= 4;
        };
        // pairs with a wide PT group: cut like a product space (k_solve_tile_w, TM_WIDE in mmh_device.cuh)
        auto wide_t = [&](const SpaceDev& s) { return s.kind == K_JOINT && s.splitA != 0 && tiled(s); };
        auto levels_of = [&](auto pred, std::vector<Range>& lv) {
            int maxkh = -1;
            for (uint32_t i = 0; i < ck.nspaces; ++i) if (pred(sp[i]) && bits(sp[i]) >= BIGK && !tiled(sp[i]) && !rb_space(sp[i])) maxkh = std::max(maxkh, bits(sp[i]) - 7);
            if (maxkh < 0) return;
            lv.resize(maxkh + 1);
            for (int l = 0; l <= maxkh; ++l) {
                lv[l].off = items.size();
                for (uint32_t i = 0; i < ck.nspaces; ++i) {
                    if (!pred(sp[i]) || bits(sp[i]) < BIGK || tiled(sp[i]) || rb_space(sp[i])) continue;
                    const int kh = bits(sp[i]) - 7;          // blocks of 128 states
                    if (l > kh) continue;
                    need_hs(kh);
                    const uint32_t a = hs_lvl[kh][l], b = hs_lvl[kh][l + 1];
                    for (uint32_t r = a; r < b; r += SEGB)
                        items.push_back({i, (uint32_t)(hs_off[kh] + r), std::min<uint32_t>(SEGB, b - r)});
                }
                lv[l].cnt = (uint32_t)(items.size() - lv[l].off);
            }
        };
        // tiled kernel: rows x column blocks, one launch per level lA + lB; a CTA item is up to TILES_PER_CTA warp
        // tiles (8 rows x 16 columns) of one (lA, lB) split
        auto levels_of_t = [&](auto pred, std::vector<Range>& lv) {
            int maxl = -1;
            auto dims = [&](const SpaceDev& s, int& kbA, int& kbB) {
                if (s.kind == K_JOINT && s.splitA) { kbA = s.splitA - 4; kbB = s.KB + s.KA - s.splitA; }
                else if (s.kind == K_JOINT) { kbA = s.KA - 4; kbB = s.KB; }
                else { kbA = s.splitA - 4; kbB = s.KA - s.splitA; }
            };
            for (uint32_t i = 0; i < ck.nspaces; ++i)
                if (pred(sp[i]) && tiled(sp[i])) { int a, b; dims(sp[i], a, b); maxl = std::max(maxl, a + b); }
            if (maxl < 0) return;
            lv.resize(maxl + 1);
            for (int l = 0; l <= maxl; ++l) {
                lv[l].off = items.size();
                // thin levels: one tile per warp (the launch is a single wave and its duration the latency of the
                // tiles a warp runs back to back); fat levels: up to TILES_PER_CTA tiles per CTA
                uint64_t total = 0;
                for (int pass = 0; pass < 2; ++pass) {
                    const uint64_t per_cta = total <= 8ull * 3 * 148 ? 8 : total <= 16ull * 3 * 148 ? 16 : TILES_PER_CTA;
                    for (uint32_t i = 0; i < ck.nspaces; ++i) {
                        if (!pred(sp[i]) || !tiled(sp[i])) continue;
                        int kbA, kbB;
                        dims(sp[i], kbA, kbB);
                        if (l > kbA + kbB) continue;
                        need_hs(kbA); need_hs(kbB);
                        for (int lA = std::max(0, l - kbB); lA <= std::min(kbA, l); ++lA) {
                            const int lB = l - lA;
                            const uint64_t nA = hs_lvl[kbA][lA + 1] - hs_lvl[kbA][lA];
                            const uint64_t nB = hs_lvl[kbB][lB + 1] - hs_lvl[kbB][lB];
                            const uint64_t T = nA * ((nB + 7) / 8);
                            if (pass == 0) { total += T; continue; }
                            for (uint64_t t0 = 0; t0 < T; t0 += per_cta)
                                items.push_back({i, (uint32_t)lA | ((uint32_t)lB << 8) | ((uint32_t)std::min<uint64_t>(per_cta, T - t0) << 16), (uint32_t)t0});
                        }
                    }
                }
                lv[l].cnt = (uint32_t)(items.size() - lv[l].off);
            }
        };
        // adjoint pass with fused group-B statistics: CTA = G row groups x C column blocks of one (lA, lB) split
        auto levels_of_adjb = [&](std::vector<Range>& lv) {
            int maxl = -1;
            for (uint32_t i = 0; i < ck.nspaces; ++i)
                if (fused_b(sp[i]) && !rb_space(sp[i])) maxl = std::max(maxl, (int)sp[i].KA - 4 + (int)sp[i].KB);
            if (maxl < 0) return;
            lv.resize(maxl + 1);
            for (int l = 0; l <= maxl; ++l) {
                lv[l].off = items.size();
                for (uint32_t i = 0; i < ck.nspaces; ++i) {
                    if (!fused_b(sp[i]) || rb_space(sp[i])) continue;
                    const int kbA = sp[i].KA - 4, kbB = sp[i].KB;
                    if (l > kbA + kbB) continue;
                    need_hs(kbA); need_hs(kbB);
                    uint32_t base[MMH_MAX_BITS + 1];
                    adjb_slots(kbA, base);
                    for (int lA = std::max(0, l - kbB); lA <= std::min(kbA, l); ++lA) {
                        const int lB = l - lA;
                        const uint32_t nA = hs_lvl[kbA][lA + 1] - hs_lvl[kbA][lA];
                        const uint32_t nB = hs_lvl[kbB][lB + 1] - hs_lvl[kbB][lB];
                        const uint32_t nBg = (nB + 7) / 8;
                        uint32_t C = 32;
                        if (nA < 32) { C = 1; while (C < nA) C <<= 1; }
                        const uint32_t G = 32 / C, nch = nA < 32 ? 1 : (nA + 31) / 32;
                        for (uint32_t jB = 0; jB < nBg; jB += G)
                            for (uint32_t k = 0; k < nch; ++k)
                                items.push_back({i, (uint32_t)lA | ((uint32_t)lB << 8), jB, k | ((base[lA] + k) << 16)});
                    }
                }
                lv[l].cnt = (uint32_t)(items.size() - lv[l].off);
            }
        };
        // row-block kernel: a CTA is a block of 2^12 states = R rows x 2^KI columns; launch levels over the outer bits
        {
            int maxl = -1;
            auto outer = [&](const SpaceDev& s) { return (int)s.KB + std::max(0, (int)s.KA - RB_KI); };
            for (uint32_t i = 0; i < ck.nspaces; ++i) if (rb_space(sp[i])) maxl = std::max(maxl, outer(sp[i]));
            ck.main_rb.resize(maxl + 1);
            for (int l = 0; l <= maxl; ++l) {
                ck.main_rb[l].off = items.size();
                // fat levels: up to RB_MAXBLK blocks per CTA (the per-CTA tables are loaded once)
                uint64_t total = 0;
                for (int pass = 0; pass < 2; ++pass) {
                    const uint32_t per_cta = (uint32_t)std::min<uint64_t>(RB_MAXBLK, std::max<uint64_t>(1, total / (148ull * 8)));
                    for (uint32_t i = 0; i < ck.nspaces; ++i) {
                        if (!rb_space(sp[i])) continue;
                        const int KO = outer(sp[i]);
                        if (l > KO) continue;
                        const int KI = std::min<int>(sp[i].KA, RB_KI);
                        need_hs(KO); need_hs(KI - 2);
                        const uint32_t nL = hs_lvl[KO][l + 1] - hs_lvl[KO][l], R = 1u << (RB_KI - KI);
                        if (pass == 0) { total += (nL + R - 1) / R; continue; }
                        for (uint32_t f = 0; f < nL; f += R * per_cta)
                            items.push_back({i, (uint32_t)l | (std::min(R * per_cta, nL - f) << 8), f});
                    }
                }
                ck.main_rb[l].cnt = (uint32_t)(items.size() - ck.main_rb[l].off);
            }
        }
        levels_of(is_main, ck.main_lv);
        levels_of(is_sec, ck.sec_lv);
        levels_of_t([&](const SpaceDev& s) { return is_main(s) && !wide_t(s); }, ck.main_lvt);
        levels_of_t([&](const SpaceDev& s) { return is_main(s) && !fused_b(s) && !wide_t(s); }, ck.main_lvt_adj);
        levels_of_t(wide_t, ck.main_lvw);
        levels_of_adjb(ck.main_lvt_adjb);
        levels_of_t(is_sec, ck.sec_lvt);
        ck.st_a.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i)
            if (sp[i].kind == K_JOINT) {
                const uint32_t nblk = std::max<uint32_t>(1u, (1u << sp[i].KA) >> 6);     // chunks of 64 uA
                for (uint32_t sl = 0; sl < sp[i].slices; ++sl)
                    for (uint32_t b = 0; b < nblk; ++b) items.push_back({i, b, sl});
            }
        ck.st_a.cnt = (uint32_t)(items.size() - ck.st_a.off);
        ck.st_ar.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i)
            if (sp[i].kind == K_JOINT)
                for (uint32_t g = 0; g < 2; ++g) {
                    const int KG = g ? sp[i].KB : sp[i].KA;
                    const uint64_t len = (uint64_t)(KG + 1) << KG;
                    for (uint64_t t = 0; t < len; t += 1024) items.push_back({i, g, (uint32_t)(t / 1024)});
                }
        ck.st_ar.cnt = (uint32_t)(items.size() - ck.st_ar.off);
        ck.st_b.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i)
            if (sp[i].kind == K_JOINT && !fused_b(sp[i]) && !narrow_b(sp[i]))
                for (uint32_t sl = 0; sl < sp[i].slicesB; ++sl)
                    for (uint32_t u = 0; u < (1u << sp[i].KB); ++u) items.push_back({i, u, sl});
        ck.st_b.cnt = (uint32_t)(items.size() - ck.st_b.off);
        ck.st_bn.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i)
            if (narrow_b(sp[i]))
                for (uint32_t u = 0; u < (1u << sp[i].KB); u += 32) items.push_back({i, u, 0u});
        ck.st_bn.cnt = (uint32_t)(items.size() - ck.st_bn.off);
        ck.pf_lo.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i)
            if (is_prod(sp[i])) {
                const uint32_t nch = std::max<uint32_t>(1u, (1u << sp[i].splitA) >> 7);     // blocks of 128 lo
                for (uint32_t sl = 0; sl < sp[i].slices; ++sl)
                    for (uint32_t c = 0; c < nch; ++c) items.push_back({i, c, sl});
            }
        ck.pf_lo.cnt = (uint32_t)(items.size() - ck.pf_lo.off);
        ck.fin.off = items.size();
        for (uint32_t i = 0; i < ck.nspaces; ++i) {
            if (is_prod(sp[i])) {
                for (uint32_t u = 0; u < (1u << sp[i].splitA); u += FIN_U) items.push_back({i, 2u, u});
                for (uint32_t u = 0; u < (1u << (sp[i].KA - sp[i].splitA)); u += FIN_U) items.push_back({i, 3u, u});
                continue;
            }
            // groups with 2^15 sub-states or more (wide pairs: thousands of items per space) take coarser items
            const uint32_t UA = sp[i].KA >= 15 ? FIN_U_WIDE : FIN_U, UB = sp[i].KB >= 15 ? FIN_U_WIDE : FIN_U;
            for (uint32_t u = 0; u < (1u << sp[i].KA); u += UA) items.push_back({i, 0u, u, UA});
            if (sp[i].kind == K_JOINT)
                for (uint32_t u = 0; u < (1u << sp[i].KB); u += UB) items.push_back({i, 1u, u, UB});
        }
        ck.fin.cnt = (uint32_t)(items.size() - ck.fin.off);
        h->chunks.push_back(std::move(ck));
    }
    h->st.n_spaces = (int64_t)spaces.size();
    h->st.n_chunks = (int64_t)h->chunks.size();

    // ---- upload -----------------------------------------------------------------------------------------
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    h->nsm = prop.multiProcessorCount;
    h->fin_ctas = prop.multiProcessorCount * 3;          // three k_finish CTAs fit one SM (58 KB of shared memory each)
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    CK(up((void**)&h->d_spaces, spaces.data(), spaces.size() * sizeof(SpaceDev)));
    CK(up((void**)&h->d_lists, lists.data(), lists.size() * sizeof(uint32_t)));
    CK(up((void**)&h->d_items, items.data(), items.size() * sizeof(Item)));
    CK(up((void**)&h->d_hs, hs_all.data(), hs_all.size() * sizeof(uint32_t)));
    {
        std::vector<uint32_t> hsidx((size_t)MMH_MAX_BITS * 32, 0u);
        for (int kb = 0; kb < MMH_MAX_BITS; ++kb)
            for (size_t l = 0; l < hs_lvl[kb].size() && l < 32; ++l) hsidx[(size_t)kb * 32 + l] = (uint32_t)(hs_off[kb] + hs_lvl[kb][l]);
        CK(up((void**)&h->d_hsidx, hsidx.data(), hsidx.size() * sizeof(uint32_t)));
    }
    {
        // Level lists of the row-block kernel, [bits][2^10]: popcount-sorted like d_hs, but inside a level the entries
        // are dealt round-robin from the eight residue classes modulo 8, so that eight consecutive lanes touch eight
        // different 16-byte bank groups of the shared-memory block (mmh_rowblock.cuh)
        std::vector<uint16_t> rblv((size_t)(RB_KI - 1) << (RB_KI - 2), 0);
        for (int kg = 0; kg <= RB_KI - 2; ++kg) {
            size_t pos = (size_t)kg << (RB_KI - 2);
            for (int l = 0; l <= kg; ++l) {
                std::vector<uint16_t> bucket[8];
                for (uint32_t c = 0; c < (1u << kg); ++c) if (__builtin_popcount(c) == l) bucket[c & 7u].push_back((uint16_t)c);
                size_t taken[8] = {};
                for (bool any = true; any;) {
                    any = false;
                    for (int k = 0; k < 8; ++k)
                        if (taken[k] < bucket[k].size()) { rblv[pos++] = bucket[k][taken[k]++]; any = true; }
                }
            }
        }
        CK(up((void**)&h->d_rblv, rblv.data(), rblv.size() * sizeof(uint16_t)));
    }
    CK(up((void**)&h->d_cls, cls.data(), cls.size()));
    CK(up((void**)&h->d_cnt, cnt_dm2.data(), NR * sizeof(double)));
    const size_t npar = (size_t)h->n_tot * (h->n_tot + 2);
    CK(cudaMalloc((void**)&h->d_par, sizeof(EvalPar)));
    CK(cudaMalloc((void**)&h->d_params, npar * sizeof(double)));
    if (const char* e = std::getenv("MMH_GRAPH")) h->use_graph = std::atoi(e) != 0;
    if (const char* e = std::getenv("MMH_STREAMS")) h->ns = std::max(1, std::min((int)mmh_handle::NS, std::atoi(e)));
    h->ns = (int)std::max<size_t>(1, std::min<size_t>((size_t)h->ns, h->chunks.size()));
    for (int q = 0; q < h->ns; ++q) {
        CK(cudaMalloc((void**)&h->d_scratch_s[q], std::max<uint64_t>(max_scratch, 4) * sizeof(double)));
        CK(cudaStreamCreateWithFlags(&h->side[q], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_side[q], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&h->ev_prep, cudaEventDisableTiming));
    h->d_scratch = h->d_scratch_s[0];
    CK(cudaMalloc((void**)&h->d_logp, std::max<int64_t>(n_dat, 1) * sizeof(double)));
    CK(cudaMemset(h->d_logp, 0, std::max<int64_t>(n_dat, 1) * sizeof(double)));
    CK(cudaMalloc((void**)&h->d_partial, (size_t)h->ns * h->fin_ctas * NACC * NR * NR * sizeof(double)));
    CK(cudaMalloc((void**)&h->d_partial2, (size_t)RED_SLICES * NACC * NR * NR * sizeof(double)));
    CK(cudaMalloc((void**)&h->d_diracc, (size_t)h->ns * 2 * NR * sizeof(double)));
    CK(cudaMalloc((void**)&h->d_tdir, (size_t)h->ns * std::max<uint32_t>(h->max_joints, 1) * 2 * sizeof(double)));
    CK(cudaMalloc((void**)&h->d_out, (npar + 1) * sizeof(double)));
    CK(cudaMallocHost((void**)&h->h_out, (npar + 1) * sizeof(double)));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaFuncSetAttribute(k_finish<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FIN_SMEM));
    CK(cudaFuncSetAttribute(k_finish<MAXG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FIN_SMEM));
    CK(cudaFuncSetAttribute(k_solve_rb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RbSmem)));
    CK(cudaFuncSetAttribute(k_solve_rb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RbSmem)));
    h->st.scratch_bytes = (double)max_scratch * 8.0 * h->ns;
    return MMH_OK;
}

// Enqueue one evaluation on the handle's streams (asynchronous).  d_params is a DEVICE pointer.
static int enqueue_eval(mmh_handle* h, const double* d_params, double w0, double w1, int want_grad, int64_t* n_launches)
{
    cudaStream_t st = h->stream;
    const int NS = h->ns;                               // slots in use (profile mode runs everything in slot 0)
    const int ns = h->profile ? 1 : h->ns;              // profile mode serialises everything on the main stream
    int64_t launches = 0;
    // optional per-class timing (profile mode): CUDA events around every launch group
    size_t evn = 0;
    std::vector<int> evcls;
    auto tick = [&](int cls) {
        if (!h->profile) return;
        if (evn >= h->evpool.size()) { cudaEvent_t e; cudaEventCreate(&e); h->evpool.push_back(e); }
        cudaEventRecord(h->evpool[evn++], st);
        evcls.push_back(cls);
    };
    tick(5);
    k_prep<<<1, 1024, 0, st>>>(d_params, h->n_tot, h->d_par); ++launches;
    if (want_grad) {
        CK(cudaMemsetAsync(h->d_partial, 0, (size_t)NS * h->fin_ctas * NACC * NR * NR * sizeof(double), st));
        CK(cudaMemsetAsync(h->d_diracc, 0, (size_t)NS * 2 * NR * sizeof(double), st));
    }
    const cudaStream_t main_stream = st;
    if (ns > 1) {
        CK(cudaEventRecord(h->ev_prep, main_stream));
        for (int q = 0; q < ns; ++q) CK(cudaStreamWaitEvent(h->side[q], h->ev_prep, 0));
    }
    const size_t part_stride = (size_t)h->fin_ctas * NACC * NR * NR;
    const size_t tdir_stride = (size_t)std::max<uint32_t>(h->max_joints, 1) * 2;
    size_t chunk_no = 0;
    for (const ChunkPlan& ck : h->chunks) {
        const int slot = ns > 1 ? (int)(chunk_no % ns) : 0;
        ++chunk_no;
        st = ns > 1 ? h->side[slot] : main_stream;
        double* S = h->d_scratch_s[slot];
        double* d_partial = h->d_partial + slot * part_stride;
        double* d_diracc = h->d_diracc + (size_t)slot * 2 * NR;
        double* d_tdir = h->d_tdir + slot * tdir_stride;
        const SpaceDev* sp = h->d_spaces + ck.space0;
        auto small = [&](const Range& r, bool adj) {
            if (!r.cnt) return;
            const uint32_t grid = (r.cnt + 7) / 8;
            if (adj) k_solve_small<true><<<grid, 256, 0, st>>>(sp, h->d_lists + r.off, r.cnt, S);
            else     k_solve_small<false><<<grid, 256, 0, st>>>(sp, h->d_lists + r.off, r.cnt, S);
            ++launches;
        };
        auto small4 = [&](const Range& r, bool adj) {
            if (!r.cnt) return;
            const uint32_t grid = (r.cnt + 7) / 8;
            if (adj) k_solve_small4<true><<<grid, 256, 0, st>>>(sp, h->d_lists + r.off, r.cnt, S);
            else     k_solve_small4<false><<<grid, 256, 0, st>>>(sp, h->d_lists + r.off, r.cnt, S);
            ++launches;
        };
        auto big = [&](const std::vector<Range>& lv, bool adj) {
            const int L = (int)lv.size();
            for (int q = 0; q < L; ++q) {
                const Range& r = lv[adj ? L - 1 - q : q];
                if (!r.cnt) continue;
                if (adj) k_solve_big4<true><<<r.cnt, 256, 0, st>>>(sp, h->d_items + r.off, h->d_hs, S);
                else     k_solve_big4<false><<<r.cnt, 256, 0, st>>>(sp, h->d_items + r.off, h->d_hs, S);
                ++launches;
            }
        };
        auto bigt = [&](const std::vector<Range>& lv, bool adj) {
            const int L = (int)lv.size();
            for (int q = 0; q < L; ++q) {
                const Range& r = lv[adj ? L - 1 - q : q];
                if (!r.cnt) continue;
                if (adj) k_solve_tile<true><<<r.cnt, 256, 0, st>>>(sp, h->d_items + r.off, h->d_hs, h->d_hsidx, S);
                else     k_solve_tile<false><<<r.cnt, 256, 0, st>>>(sp, h->d_items + r.off, h->d_hs, h->d_hsidx, S);
                ++launches;
            }
        };
        auto bigw = [&](bool adj) {
            const int L = (int)ck.main_lvw.size();
            for (int q = 0; q < L; ++q) {
                const Range& r = ck.main_lvw[adj ? L - 1 - q : q];
                if (!r.cnt) continue;
                if (adj) k_solve_tile_w<true><<<r.cnt, 256, 0, st>>>(sp, h->d_items + r.off, h->d_hs, h->d_hsidx, S);
                else     k_solve_tile_w<false><<<r.cnt, 256, 0, st>>>(sp, h->d_items + r.off, h->d_hs, h->d_hsidx, S);
                ++launches;
            }
        };
        auto rowblk = [&](bool adj) {
            const int L = (int)ck.main_rb.size();
            for (int q = 0; q < L; ++q) {
                const Range& r = ck.main_rb[adj ? L - 1 - q : q];
                if (!r.cnt) continue;
                if (adj) k_solve_rb<true><<<r.cnt, RB_T, sizeof(RbSmem), st>>>(sp, h->d_items + r.off, h->d_hs, h->d_hsidx, h->d_rblv, S);
                else     k_solve_rb<false><<<r.cnt, RB_T, sizeof(RbSmem), st>>>(sp, h->d_items + r.off, h->d_hs, h->d_hsidx, h->d_rblv, S);
                ++launches;
            }
        };
        tick(0);
        k_setup<<<ck.setup.cnt, 256, 0, st>>>(sp, h->d_items + ck.setup.off, h->d_par, S); ++launches;
        if (ck.setup_wide.cnt) { k_setup_wide<<<ck.setup_wide.cnt, 1024, 0, st>>>(sp, h->d_items + ck.setup_wide.off, h->d_par, S); ++launches; }
        if (ck.diag.cnt) { k_diag_prod<<<ck.diag.cnt, 256, 0, st>>>(sp, h->d_items + ck.diag.off, S); ++launches; }
        if (ck.rb_list.cnt) { k_rb_tables<<<ck.rb_list.cnt, 256, 0, st>>>(sp, h->d_lists + ck.rb_list.off, h->d_par, S); ++launches; }
        tick(1);
        small(ck.pre, false); small4(ck.pre4, false);
        small(ck.main_small, false); small4(ck.main_small4, false);
        big(ck.main_lv, false); bigt(ck.main_lvt, false); bigw(false); rowblk(false);
        small(ck.sec_small, false); small4(ck.sec_small4, false);
        big(ck.sec_lv, false); bigt(ck.sec_lvt, false);
        tick(5);
        if (ck.logp.cnt) { k_logp<<<(ck.logp.cnt + 127) / 128, 128, 0, st>>>(sp, h->d_lists + ck.logp.off, ck.logp.cnt, S, h->d_logp); ++launches; }
        if (!want_grad) continue;
        tick(2);
        small(ck.sec_small, true); small4(ck.sec_small4, true);
        big(ck.sec_lv, true); bigt(ck.sec_lvt, true);
        tick(5);
        if (ck.joints.cnt) {
            k_direct<<<(ck.joints.cnt + 255) / 256, 256, 0, st>>>(sp, h->d_lists + ck.joints.off, ck.joints.cnt, S, d_tdir);
            k_direct_acc<<<dim3(h->n_tot, 2), 256, 0, st>>>(sp, h->d_lists + ck.joints.off, ck.joints.cnt, d_tdir, w1, d_diracc);
            launches += 2;
        }
        tick(2);
        small(ck.main_small, true); small4(ck.main_small4, true);
        big(ck.main_lv, true); bigt(ck.main_lvt_adj, true); bigw(true);
        for (int q = (int)ck.main_lvt_adjb.size() - 1; q >= 0; --q) {
            const Range& r = ck.main_lvt_adjb[q];
            if (!r.cnt) continue;
            k_solve_tile_adjb<<<r.cnt, 256, 0, st>>>(sp, h->d_items + r.off, h->d_hs, h->d_hsidx, S);
            ++launches;
        }
        rowblk(true);
        small(ck.pre, true); small4(ck.pre4, true);
        tick(3);
        if (ck.st_a.cnt) {
            k_stats_a<<<(ck.st_a.cnt + 7) / 8, 256, 0, st>>>(sp, h->d_items + ck.st_a.off, ck.st_a.cnt, S);
            if (ck.st_b.cnt) k_stats_b<<<(ck.st_b.cnt + 7) / 8, 256, 0, st>>>(sp, h->d_items + ck.st_b.off, ck.st_b.cnt, S);
            if (ck.st_bn.cnt) { k_stats_b_narrow<<<(ck.st_bn.cnt + 7) / 8, 256, 0, st>>>(sp, h->d_items + ck.st_bn.off, ck.st_bn.cnt, S); ++launches; }
            k_stats_reduce<<<ck.st_ar.cnt, 1024, 0, st>>>(sp, h->d_items + ck.st_ar.off, S);
            launches += 3;
        }
        tick(6);
        if (ck.pf_lo.cnt) {
            k_pf<<<ck.pf_lo.cnt, 32 * PF_WARPS, 0, st>>>(sp, h->d_items + ck.pf_lo.off, S);
            ++launches;
        }
        tick(4);
        const size_t fin_smem = FIN_SMEM;
        const int fin_grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)h->fin_ctas, (ck.fin.cnt + 2 * FIN_WARPS - 1) / (2 * FIN_WARPS)));
        if (ck.wide) k_finish<MAXG><<<fin_grid, FIN_WARPS * 32, fin_smem, st>>>(sp, h->d_items + ck.fin.off, ck.fin.cnt, S, w0, w1, d_partial);
        else         k_finish<MAXT><<<fin_grid, FIN_WARPS * 32, fin_smem, st>>>(sp, h->d_items + ck.fin.off, ck.fin.cnt, S, w0, w1, d_partial);
        ++launches;
    }
    st = main_stream;
    if (ns > 1)
        for (int q = 0; q < ns; ++q) {
            CK(cudaEventRecord(h->ev_side[q], h->side[q]));
            CK(cudaStreamWaitEvent(main_stream, h->ev_side[q], 0));
        }
    tick(5);
    if (want_grad) {
        k_reduce_partials<<<dim3(NACC * NR * NR / 256, RED_SLICES), 256, 0, st>>>(h->d_partial, NS * h->fin_ctas, h->d_partial2);
        ++launches;
    }
    k_final<<<1, 1024, 0, st>>>(h->d_partial2, RED_SLICES, h->d_diracc, NS, h->d_logp, h->d_cls, h->n_dat, h->d_cnt, w0, w1,
                                h->n_tot, want_grad, h->d_out);
    ++launches;
    tick(-1);
    CK(cudaGetLastError());
    *n_launches = launches;
    if (h->profile) {
        CK(cudaStreamSynchronize(st));
        for (int c = 0; c < 8; ++c) h->st.class_ms[c] = 0.0;
        for (size_t i = 0; i + 1 < evn; ++i) {
            float ms = 0.f;
            if (evcls[i] >= 0 && cudaEventElapsedTime(&ms, h->evpool[i], h->evpool[i + 1]) == cudaSuccess)
                h->st.class_ms[evcls[i]] += ms;
        }
    }
    return MMH_OK;
}

// Launch one evaluation.  The launch sequence is static (it depends only on the dataset), so it is captured once
// per (parameter buffer, class weights, value/grad) into a CUDA graph -- side streams included -- and replayed:
// ~4 900 launches at n = 25 / 100 000 patients, 53 for LUAD, become one graph launch.
static int run_eval(mmh_handle* h, const double* d_params, double w0, double w1, int want_grad)
{
    CK(cudaSetDevice(h->device));
    int64_t launches = 0;
    if (h->profile || !h->use_graph) {
        CK(cudaEventRecord(h->ev0, h->stream));
        int rc = enqueue_eval(h, d_params, w0, w1, want_grad, &launches);
        if (rc != MMH_OK) return rc;
        CK(cudaEventRecord(h->ev1, h->stream));
        h->st.n_launches = launches;
        return MMH_OK;
    }
    mmh_handle::GraphEntry* hit = nullptr;
    for (auto& g : h->graphs)
        if (g.params == d_params && g.w0 == w0 && g.w1 == w1 && g.want_grad == want_grad) hit = &g;
    if (!hit) {
        if (h->graphs.size() >= 4) { cudaGraphExecDestroy(h->graphs.front().exec); h->graphs.erase(h->graphs.begin()); }
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_eval(h, d_params, w0, w1, want_grad, &launches);
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        if (rc != MMH_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(MMH_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(MMH_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        h->graphs.push_back({d_params, w0, w1, want_grad, exec, launches});
        hit = &h->graphs.back();
    }
    CK(cudaEventRecord(h->ev0, h->stream));
    CK(cudaGraphLaunch(hit->exec, h->stream));
    CK(cudaEventRecord(h->ev1, h->stream));
    h->st.n_launches = hit->launches;
    return MMH_OK;
}

#ifdef RB_TIMING
extern "C" int mmh_debug_rb_timing(unsigned long long* out16, int reset)
{
    if (cudaMemcpyFromSymbol(out16, rb_timing, 16 * sizeof(unsigned long long)) != cudaSuccess) return MMH_ECUDA;
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(rb_timing, z, sizeof(z)); }
    return MMH_OK;
}
#endif

static void class_weights(const mmh_handle* h, double perc_met, double& w0, double& w1)
{
    // regularized_optimization.py:256-262
    const double n_em = (double)h->n_em, n_nm = (double)h->n_dat - n_em;
    const double w = (n_em * n_nm != 0.0) ? perc_met * n_nm / ((1.0 - perc_met) * n_em) : 1.0;
    const double n_full = w * n_em + n_nm;
    w0 = 1.0 / n_full;
    w1 = w / n_full;
}


// ---- NCCL (loaded lazily with dlopen: the library has no link-time dependency on it) ------------------------
namespace {
struct ncclUniqueIdT { char internal[MMH_NCCL_ID_BYTES]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueIdT*) = nullptr;
    int (*CommInitRank)(void**, int, ncclUniqueIdT, int) = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;      // ncclFloat64, ncclSum (nccl.h; stable since NCCL 2.0)

NcclApi& nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        const char* names[] = {std::getenv("MMH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm) continue;
            a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);   // an already loaded libnccl.so.2 (e.g. torch's) is reused
            if (a.lib) break;
        }
        if (!a.lib) return a;
        auto sym = [&](const char* n) { return dlsym(a.lib, n); };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommInitAll = (decltype(a.CommInitAll))sym("ncclCommInitAll");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart &&
               a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}
int nccl_fail(const char* what, int rc)
{
    return fail(MMH_ENCCL, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error"));
}
}  // namespace

#define NK(call)                                                            \
    do {                                                                    \
        int r_ = (call);                                                    \
        if (r_ != 0) return nccl_fail(#call, r_);                           \
    } while (0)

extern "C" int mmh_nccl_unique_id(char id[MMH_NCCL_ID_BYTES])
{
    if (!id) return fail(MMH_EINVAL, "mmh_nccl_unique_id: null argument");
    if (!nccl().ok) return fail(MMH_ENCCL, "libnccl.so.2 could not be loaded (set MMH_NCCL_LIB)");
    ncclUniqueIdT u;
    NK(nccl().GetUniqueId(&u));
    std::memcpy(id, u.internal, MMH_NCCL_ID_BYTES);
    return MMH_OK;
}

extern "C" int mmh_comm_init(mmh_handle* h, const char id[MMH_NCCL_ID_BYTES], int nranks, int rank)
{
    if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(MMH_EINVAL, "mmh_comm_init: bad arguments");
    if (!nccl().ok) return fail(MMH_ENCCL, "libnccl.so.2 could not be loaded (set MMH_NCCL_LIB)");
    if (h->comm) return fail(MMH_EINVAL, "mmh_comm_init: the handle already has a communicator");
    CK(cudaSetDevice(h->device));
    ncclUniqueIdT u;
    std::memcpy(u.internal, id, MMH_NCCL_ID_BYTES);
    void* comm = nullptr;
    NK(nccl().CommInitRank(&comm, nranks, u, rank));
    h->comm = comm;
    h->comm_owned = true;
    return MMH_OK;
}

extern "C" int mmh_comm_destroy(mmh_handle* h)
{
    if (!h) return fail(MMH_EINVAL, "mmh_comm_destroy: null argument");
    if (h->comm) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        if (h->comm_owned) nccl().CommDestroy(h->comm);
        h->comm = nullptr;
    }
    return MMH_OK;
}

// the shard results add up: one all-reduce of the result vector on the stream that produced it
static int reduce_result(mmh_handle* h, size_t len)
{
    if (!h->comm) return MMH_OK;
    NK(nccl().AllReduce(h->d_out, h->d_out, len, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream));
    CK(cudaEventRecord(h->ev1, h->stream));          // the device time of an evaluation includes its collective
    return MMH_OK;
}

extern "C" int mmh_eval_weighted(mmh_handle* h, const double* params, double w_type0, double w_other,
                                 int want_grad, double* out_host, double* out_dev)
{
    if (!h || !params) return fail(MMH_EINVAL, "mmh_eval_weighted: null argument");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->d_params, params, (size_t)h->n_tot * (h->n_tot + 2) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int rc = run_eval(h, h->d_params, w_type0, w_other, want_grad);
    if (rc != MMH_OK) return rc;
    const size_t len = want_grad ? (size_t)h->n_tot * (h->n_tot + 2) + 1 : 1;
    rc = reduce_result(h, len);
    if (rc != MMH_OK) return rc;
    if (out_dev) CK(cudaMemcpyAsync(out_dev, h->d_out, len * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (out_host) CK(cudaMemcpyAsync(h->h_out, h->d_out, len * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->st.last_ms = ms;
    if (out_host) std::memcpy(out_host, h->h_out, len * sizeof(double));
    return MMH_OK;
}

extern "C" int mmh_eval_device(mmh_handle* h, const double* d_params, double w_type0, double w_other,
                               int want_grad, double* d_out)
{
    if (!h || !d_params || !d_out) return fail(MMH_EINVAL, "mmh_eval_device: null argument");
    int rc = run_eval(h, d_params, w_type0, w_other, want_grad);
    if (rc != MMH_OK) return rc;
    const size_t len = want_grad ? (size_t)h->n_tot * (h->n_tot + 2) + 1 : 1;
    rc = reduce_result(h, len);
    if (rc != MMH_OK) return rc;
    CK(cudaMemcpyAsync(d_out, h->d_out, len * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return MMH_OK;
}

extern "C" int mmh_sync(mmh_handle* h)
{
    if (!h) return fail(MMH_EINVAL, "mmh_sync: null argument");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->st.last_ms = ms;
    return MMH_OK;
}

extern "C" int mmh_set_profile(mmh_handle* h, int on)
{
    if (!h) return fail(MMH_EINVAL, "mmh_set_profile: null argument");
    h->profile = on ? 1 : 0;
    return MMH_OK;
}

extern "C" int mmh_value_grad(mmh_handle* h, const double* params, double perc_met, double* score, double* grad)
{
    if (!h || !params || !score || !grad) return fail(MMH_EINVAL, "mmh_value_grad: null argument");
    double w0, w1;
    class_weights(h, perc_met, w0, w1);
    const size_t npar = (size_t)h->n_tot * (h->n_tot + 2);
    std::vector<double> buf(npar + 1);
    int rc = mmh_eval_weighted(h, params, w0, w1, 1, buf.data(), nullptr);
    if (rc != MMH_OK) return rc;
    *score = buf[0];
    std::memcpy(grad, buf.data() + 1, npar * sizeof(double));
    return MMH_OK;
}

extern "C" int mmh_value(mmh_handle* h, const double* params, double perc_met, double* score)
{
    if (!h || !params || !score) return fail(MMH_EINVAL, "mmh_value: null argument");
    double w0, w1;
    class_weights(h, perc_met, w0, w1);
    return mmh_eval_weighted(h, params, w0, w1, 0, score, nullptr);
}

extern "C" int mmh_per_patient(mmh_handle* h, const double* params, double* logp)
{
    if (!h || !params || !logp) return fail(MMH_EINVAL, "mmh_per_patient: null argument");
    double s;
    int rc = mmh_eval_weighted(h, params, 1.0, 1.0, 0, &s, nullptr);
    if (rc != MMH_OK) return rc;
    if (h->n_dat) CK(cudaMemcpy(logp, h->d_logp, (size_t)h->n_dat * sizeof(double), cudaMemcpyDeviceToHost));
    return MMH_OK;
}

// Per-row log-likelihood AND gradient (test hook, the per-patient `_g_coupled_*` / `_grad_*` of metmhn/jx/likelihood.py):
// every requested row is evaluated as its own one-row dataset with unit weights.  Milliseconds per row -- not a data path.
extern "C" int mmh_per_patient_grads(mmh_handle* h, const double* params, int64_t first_row, int64_t n_rows,
                                     double* logp, double* grads)
{
    if (!h || !params || !logp || first_row < 0 || n_rows < 0 || first_row + n_rows > h->n_dat)
        return fail(MMH_EINVAL, "mmh_per_patient_grads: bad arguments");
    const int width = 2 * h->n + 3;
    const size_t npar = (size_t)h->n_tot * (h->n_tot + 2);
    std::vector<double> buf(npar + 1);
    for (int64_t r = 0; r < n_rows; ++r) {
        mmh_handle* one = nullptr;
        int rc = mmh_create(&one, h->n, h->rows.data() + (size_t)(first_row + r) * width, 1, width, h->device, 0);
        if (rc != MMH_OK) return rc;
        rc = mmh_eval_weighted(one, params, 1.0, 1.0, grads ? 1 : 0, buf.data(), nullptr);
        mmh_destroy(one);
        if (rc != MMH_OK) return rc;
        logp[r] = buf[0];
        if (grads) std::memcpy(grads + (size_t)r * npar, buf.data() + 1, npar * sizeof(double));
    }
    return MMH_OK;
}

extern "C" int mmh_stats(mmh_handle* h, mmh_stats_t* out)
{
    if (!h || !out) return fail(MMH_EINVAL, "mmh_stats: null argument");
    *out = h->st;
    return MMH_OK;
}

extern "C" void mmh_destroy(mmh_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    mmh_comm_destroy(h);
    cudaFree(h->d_spaces); cudaFree(h->d_lists); cudaFree(h->d_items); cudaFree(h->d_hs); cudaFree(h->d_hsidx); cudaFree(h->d_rblv); cudaFree(h->d_cls);
    cudaFree(h->d_cnt); cudaFree(h->d_par); cudaFree(h->d_params); cudaFree(h->d_logp);
    for (int q = 0; q < mmh_handle::NS; ++q) {
        cudaFree(h->d_scratch_s[q]);
        if (h->side[q]) cudaStreamDestroy(h->side[q]);
        if (h->ev_side[q]) cudaEventDestroy(h->ev_side[q]);
    }
    if (h->ev_prep) cudaEventDestroy(h->ev_prep);
    cudaFree(h->d_partial); cudaFree(h->d_partial2); cudaFree(h->d_diracc); cudaFree(h->d_tdir); cudaFree(h->d_out);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t e : h->evpool) cudaEventDestroy(e);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    delete h;
}


// ---- one process, several GPUs ---------------------------------------------------------------------------
struct mmh_multi {
    std::vector<mmh_handle*> hs;
    int n_tot = 0;
    int64_t n_dat = 0, n_em = 0;
};

// work estimate of a row: lattice states times the measured cost per state (ps, B200) of the row's kind and size tier --
// the table of metmhn_b200/sharded.py (scripts/calibrate_cost.py)
static double row_cost(const int8_t* row, int n)
{
    int pt = 0, mt = 0;
    for (int e = 0; e < n; ++e) { pt += row[2 * e] != 0; mt += row[2 * e + 1] != 0; }
    const int seed = row[2 * n] != 0, typ = row[2 * n + 2];
    auto single = [](int k) { return std::ldexp(1.0, k) * (k >= 17 ? 85.0 : k >= 13 ? 112.0 : k >= 9 ? 232.0 : 800.0); };
    // pairs with fewer than 4 PT events or a tumour with more than 16 take the generic solve kernel
    const bool generic = pt < 4 || pt > 16 || mt > 16;
    auto pair = [generic](int k) { return std::ldexp(1.0, k) * (k < 13 ? 485.0 : generic ? (k >= 20 ? 195.0 : 590.0) : (k >= 20 ? 37.0 : 66.0)); };
    if (typ == 0 || typ == 1) return single(pt + seed);
    if (typ == 2) return single(mt + 1);
    if (typ == 3) return pair(pt + mt);
    return 1.0;
}

extern "C" double mmh_row_cost(const int8_t* row, int n_mut)
{
    return (row && n_mut >= 1 && n_mut <= MMH_MAX_MUT) ? row_cost(row, n_mut) : 0.0;
}

extern "C" void mmh_multi_destroy(mmh_multi* m)
{
    if (!m) return;
    for (mmh_handle* h : m->hs) mmh_destroy(h);
    delete m;
}

extern "C" int mmh_multi_create(mmh_multi** out, int n_mut, const int8_t* dat, int64_t n_dat, int64_t row_stride,
                                const int* device_ids, int n_devices, int64_t chunk_bytes)
{
    if (!out || !dat || !device_ids || n_devices < 1 || n_mut < 1 || n_mut > MMH_MAX_MUT || n_dat < 0 || row_stride < 2 * n_mut + 3)
        return fail(MMH_EINVAL, "mmh_multi_create: bad arguments");
    if (n_devices > 1 && !nccl().ok) return fail(MMH_ENCCL, "libnccl.so.2 could not be loaded (set MMH_NCCL_LIB)");
    const int n = n_mut, width = 2 * n + 3;
    // longest-processing-time assignment, deterministic (ties: lower row index first, lower device first)
    std::vector<int64_t> order((size_t)n_dat);
    std::vector<double> cost((size_t)n_dat);
    for (int64_t p = 0; p < n_dat; ++p) { order[(size_t)p] = p; cost[(size_t)p] = row_cost(dat + p * row_stride, n); }
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cost[(size_t)a] > cost[(size_t)b]; });
    std::vector<double> load((size_t)n_devices, 0.0);
    std::vector<std::vector<int8_t>> shard((size_t)n_devices);
    int64_t n_em = 0;
    for (int64_t p : order) {
        int best = 0;
        for (int d = 1; d < n_devices; ++d) if (load[(size_t)d] < load[(size_t)best]) best = d;
        load[(size_t)best] += cost[(size_t)p];
        const int8_t* row = dat + p * row_stride;
        shard[(size_t)best].insert(shard[(size_t)best].end(), row, row + width);
        n_em += row[2 * n];
    }
    mmh_multi* m = new mmh_multi();
    m->n_tot = n + 1; m->n_dat = n_dat; m->n_em = n_em;
    for (int d = 0; d < n_devices; ++d) {
        mmh_handle* h = nullptr;
        static const int8_t none = 0;
        const int8_t* rows = shard[(size_t)d].empty() ? &none : shard[(size_t)d].data();
        const int rc = mmh_create(&h, n_mut, rows, (int64_t)(shard[(size_t)d].size() / width), width, device_ids[d], chunk_bytes);
        if (rc != MMH_OK) { const std::string keep = g_err; mmh_multi_destroy(m); g_err = keep; return rc; }
        m->hs.push_back(h);
    }
    if (n_devices > 1) {
        std::vector<void*> comms((size_t)n_devices, nullptr);
        const int rc = nccl().CommInitAll(comms.data(), n_devices, device_ids);
        if (rc != 0) { mmh_multi_destroy(m); return nccl_fail("ncclCommInitAll", rc); }
        for (int d = 0; d < n_devices; ++d) { m->hs[(size_t)d]->comm = comms[(size_t)d]; m->hs[(size_t)d]->comm_owned = true; }
    }
    *out = m;
    return MMH_OK;
}

static int multi_eval(mmh_multi* m, const double* params, double perc_met, int want_grad, double* out)
{
    const double n_em = (double)m->n_em, n_nm = (double)m->n_dat - n_em;      // regularized_optimization.py:256-262
    const double w = (n_em * n_nm != 0.0) ? perc_met * n_nm / ((1.0 - perc_met) * n_em) : 1.0;
    const double n_full = w * n_em + n_nm, w0 = 1.0 / n_full, w1 = w / n_full;
    const size_t npar = (size_t)m->n_tot * (m->n_tot + 2), len = want_grad ? npar + 1 : 1;
    for (mmh_handle* h : m->hs) {
        CK(cudaSetDevice(h->device));
        CK(cudaMemcpyAsync(h->d_params, params, npar * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        const int rc = run_eval(h, h->d_params, w0, w1, want_grad);
        if (rc != MMH_OK) return rc;
    }
    if (m->hs.size() > 1) {
        NK(nccl().GroupStart());
        for (mmh_handle* h : m->hs)
            NK(nccl().AllReduce(h->d_out, h->d_out, len, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream));
        NK(nccl().GroupEnd());
    }
    mmh_handle* h0 = m->hs[0];
    CK(cudaSetDevice(h0->device));
    CK(cudaMemcpyAsync(h0->h_out, h0->d_out, len * sizeof(double), cudaMemcpyDeviceToHost, h0->stream));
    for (mmh_handle* h : m->hs) { CK(cudaSetDevice(h->device)); CK(cudaStreamSynchronize(h->stream)); }
    std::memcpy(out, h0->h_out, len * sizeof(double));
    return MMH_OK;
}

extern "C" int mmh_multi_value_grad(mmh_multi* m, const double* params, double perc_met, double* score, double* grad)
{
    if (!m || !params || !score || !grad) return fail(MMH_EINVAL, "mmh_multi_value_grad: null argument");
    const size_t npar = (size_t)m->n_tot * (m->n_tot + 2);
    std::vector<double> buf(npar + 1);
    const int rc = multi_eval(m, params, perc_met, 1, buf.data());
    if (rc != MMH_OK) return rc;
    *score = buf[0];
    std::memcpy(grad, buf.data() + 1, npar * sizeof(double));
    return MMH_OK;
}

extern "C" int mmh_multi_value(mmh_multi* m, const double* params, double perc_met, double* score)
{
    if (!m || !params || !score) return fail(MMH_EINVAL, "mmh_multi_value: null argument");
    return multi_eval(m, params, perc_met, 0, score);
}

// ---- learn_mhn inside the library (regularized_optimization.py:270-334) ------------------------------------------------
// out[0] = -score + lambda * penalty, out[1..] = -grad + lambda * penalty': `symmetric_penal` (:46-52) = the symmetrised group
// penalty on the off-diagonal theta pairs (:31-43) + smoothed L1 on d_p, d_m (:11-28), on the device so that the objective of
// an L-BFGS iteration is one stream of work and one 7 KB read-back.
__global__ void k_penalty(const double* __restrict__ params, int n_tot, double eps, double lambda, double* __restrict__ out)
{
    __shared__ double red[1024];
    const int sq = n_tot * n_tot;
    double pen = 0.0;
    for (int t = threadIdx.x; t < sq + 2 * n_tot; t += blockDim.x) {
        double dpen;
        if (t < sq) {
            const int i = t / n_tot, j = t % n_tot;
            if (i == j) { pen += sqrt(eps); dpen = 0.0; }                      // pair.sum() runs over the zeroed diagonal too
            else {
                const double a = params[t], b = params[j * n_tot + i];
                const double pr = sqrt(a * a + b * b - a * b + eps);
                pen += pr;
                dpen = (2.0 * a - b) / (2.0 * pr);
            }
        } else {
            const double d = params[t], r = sqrt(d * d + eps);
            pen += 2.0 * r;                                                     // the theta part is halved below
            dpen = d / r;
        }
        out[1 + t] = -out[1 + t] + lambda * dpen;
    }
    red[threadIdx.x] = pen;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = -out[0] + lambda * 0.5 * (red[0] - n_tot * sqrt(eps));
}

extern "C" int mmh_learn(mmh_handle* h, const double* x0, double perc_met, double w_penal, double eps, int64_t max_iter,
                         double ftol, double* x_out, double* f_out, int64_t* n_iter, int64_t* n_eval)
{
    if (!h || !x0 || !x_out) return fail(MMH_EINVAL, "mmh_learn: null argument");
    CK(cudaSetDevice(h->device));
    double w0, w1;
    class_weights(h, perc_met, w0, w1);
    const size_t npar = (size_t)h->n_tot * (h->n_tot + 2);
    std::vector<double> x(x0, x0 + npar);
    int rc_eval = MMH_OK;
    auto fun = [&](const double* xv, double* g) -> double {
        cudaError_t e = cudaMemcpyAsync(h->d_params, xv, npar * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) {
            rc_eval = run_eval(h, h->d_params, w0, w1, 1);
            if (rc_eval == MMH_OK) rc_eval = reduce_result(h, npar + 1);
            if (rc_eval != MMH_OK) return std::nan("");
            k_penalty<<<1, 1024, 0, h->stream>>>(h->d_params, h->n_tot, eps, w_penal, h->d_out);
            e = cudaMemcpyAsync(h->h_out, h->d_out, (npar + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { rc_eval = fail(MMH_ECUDA, std::string("mmh_learn: ") + cudaGetErrorString(e)); return std::nan(""); }
        std::memcpy(g, h->h_out + 1, npar * sizeof(double));
        return h->h_out[0];
    };
    const int iters = (int)std::min<int64_t>(std::max<int64_t>(max_iter, 0), 2000000000);
    const LbfgsResult r = lbfgs_minimize(fun, x, iters, ftol);
    if (r.status == -1) return rc_eval != MMH_OK ? rc_eval : fail(MMH_EINVAL, "mmh_learn: the objective is not finite at a trial point");
    std::memcpy(x_out, x.data(), npar * sizeof(double));
    if (f_out) *f_out = r.f;
    if (n_iter) *n_iter = r.iterations;
    if (n_eval) *n_eval = r.evaluations;
    return MMH_OK;
}

// ---- GPU Gillespie sampler (metmhn/simulations.py:8-147) ----------------------------------------------------
extern "C" int mmh_simulate(int n_mut, const double* params, int64_t n_sim, uint64_t seed, int device,
                            int8_t* geno, int8_t* order)
{
    if (!params || !geno || !order || n_mut < 1 || n_mut > MMH_MAX_MUT || n_sim < 0)
        return fail(MMH_EINVAL, "mmh_simulate: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(MMH_ECUDA, "mmh_simulate: no usable CUDA device (this library has no CPU fallback)");
    if (n_sim == 0) return MMH_OK;
    CK(cudaSetDevice(device));
    const int n_tot = n_mut + 1;
    const size_t npar = (size_t)n_tot * (n_tot + 2), width = (size_t)2 * n_mut + 1;
    double* d_params = nullptr;
    SimPar* d_par = nullptr;
    int8_t *d_geno = nullptr, *d_order = nullptr;
    auto cleanup = [&]() { cudaFree(d_params); cudaFree(d_par); cudaFree(d_geno); cudaFree(d_order); };
    cudaError_t e = cudaMalloc((void**)&d_params, npar * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_par, sizeof(SimPar));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_geno, (size_t)n_sim * width);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_order, (size_t)n_sim);
    if (e == cudaSuccess) e = cudaMemcpy(d_params, params, npar * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_sim_prep<<<1, 256>>>(d_params, n_tot, d_par);
        k_simulate<<<(unsigned)((n_sim + 127) / 128), 128>>>(d_par, n_tot, n_sim, seed, d_geno, d_order);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(geno, d_geno, (size_t)n_sim * width, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(order, d_order, (size_t)n_sim, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess)
        return fail(e == cudaErrorMemoryAllocation ? MMH_ENOMEM : MMH_ECUDA, std::string("mmh_simulate: ") + cudaGetErrorString(e));
    return MMH_OK;
}

extern "C" const char* mmh_last_error(void) { return g_err.c_str(); }

extern "C" int mmh_measure_fp64_tflops(int device, double* tflops)
{
    if (!tflops) return fail(MMH_EINVAL, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(MMH_ECUDA, "no usable CUDA device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, 8));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = prop.multiProcessorCount * 8, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(a));
        k_fp64_peak<<<grid, 256>>>(d, iters);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = std::max(best, (double)grid * 256.0 * 8.0 * 2.0 * iters / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    *tflops = best;
    return MMH_OK;
}
