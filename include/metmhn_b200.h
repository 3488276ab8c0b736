/* metmhn_b200 -- C-ABI of the B200-native metMHN likelihood/gradient engine.
 *
 * The reference (cbg-ethz/metMHN) has no FFI layer: its training hot path sits behind the
 * plain Python functions of metmhn/regularized_optimization.py.  This header is the boundary
 * a maintainer binds (ctypes stub in INTEGRATION.md) to run that path on a B200:
 *
 *   reference call                                          replaced by
 *   ------------------------------------------------------  ---------------------------
 *   score_and_grad(...)      regularized_optimization.py:163   mmh_value_grad
 *   score(...)               regularized_optimization.py:55    mmh_value
 *   per-row _lp_* dispatch   regularized_optimization.py:75-119  mmh_per_patient
 *   (dataset constant across L-BFGS iterations, :328)       mmh_create / mmh_destroy
 *
 * Plain pointers and sizes only.  All entry points return 0 on success or a negative
 * MMH_E* code; mmh_last_error() gives a thread-local message.  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with MMH_ECUDA.
 *
 * Layouts (identical to the reference):
 *   dat    : int8, n_dat rows of 2*n_mut+3 columns
 *            [PT_0, MT_0, ..., PT_{n-1}, MT_{n-1}, seeding, order, type]   (:63-66)
 *   params : double[(n+1)*(n+3)] = [log_theta row-major (n+1)^2, log_d_p (n+1), log_d_m (n+1)]
 *            (:292-294); gradients use the same packing (:296).
 */
#ifndef METMHN_B200_H
#define METMHN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMH_OK          0
#define MMH_EINVAL    (-1)   /* bad shape, non-binary genotype, paired row without seeding, n_mut too large */
#define MMH_ECUDA     (-2)   /* CUDA runtime failure or no device */
#define MMH_ENOMEM    (-3)
#define MMH_ETOOLARGE (-4)   /* a patient's restricted state space exceeds the supported size */
#define MMH_ENCCL     (-5)   /* NCCL missing (libnccl.so.2 could not be loaded) or a NCCL call failed */

#define MMH_MAX_MUT    28    /* events excluding seeding (LUAD uses 28) */
#define MMH_MAX_BITS   26    /* largest restricted lattice (bits) one space may have */

typedef struct mmh_handle mmh_handle;

typedef struct mmh_stats_t {
    int64_t n_dat;            /* rows given to mmh_create */
    int64_t n_em;             /* sum of the seeding column (regularized_optimization.py:256) */
    int64_t n_spaces;         /* lattices built (pre-seeding, joint, second-phase, unpaired) */
    int64_t n_chunks;         /* scratch-sized batches the evaluation streams through */
    int64_t n_launches;       /* kernel launches of the last evaluation */
    double  states_value_grad;/* sum over spaces of N_eff */
    double  alg_bytes;        /* 32 * sum N_eff : algorithmic bytes of one value+grad evaluation */
    double  alg_flops;        /* sum N_eff * (2 k_eff + 4 c n_tot) */
    double  exec_fma;         /* FP64 FMAs this implementation executes (estimate) */
    double  last_ms;          /* device time of the last evaluation (CUDA events) */
    double  scratch_bytes;    /* device scratch allocated */
    int64_t k_hist[4][64];    /* [type][k] histogram of restricted sizes */
    double  class_ms[8];      /* profile mode: device ms by kernel class of the last evaluation:
                                 0 table setup, 1 forward solves, 2 adjoint solves, 3 marginal statistics,
                                 4 gradient contraction (k_finish), 5 other, 6 weighted marginals of product-form spaces */
} mmh_stats_t;

/* Copy + preprocess the dataset (parse rows, canonical bit layout, bucket by lattice size,
 * plan scratch chunks, upload).  `device` is the CUDA ordinal.  `chunk_bytes` bounds the
 * scratch of one batch (0 = default). */
int mmh_create(mmh_handle** out, int n_mut, const int8_t* dat, int64_t n_dat, int64_t row_stride,
               int device, int64_t chunk_bytes);

/* score_and_grad: *score and grad[(n+1)(n+3)] as the reference returns them (means, weighted). */
int mmh_value_grad(mmh_handle* h, const double* params, double perc_met, double* score, double* grad);

/* score (value only; skips the adjoint pass). */
int mmh_value(mmh_handle* h, const double* params, double perc_met, double* score);

/* Same computation with caller-supplied class weights instead of the handle's own
 * n_em / n_nm (used when the dataset is sharded over several handles / GPUs: every shard is
 * scaled with the GLOBAL weights so that the shard results simply add up).
 *   result = w_type0 * sum_{type 0} + w_other * sum_{types 1,2,3}
 * out_host (1 + (n+1)(n+3) doubles: score, grad) and/or out_dev (same layout, device
 * pointer on the handle's device, e.g. the buffer a NCCL all-reduce runs on) may be NULL.
 * want_grad = 0 writes only out[0]. */
int mmh_eval_weighted(mmh_handle* h, const double* params, double w_type0, double w_other,
                      int want_grad, double* out_host, double* out_dev);

/* Asynchronous variant for device-resident callers: d_params and d_out are DEVICE pointers on the
 * handle's device (d_out: 1 + (n+1)(n+3) doubles).  Work is queued on the handle's stream; call
 * mmh_sync() before reading d_out from another stream or the host. */
int mmh_eval_device(mmh_handle* h, const double* d_params, double w_type0, double w_other,
                    int want_grad, double* d_out);
int mmh_sync(mmh_handle* h);

/* Profile mode: time every kernel class of an evaluation with CUDA events on the launching stream
 * (mmh_stats_t.class_ms).  Adds a synchronisation per evaluation; off by default. */
int mmh_set_profile(mmh_handle* h, int on);

/* Per-row log-likelihoods of the last evaluation's parameters (test hook).  Rows with an
 * unknown type get 0. */
int mmh_per_patient(mmh_handle* h, const double* params, double* logp);
/* The same with gradients (the per-patient `_g_coupled_{0,1,2}`, `_grad_prim_obs`, `_grad_met_obs` of
 * metmhn/jx/likelihood.py:442-730): rows first_row .. first_row + n_rows - 1, each evaluated as its own one-row dataset
 * with unit weights; logp[n_rows], grads[n_rows * (n+1)(n+3)] or NULL.  A test hook (milliseconds per row). */
int mmh_per_patient_grads(mmh_handle* h, const double* params, int64_t first_row, int64_t n_rows, double* logp, double* grads);

/* ---- multi-GPU: one handle per GPU, the shard results are summed by ONE in-library ncclAllReduce ----------
 * The reference sums per-patient results (regularized_optimization.py:187-267); with the dataset sharded over
 * handles that are evaluated with the GLOBAL class weights (mmh_eval_weighted / mmh_eval_device) the only
 * exchange is the sum of 1 + (n+1)(n+3) doubles.  libnccl.so.2 is loaded with dlopen() on first use, so the
 * library itself has no link-time NCCL dependency.
 *
 * One process per GPU (torchrun / MPI): rank 0 calls mmh_nccl_unique_id() and ships the 128 bytes to the other
 * ranks by whatever means it has; every rank then calls mmh_comm_init() (collective, blocks until all ranks
 * joined).  From then on every mmh_eval_weighted / mmh_eval_device / mmh_value_grad / mmh_value on that handle
 * ends with ncclAllReduce(sum) of the result ON THE HANDLE'S STREAM, i.e. ordered with the evaluation that
 * produced it and with the next one that overwrites it; all ranks must make the same sequence of calls. */
#define MMH_NCCL_ID_BYTES 128
int mmh_nccl_unique_id(char id[MMH_NCCL_ID_BYTES]);
int mmh_comm_init(mmh_handle* h, const char id[MMH_NCCL_ID_BYTES], int nranks, int rank);
int mmh_comm_destroy(mmh_handle* h);

/* One process, several GPUs: the rows are partitioned over `n_devices` devices by a longest-processing-time
 * cost model (same as metmhn_b200/sharded.py), one handle per device, ncclCommInitAll, and every evaluation
 * runs on all devices concurrently and ends with the all-reduce.  Same result layout as mmh_value_grad. */
typedef struct mmh_multi mmh_multi;
int mmh_multi_create(mmh_multi** out, int n_mut, const int8_t* dat, int64_t n_dat, int64_t row_stride,
                     const int* device_ids, int n_devices, int64_t chunk_bytes);
int mmh_multi_value_grad(mmh_multi* m, const double* params, double perc_met, double* score, double* grad);
int mmh_multi_value(mmh_multi* m, const double* params, double perc_met, double* score);
void mmh_multi_destroy(mmh_multi* m);
/* The partition's work estimate of one row (lattice states times the measured cost per state of the row's kind and size,
 * picoseconds on a B200); pure host arithmetic, no device needed.  metmhn_b200/sharded.py: patient_cost is the same table. */
double mmh_row_cost(const int8_t* row, int n_mut);

/* learn_mhn (regularized_optimization.py:301-334) as ONE call: minimises -score + w_penal * symmetric_penal(params)
 * (:46-52, smoothing eps, 1e-5 in the reference) from x0 with L-BFGS-B without bounds (m = 10, More'-Thuente line search,
 * stopping tests of SciPy's driver: relative decrease <= ftol, max|gradient| <= 1e-5, max_iter iterations).  The
 * likelihood, its gradient AND the penalty are evaluated on the device; the host keeps the limited-memory vectors.  x_out
 * receives (n+1)(n+3) doubles; f_out, n_iter, n_eval may be NULL.  With a communicator attached (mmh_comm_init) every
 * rank runs the same loop on the all-reduced objective. */
int mmh_learn(mmh_handle* h, const double* x0, double perc_met, double w_penal, double eps, int64_t max_iter, double ftol,
              double* x_out, double* f_out, int64_t* n_iter, int64_t* n_eval);

/* GPU Gillespie sampler of the metMHN process (metmhn/simulations.py:8-147 `single_traject` / `simulate_dat`): n_sim
 * trajectories, one thread each, Philox4x32-10 random numbers keyed by (seed, trajectory index).  geno receives n_sim rows
 * of 2*n_mut+1 int8 [PT_0, MT_0, ..., PT_{n-1}, MT_{n-1}, seeding], order one int8 per row: 1 = PT diagnosed first,
 * 2 = MT diagnosed first, 0 = the trajectory ended before seeding (never-metastasised primary tumour). */
int mmh_simulate(int n_mut, const double* params, int64_t n_sim, uint64_t seed, int device, int8_t* geno, int8_t* order);

int mmh_stats(mmh_handle* h, mmh_stats_t* out);
void mmh_destroy(mmh_handle* h);
const char* mmh_last_error(void);
/* Measured FP64 FMA throughput of the device in TFLOP/s (independent-DFMA micro-kernel);
 * denominator for the FP64 roofline since MEASURED_PEAKS.json has none. */
int mmh_measure_fp64_tflops(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
