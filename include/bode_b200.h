/* bode_b200 -- C ABI of the B200-native EKLD acquisition hot path.
 *
 * Drop-in boundary for the reference path (piyushpandita92/bode)
 *     KLSampler.mcmc_kld            bode/_sampler.py:483-492
 *       -> make_mcmc_model          bode/_sampler.py:449-462   (GPy exact-GP inference per hyper-sample)
 *       -> avg_kld_mean             bode/_sampler.py:464-481   (closed-form Gaussian EKLD)
 *       -> get_sigma_1/get_params_2 bode/_sampler.py:404-410, 391-402
 *       -> xik / bk                 bode/_core.py:86-101
 *     KLSampler.get_mu_sigma        bode/_sampler.py:434-444   (Q-posterior summary, by-product of prepare)
 * The reference has no FFI (pure Python on GPy); these entry points are what a ctypes binding inside
 * `KLSampler.mcmc_kld` / `get_mu_sigma` would call.
 *
 * Conventions
 *   - plain C, no torch types.  All arrays are FP64, row-major (NumPy C order).
 *   - `*_dev` entry points take DEVICE pointers and a `cudaStream_t` passed as `void*`; they never allocate,
 *     never synchronise and never touch the host.  `*_host` entry points take HOST pointers, allocate device
 *     scratch internally and synchronise before returning.
 *   - hyper-sample layout is the reference's (`_sampler.py:273-274, 455-461`): `hyp[S][ndim]`, row =
 *     [variance, ell_1..ell_d] (ndim == d+1, noise variance = nugget_var) or
 *     [variance, ell_1..ell_d, noise_std] (ndim == d+2, noise variance = noise_std^2).
 *   - every function returns 0 on success or a negative `bode_status`; nothing throws across the ABI.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns BODE_ERR_CUDA.
 */
#ifndef BODE_B200_H_
#define BODE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bode_status {
  BODE_OK = 0,
  BODE_ERR_BAD_ARG = -1,     /* null pointer, non-positive size, ndim not in {d+1, d+2} */
  BODE_ERR_UNSUPPORTED = -2, /* n > BODE_MAX_N or d > BODE_MAX_D */
  BODE_ERR_WORKSPACE = -3,   /* workspace too small / misaligned */
  BODE_ERR_CUDA = -4,        /* CUDA runtime error (see bode_last_cuda_error) */
  BODE_ERR_NOT_PD = -5,      /* host entry points only: a Cholesky pivot was non-positive (info[s] > 0) */
  BODE_ERR_NCCL = -6         /* libnccl.so.2 not loadable, an NCCL call failed (see bode_last_cuda_error), or a peer did not
                                deliver its record to bode_ekld_argmax_p2p */
} bode_status;

#define BODE_MAX_N 1024 /* observations */
#define BODE_MAX_D 16   /* input dimensions */

/* EKLD expression: the reference's literal (uncancelled) expression `_sampler.py:480`, or the analytically
 * identical -1/2 log1p(-v^2/(sigma_1 k_xx)) which does not lose digits for small EKLD. */
#define BODE_FORM_REFERENCE 0
#define BODE_FORM_LOG1P 1

const char* bode_version(void);
const char* bode_strerror(int status);
const char* bode_last_cuda_error(void);

/* The developer tuning knobs (environment variables BODE_SCORE_VARIANT, BODE_CHUNK_KB, BODE_WAVES, BODE_STAGES,
 * BODE_XIK_SERIES, ...) are read once when the library is loaded, never on the call path; this re-reads them. */
int bode_tuning_reload(void);

/* Bytes of device workspace `bode_ekld_prepare_dev` needs for (n observations, d dims, S hyper-samples).
 * Returns 0 if the shape is unsupported. 256-byte alignment required. */
size_t bode_ekld_workspace_bytes(int n, int d, int S);

/* K1 -- per-hyper-sample preparation; replaces S calls of make_mcmc_model (`_sampler.py:449-462`) plus the
 * candidate-independent parts of get_sigma_1 / get_mu_1 (`:404-422`):
 *   A_s = K_ARD-RBF(X, X; theta_s) + (noise_var_s + jitter) I,  L_s = chol(A_s),  M_s = L_s^-1,
 *   e_s = xik(X),  sigma0_s = bk,  a = L^-1 e,  w_s = L^-T a,  sigma1_s = sigma0 - a.a,  mu1_s = (L^-1 Y).a,
 *   loglik_s = 0.5(-n log 2pi - logdet A_s - Y^T A_s^-1 Y)      (GPy log_marginal, `_sampler.py:232`),
 * plus the Chebyshev series of the d factors of xik(x; theta_s) (`_core.py:86-92`) that the scoring call evaluates.
 * X[n*d], Y[n], hyp[S*ndim], info[S] are device pointers.  info[s] = 0, or j+1 if pivot j of sample s was not
 * positive (LAPACK dpotrf convention), or -1 if theta_s is outside the supported range (variance in (9.5e-7, 1e6),
 * 0 <= noise_var < 1e6, finite positive lengthscales with sum_k 1/ell_k^2 < 4.3e7, i.e. no shorter than ~2e-4 of the
 * unit box); such a sample yields NaN downstream.
 * `jitter` is GPy's hidden diagonal addition (1e-8 reproduces the reference fixture). */
int bode_ekld_prepare_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                          double nugget_var, double jitter, void* workspace, size_t workspace_bytes, int* info,
                          void* stream);

/* Same with a per-sample diagonal addition jitter_per_sample[S] (device) instead of the scalar: what a caller uses to
 * mirror GPy's jitchol, which on a failed factorisation retries that model with mean(diag) * 1e-6 * 10^k, k = 0..4
 * (quadrature mode is not covered by this variant). */
int bode_ekld_prepare_jitter_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                                 double nugget_var, const double* jitter_per_sample, void* workspace, size_t workspace_bytes,
                                 int* info, void* stream);

/* Per-sample by-products of prepare, device -> device: out[s*8 + {0..7}] =
 * {variance, noise_var, sigma0, sigma1, mu1, logdet, Y^T A^-1 Y, loglik}. */
int bode_ekld_sample_terms_dev(const void* workspace, int n, int d, int S, double* out, void* stream);

/* Q-posterior summary (`get_mu_sigma`, `_sampler.py:434-444`): out[0] = mean_s mu1, out[1] = mean_s sigma1,
 * accumulated sequentially s = 0..S-1 like the reference loop.  Device pointer. */
int bode_ekld_mu_sigma_dev(const void* workspace, int n, int d, int S, double* out, void* stream);

/* K2+K3+K4 -- dense scoring; replaces Nd calls of mcmc_kld (`_sampler.py:483-492`):
 *   score[c] = (1/S) sum_s log EKLD_s(X_design[c])  (sequential in s),
 *   var[c] = (1/S) sum_s (log EKLD_s - score[c])^2  (nullable),
 *   best[0] = max score (NaN counts as largest, like np.argmax), best_idx[0] = idx_offset + first index attaining it.
 *   logk[s*Nd + c] = log EKLD_s(X_design[c]) -- NULLABLE: pass S*Nd doubles only if the per-sample matrix is wanted (the
 *   `kld_j` of `:489-491`).
 * Two launches: k_score (cross-covariance and L^-1 contraction on the FP64 tensor pipe; writes sigma_x and w.k per
 * (sample, candidate)) and k_ekld_reduce (xik(x), v, the closed-form EKLD and its log, mean / var over samples in sample
 * order, block argmax, final argmax by the last block to finish).
 * `idx_offset` is the global index of X_design[0] (multi-GPU candidate sharding).
 * `scratch` (256-byte aligned) holds the argmax partials, the w.k plane and -- when logk is NULL -- the sigma_x plane:
 * bode_ekld_score_scratch_bytes(Nd, S) bytes when logk is given, bode_ekld_score_scratch_bytes(Nd, 2 * S) when it is NULL.
 * `best` and `best_idx` may point at the two halves of one 16-byte record (what bode_ekld_argmax_nccl exchanges). */
size_t bode_ekld_score_scratch_bytes(int64_t Nd, int S);
int bode_ekld_score_dev(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                        int64_t idx_offset, int form, double* logk, double* score, double* var, double* best,
                        int64_t* best_idx, void* scratch, void* stream);

/* The other EKLD form from the planes of the previous call: re-runs only k_ekld_reduce (xik, v, EKLD, log, mean / var,
 * argmax) on the sigma_x and w.k planes that the last bode_ekld_score_dev call with the same workspace, X_design, Nd,
 * scratch and with logk == NULL and var == NULL left in `scratch` (a call with logk or var turns the sigma_x plane into
 * log EKLD and cannot be followed by this).  What `KLSampler`'s check of the `log1p` form against the literal expression
 * of `_sampler.py:480` costs: 3 % of a scoring call instead of a second one.  `var` is nullable. */
int bode_ekld_rescore_dev(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                          int64_t idx_offset, int form, double* score, double* var, double* best, int64_t* best_idx,
                          void* scratch, void* stream);

/* Same as bode_ekld_score_dev, and additionally times the dominant kernel (the fused K2-K4 scoring kernel) with
 * CUDA events recorded on `stream`; synchronises on the stop event before returning.  Used by bench.py for the
 * roofline figure. */
int bode_ekld_score_dev_timed(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                              int64_t idx_offset, int form, double* logk, double* score, double* var, double* best,
                              int64_t* best_idx, void* scratch, void* stream, float* score_kernel_ms);

/* Same as bode_ekld_score_dev, and records the caller's CUDA events (cudaEvent_t, created with timing enabled) on
 * `stream` immediately before and after the fused K2-K3 scoring kernel, without synchronising: the caller reads
 * cudaEventElapsedTime once its own timed region is over (bench.py's roofline figure). */
int bode_ekld_score_dev_events(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                               int64_t idx_offset, int form, double* logk, double* score, double* var, double* best,
                               int64_t* best_idx, void* scratch, void* stream, void* ev_start, void* ev_stop);

/* ---- quadrature mode (north star: k(X_design, X_quad) / k(X, X_quad) averaged over quadrature points) ------------------
 * The reference integrates the kernel over [0,1]^d in closed form (xik / bk, `_core.py:86-101`); quadrature mode
 * replaces the three integrals by (weighted) means over user-supplied quadrature points X_quad[Nq*d] (the drivers'
 * `quad_points`, `samp_ex1.py:61-62`; `wq` nullable = uniform weights):
 *     e_q[j] = mean_q k(X_j, x_q),   xi_q(x) = mean_q k(x, x_q),   sigma0_q = mean_{q,q'} k(x_q, x_q').
 * It differs from the reference by quadrature error (SURVEY note Q) and is parity-checked against the oracle's own
 * quadrature variant on the same X_quad. */
size_t bode_quad_scratch_bytes(int n, int S, int Nq);
/* out[s*Np + i] = (variance_s / sum w) * sum_q w_q exp(-|| (P_i - x_q) / (sqrt2 ell_s) ||^2); scratch >= 256 bytes. */
int bode_quad_rowmean_dev(const double* hyp, int S, int ndim, int d, const double* P, int64_t Np, const double* Xq,
                          const double* wq, int Nq, double* out, void* scratch, void* stream);
/* bode_ekld_prepare_dev with e_q and sigma0_q in place of xik(X) and bk; scratch >= bode_quad_scratch_bytes(n,S,Nq). */
int bode_ekld_prepare_quad_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                               double nugget_var, double jitter, const double* Xq, const double* wq, int Nq,
                               void* workspace, size_t workspace_bytes, int* info, void* scratch, void* stream);
/* The same with the S*Nq row means  rowmean_q[s*Nq + q] = mean_q' k_s(x_q, x_q')  supplied by the caller: the O(Nq^2 S) part of
 * the preparation, which a multi-GPU caller computes in row shards with bode_quad_rowmean_dev (P = its rows of X_quad) and
 * all-gathers -- every rank then forms sigma0_q from identical numbers in the same order. */
int bode_ekld_prepare_quad_rm_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                                  double nugget_var, double jitter, const double* Xq, const double* wq, int Nq,
                                  const double* rowmean_q, void* workspace, size_t workspace_bytes, int* info, void* scratch,
                                  void* stream);
/* bode_ekld_score_dev with xi_q(x) in place of xik(x); xi is an S*Nd buffer (returned filled); logk nullable as above;
 * scratch >= 256 + bode_ekld_score_scratch_bytes(Nd, logk ? S : 2 * S). */
int bode_ekld_score_quad_dev(const void* workspace, int n, int d, int S, const double* hyp, int ndim,
                             const double* X_design, int64_t Nd, int64_t idx_offset, int form, const double* Xq,
                             const double* wq, int Nq, double* xi, double* logk, double* score, double* var, double* best,
                             int64_t* best_idx, void* scratch, void* stream);

/* Uncertainty-sampling comparator (`update_comp_models`, `_sampler.py:506-514`): pred_var[c] = mean_s latent
 * predictive variance at X_grid[c]; best/best_idx as above. Reuses the scoring kernel (no v / EKLD epilogue). */
int bode_us_variance_dev(const void* workspace, int n, int d, int S, const double* X_grid, int64_t Ng,
                         int64_t idx_offset, double* sigx /* S*Ng scratch */, double* pred_var, double* best,
                         int64_t* best_idx, void* scratch, void* stream);

/* Host-buffer entry point: the call a non-torch caller (or the reference's Python via ctypes) makes.
 * Copies X, Y, hyp, X_design to the device, runs prepare + score, copies results back, synchronises.
 * score[Nd] required; var[Nd], kld[S*Nd] (EKLD values, not logs), mu_sigma[2], info[S] nullable.
 * Device buffers live in a reusable context and only grow: after the first call of a given size there is no
 * cudaMalloc / cudaFree on the call path (the sequential design loop calls this once per iteration).
 * `bode_ekld_score_host` uses a process-wide context; `*_ctx` variants take an explicit one (NULL = the process-wide one).
 * A context is bound to the CUDA device current at its first use and serialises its callers. */
typedef struct bode_ctx bode_ctx;
int bode_ctx_create(bode_ctx** ctx);
int bode_ctx_destroy(bode_ctx* ctx);
long long bode_ctx_alloc_count(bode_ctx* ctx); /* cudaMalloc calls made by this context so far */
int bode_ekld_score_host(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                         double nugget_var, double jitter, const double* X_design, int64_t Nd, int form,
                         double* score, double* var, double* kld, int64_t* best_idx, double* mu_sigma, int* info);
int bode_ekld_score_host_ctx(bode_ctx* ctx, const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                             double nugget_var, double jitter, const double* X_design, int64_t Nd, int form,
                             double* score, double* var, double* kld, int64_t* best_idx, double* mu_sigma, int* info);

/* Batched GP log marginal likelihood for the host MCMC (`lnprob` -> `get_likelihood`, `_sampler.py:221-238`):
 * loglik[s] for each row of hyp; -inf where the Cholesky fails.  Host pointers. */
int bode_gp_loglik_host(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                        double nugget_var, double jitter, double* loglik);
int bode_gp_loglik_host_ctx(bode_ctx* ctx, const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                            double nugget_var, double jitter, double* loglik);

/* Device-pointer form of the same (one launch of the prepare kernel in likelihood-only mode + a gather):
 * loglik[S] device; workspace as for bode_ekld_prepare_dev (its scoring state is NOT valid afterwards). */
int bode_gp_loglik_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                       double nugget_var, double jitter, void* workspace, size_t workspace_bytes, int* info, double* loglik,
                       void* stream);

/* ---- multi-GPU: candidate blocks are independent, the only exchange is the argmax (replaces the
 * np.argmax of `_sampler.py:659` across ranks).  Each rank's scoring call leaves a 16-byte record {double best;
 * int64 best_idx} (pass best = rec, best_idx = rec + 8); bode_ekld_argmax_nccl enqueues ONE ncclAllGather of that record
 * on `stream` (no host synchronisation) into gathered[16 * world] (device); the caller copies those bytes to the host
 * once and bode_argmax_combine picks the winner with np.argmax semantics (NaN counts as the maximum, ties -> lowest
 * index; records with idx < 0 are ranks without candidates).  NCCL is resolved at run time (dlopen of libnccl.so.2 --
 * inside a torch process that is torch's own copy); the communicator is created from a 128-byte ncclUniqueId that rank 0
 * generates and the caller distributes (torch.distributed broadcast, MPI, a file, ...). */
int bode_nccl_unique_id(void* id128);
int bode_nccl_comm_create(const void* id128, int world, int rank, void** comm);
int bode_nccl_comm_destroy(void* comm);
int bode_ekld_argmax_nccl(void* comm, const void* best_rec, void* gathered, void* stream);
int bode_argmax_combine(const void* records_host, int count, double* best, int64_t* best_idx);

/* The same exchange WITHOUT NCCL on the data path: one kernel over NVLink peer memory.  Every rank owns a mailbox of
 * 2 x world 32-byte slots {best, best_idx, epoch}, exported as a CUDA IPC handle and mapped by every other rank (one
 * process per GPU, same node).  bode_ekld_argmax_p2p enqueues one warp on `stream`: lane r stores this rank's record
 * into its slot of rank r's mailbox (system-scope release of the epoch word) and then waits for rank r's record in its
 * own mailbox; gathered[16 * world] (device) is filled exactly like the NCCL form.  The epoch counter lives in device
 * memory, so the launch replays from a CUDA graph; slots alternate with the epoch's parity.  A peer that does not
 * deliver within 5 s leaves {NaN, -2} in its slot and bode_argmax_combine returns BODE_ERR_NCCL.  Collective: every rank
 * must enqueue the same number of calls.
 *   bode_xchg_create(world, rank, &x, handle)   allocates the mailbox, returns its 64-byte IPC handle
 *   bode_xchg_connect(x, handles)               handles = the world handles in rank order (distributed by the caller);
 *                                               synchronise the ranks (a barrier) between connect and the first exchange */
typedef struct bode_xchg bode_xchg;
int bode_xchg_create(int world, int rank, bode_xchg** x, unsigned char handle[64]);
int bode_xchg_connect(bode_xchg* x, const unsigned char* handles);
int bode_xchg_destroy(bode_xchg* x);
int bode_ekld_argmax_p2p(bode_xchg* x, const void* best_rec, void* gathered, void* stream);

/* Host-only: JSON description of the launch plan the scoring entry points would use for this shape on a device with
 * `sm_count` SMs and `smem_optin` bytes of opt-in shared memory per CTA (kernel variant, candidates per CTA, ring
 * chunks, row-tile assignment, work items).  No CUDA call is made. */
int bode_ekld_plan_describe(int n, int d, int S, int64_t Nd, int sm_count, size_t smem_optin, char* out, size_t cap);

/* Measured FP64 peak of this GPU (DMMA.8x8x4 register-resident loop and DFMA loop), TFLOP/s; used by bench.py
 * as the roofline denominator because MEASURED_PEAKS.json has no FP64 entry. */
int bode_fp64_peak(double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* BODE_B200_H_ */
