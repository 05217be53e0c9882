// C ABI of bode_b200 (see include/bode_b200.h).  Thin, allocation-free glue around the kernels for the *_dev
// entry points; the *_host entry points own their device scratch for the duration of the call.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/bode_b200.h"
#include "layout.h"
#include "prepare_args.h"
#include "score_args.h"

namespace bode {
cudaError_t launch_prepare(const PrepArgs& a, cudaStream_t st);
cudaError_t launch_sample_terms(const void* ws, const Layout& L, double* out, cudaStream_t st);
cudaError_t launch_mu_sigma(const void* ws, const Layout& L, double* out, cudaStream_t st);
cudaError_t launch_loglik_gather(const void* ws, const Layout& L, double* out, cudaStream_t st);
int plan_score(const Layout& L, int64_t Nd, int sm_count, size_t smem_optin, ScorePlan* P);
cudaError_t launch_score(const ScoreArgs& a, const ScorePlan& P, cudaStream_t st);
cudaError_t launch_ekld(const void* ws, const Layout& L, const double* Xd, const double* xi_ovr, double* logk,
                        const double* wkplane, int S, int64_t Nd, int form, cudaStream_t st);
cudaError_t launch_reduce(const double* logk, int S, int64_t Nd, int64_t idx_offset, double* score, double* var,
                          double* part_val, int64_t* part_idx, double* best, int64_t* best_idx, cudaStream_t st);
cudaError_t launch_exp_inplace(double* v, size_t n, cudaStream_t st);
cudaError_t measure_fp64_peak(double* dmma, double* dfma);
cudaError_t launch_rowmean(const double* hyp, int S, int ndim, int d, const double* P, int64_t Np, const double* Xq,
                           const double* wq, int Nq, const double* wsum_ptr, double* out, cudaStream_t st);
cudaError_t launch_weighted_mean(const double* rm, const double* wq, int Nq, const double* wsum_ptr, int S, double* out,
                                 cudaStream_t st);
cudaError_t launch_sum_weights(const double* wq, int Nq, double* out, cudaStream_t st);
}  // namespace bode

using namespace bode;

static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
  return BODE_ERR_CUDA;
}
#define BODE_CUDA(call)                                  \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

struct DevProps {
  int sm_count;
  size_t smem_optin;
};
static int get_props(DevProps* p) {
  static std::mutex mu;
  static std::vector<DevProps> cache;
  int dev = 0;
  BODE_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if ((int)cache.size() <= dev) cache.resize(dev + 1, DevProps{0, 0});
  if (cache[dev].sm_count == 0) {
    int sm = 0, optin = 0;
    BODE_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    BODE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev].sm_count = sm;
    cache[dev].smem_optin = (size_t)optin;
  }
  *p = cache[dev];
  return BODE_OK;
}

static bool shape_ok(int n, int d, int S) { return n >= 1 && d >= 1 && S >= 1; }
static bool shape_supported(int n, int d) { return n <= BODE_MAX_N && d <= BODE_MAX_D; }

// ---- host-buffer entry points --------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T> T* as() { return static_cast<T*>(p); }
};

extern "C" {

const char* bode_version(void) { return "bode_b200 0.1.0 (sm_100a)"; }

const char* bode_strerror(int status) {
  switch (status) {
    case BODE_OK: return "ok";
    case BODE_ERR_BAD_ARG: return "bad argument";
    case BODE_ERR_UNSUPPORTED: return "unsupported shape (n > 1024 or d > 16)";
    case BODE_ERR_WORKSPACE: return "workspace too small or misaligned";
    case BODE_ERR_CUDA: return "CUDA error";
    case BODE_ERR_NOT_PD: return "covariance matrix not positive definite";
    default: return "unknown status";
  }
}

const char* bode_last_cuda_error(void) { return g_cuda_err; }

size_t bode_ekld_workspace_bytes(int n, int d, int S) {
  if (!shape_ok(n, d, S) || !shape_supported(n, d)) return 0;
  return make_layout(n, d, S).total_bytes;
}

int bode_ekld_prepare_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                          double nugget_var, double jitter, void* workspace, size_t workspace_bytes, int* info,
                          void* stream) {
  if (!X || !Y || !hyp || !workspace || !info || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  PrepArgs a;
  a.X = X; a.Y = Y; a.hyp = hyp; a.n = n; a.d = d; a.S = S; a.ndim = ndim;
  a.nugget_var = nugget_var; a.jitter = jitter;
  a.ws = static_cast<char*>(workspace);
  a.L = make_layout(n, d, S);
  a.info = info;
  a.e_ovr = nullptr;
  a.sigma0_ovr = nullptr;
  a.loglik_only = 0;
  if (workspace_bytes < a.L.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return BODE_ERR_WORKSPACE;
  BODE_CUDA(launch_prepare(a, static_cast<cudaStream_t>(stream)));
  return BODE_OK;
}

size_t bode_quad_scratch_bytes(int n, int S, int Nq) {
  if (n < 1 || S < 1 || Nq < 1) return 0;
  // [wsum (256 B)] [e: S*n] [rowmean of X_quad: S*Nq] [sigma0: S]
  return 256 + align_up(sizeof(double) * (size_t)S * n, 256) + align_up(sizeof(double) * (size_t)S * Nq, 256) +
         align_up(sizeof(double) * (size_t)S, 256);
}

int bode_quad_rowmean_dev(const double* hyp, int S, int ndim, int d, const double* P, int64_t Np, const double* Xq,
                          const double* wq, int Nq, double* out, void* scratch, void* stream) {
  if (!hyp || !P || !Xq || !out || !scratch || S < 1 || d < 1 || Np < 1 || Nq < 1) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (d > BODE_MAX_D) return BODE_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* wsum = static_cast<double*>(scratch);
  if (wq) BODE_CUDA(launch_sum_weights(wq, Nq, wsum, st));
  BODE_CUDA(launch_rowmean(hyp, S, ndim, d, P, Np, Xq, wq, Nq, wq ? wsum : nullptr, out, st));
  return BODE_OK;
}

int bode_ekld_prepare_quad_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                               double nugget_var, double jitter, const double* Xq, const double* wq, int Nq,
                               void* workspace, size_t workspace_bytes, int* info, void* scratch, void* stream) {
  if (!X || !Y || !hyp || !workspace || !info || !Xq || !scratch || !shape_ok(n, d, S) || Nq < 1) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* sc = static_cast<char*>(scratch);
  double* wsum = reinterpret_cast<double*>(sc);
  double* e = reinterpret_cast<double*>(sc + 256);
  double* rmq = reinterpret_cast<double*>(sc + 256 + align_up(sizeof(double) * (size_t)S * n, 256));
  double* s0 = reinterpret_cast<double*>(sc + 256 + align_up(sizeof(double) * (size_t)S * n, 256) +
                                         align_up(sizeof(double) * (size_t)S * Nq, 256));
  if (wq) BODE_CUDA(launch_sum_weights(wq, Nq, wsum, st));
  const double* wsp = wq ? wsum : nullptr;
  BODE_CUDA(launch_rowmean(hyp, S, ndim, d, X, n, Xq, wq, Nq, wsp, e, st));      // e_q = mean_q k(X, x_q)
  BODE_CUDA(launch_rowmean(hyp, S, ndim, d, Xq, Nq, Xq, wq, Nq, wsp, rmq, st));  // mean_q' k(x_q, x_q')
  BODE_CUDA(launch_weighted_mean(rmq, wq, Nq, wsp, S, s0, st));                  // sigma0_q
  PrepArgs a;
  a.X = X; a.Y = Y; a.hyp = hyp; a.n = n; a.d = d; a.S = S; a.ndim = ndim;
  a.nugget_var = nugget_var; a.jitter = jitter;
  a.ws = static_cast<char*>(workspace);
  a.L = make_layout(n, d, S);
  a.info = info;
  a.e_ovr = e;
  a.sigma0_ovr = s0;
  a.loglik_only = 0;
  if (workspace_bytes < a.L.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return BODE_ERR_WORKSPACE;
  BODE_CUDA(launch_prepare(a, st));
  return BODE_OK;
}

int bode_gp_loglik_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                       double nugget_var, double jitter, void* workspace, size_t workspace_bytes, int* info, double* loglik,
                       void* stream) {
  if (!X || !Y || !hyp || !workspace || !info || !loglik || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  PrepArgs a;
  a.X = X; a.Y = Y; a.hyp = hyp; a.n = n; a.d = d; a.S = S; a.ndim = ndim;
  a.nugget_var = nugget_var; a.jitter = jitter;
  a.ws = static_cast<char*>(workspace);
  a.L = make_layout(n, d, S);
  a.info = info;
  a.e_ovr = nullptr;
  a.sigma0_ovr = nullptr;
  a.loglik_only = 1;
  if (workspace_bytes < a.L.total_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return BODE_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BODE_CUDA(launch_prepare(a, st));
  BODE_CUDA(launch_loglik_gather(workspace, a.L, loglik, st));
  return BODE_OK;
}

int bode_ekld_sample_terms_dev(const void* workspace, int n, int d, int S, double* out, void* stream) {
  if (!workspace || !out || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  BODE_CUDA(launch_sample_terms(workspace, make_layout(n, d, S), out, static_cast<cudaStream_t>(stream)));
  return BODE_OK;
}

int bode_ekld_mu_sigma_dev(const void* workspace, int n, int d, int S, double* out, void* stream) {
  if (!workspace || !out || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  BODE_CUDA(launch_mu_sigma(workspace, make_layout(n, d, S), out, static_cast<cudaStream_t>(stream)));
  return BODE_OK;
}

static size_t score_partials_bytes(int64_t Nd) {
  const size_t nb = (size_t)((Nd + 255) / 256);
  return align_up(nb * sizeof(double), 256) + align_up(nb * sizeof(int64_t), 256);
}

size_t bode_ekld_score_scratch_bytes(int64_t Nd, int S) {
  if (Nd < 1 || S < 1) return 0;
  return score_partials_bytes(Nd) + align_up(sizeof(double) * (size_t)S * (size_t)Nd, 256);  // argmax partials + v plane
}

static int score_common(const void* workspace, int n, int d, int S, const double* Xd, int64_t Nd, int64_t idx_offset,
                        int form, int mode, const double* xi_ovr, double* logk, double* score, double* var, double* best,
                        int64_t* best_idx, void* scratch, void* stream, float* kernel_ms, void* ev_start = nullptr,
                        void* ev_stop = nullptr) {
  if (!workspace || !Xd || !logk || !score || !best || !best_idx || !scratch || !shape_ok(n, d, S) || Nd < 1)
    return BODE_ERR_BAD_ARG;
  if (form != BODE_FORM_REFERENCE && form != BODE_FORM_LOG1P) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  DevProps props;
  int rc = get_props(&props);
  if (rc != BODE_OK) return rc;
  ScorePlan P;
  const Layout L = make_layout(n, d, S);
  if (plan_score(L, Nd, props.sm_count, props.smem_optin, &P) != 0) return BODE_ERR_UNSUPPORTED;
  ScoreArgs a;
  a.ws = workspace; a.L = L; a.Xd = Xd; a.Nd = Nd; a.logk = logk; a.S = S;
  a.vplane = reinterpret_cast<double*>(static_cast<char*>(scratch) + score_partials_bytes(Nd));
  a.sc = P.sc; a.ntiles = P.ntiles; a.form = form; a.mode = mode;
  {
    const char* dbg = getenv("BODE_DEBUG_SKIP");
    a.debug_skip = dbg ? atoi(dbg) : 0;
    const char* sg = getenv("BODE_STAGGER");
    a.stagger = sg ? atoi(sg) : 0;
  }
  a.stages = P.stages; a.chunk_max = P.chunk_max; a.nchunks = P.nchunks; a.u_global = P.u_global;
  memcpy(a.chunk_g0, P.chunk_g0, sizeof(a.chunk_g0));
  memcpy(a.tile_id, P.tile_id, sizeof(a.tile_id));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (kernel_ms) {
    BODE_CUDA(cudaEventCreate(&ev0));
    BODE_CUDA(cudaEventCreate(&ev1));
    BODE_CUDA(cudaEventRecord(ev0, st));
  }
  if (ev_start) BODE_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_start), st));
  BODE_CUDA(launch_score(a, P, st));
  if (ev_stop) BODE_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_stop), st));
  if (kernel_ms) BODE_CUDA(cudaEventRecord(ev1, st));
  if (mode == kModeEkld) BODE_CUDA(launch_ekld(workspace, L, Xd, xi_ovr, logk, a.vplane, S, Nd, form, st));
  const size_t nb = (size_t)((Nd + 255) / 256);
  double* part_val = static_cast<double*>(scratch);
  int64_t* part_idx = reinterpret_cast<int64_t*>(static_cast<char*>(scratch) + align_up(nb * sizeof(double), 256));
  BODE_CUDA(launch_reduce(logk, S, Nd, idx_offset, score, var, part_val, part_idx, best, best_idx, st));
  if (kernel_ms) {
    BODE_CUDA(cudaEventSynchronize(ev1));
    BODE_CUDA(cudaEventElapsedTime(kernel_ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
  }
  return BODE_OK;
}

int bode_ekld_score_dev(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                        int64_t idx_offset, int form, double* logk, double* score, double* var, double* best,
                        int64_t* best_idx, void* scratch, void* stream) {
  return score_common(workspace, n, d, S, X_design, Nd, idx_offset, form, kModeEkld, nullptr, logk, score, var, best, best_idx,
                      scratch, stream, nullptr);
}

int bode_ekld_score_dev_timed(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                              int64_t idx_offset, int form, double* logk, double* score, double* var, double* best,
                              int64_t* best_idx, void* scratch, void* stream, float* score_kernel_ms) {
  if (!score_kernel_ms) return BODE_ERR_BAD_ARG;
  return score_common(workspace, n, d, S, X_design, Nd, idx_offset, form, kModeEkld, nullptr, logk, score, var, best, best_idx,
                      scratch, stream, score_kernel_ms);
}

int bode_ekld_score_dev_events(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                               int64_t idx_offset, int form, double* logk, double* score, double* var, double* best,
                               int64_t* best_idx, void* scratch, void* stream, void* ev_start, void* ev_stop) {
  if (!ev_start || !ev_stop) return BODE_ERR_BAD_ARG;
  return score_common(workspace, n, d, S, X_design, Nd, idx_offset, form, kModeEkld, nullptr, logk, score, var, best, best_idx,
                      scratch, stream, nullptr, ev_start, ev_stop);
}

int bode_ekld_score_quad_dev(const void* workspace, int n, int d, int S, const double* hyp, int ndim,
                             const double* X_design, int64_t Nd, int64_t idx_offset, int form, const double* Xq,
                             const double* wq, int Nq, double* xi /* S*Nd scratch */, double* logk, double* score,
                             double* var, double* best, int64_t* best_idx, void* scratch, void* stream) {
  if (!hyp || !Xq || !xi || Nq < 1) return BODE_ERR_BAD_ARG;
  int rc = bode_quad_rowmean_dev(hyp, S, ndim, d, X_design, Nd, Xq, wq, Nq, xi, scratch, stream);  // xi_q(x) for every (s, x)
  if (rc != BODE_OK) return rc;
  return score_common(workspace, n, d, S, X_design, Nd, idx_offset, form, kModeEkld, xi, logk, score, var, best, best_idx,
                      static_cast<char*>(scratch) + 256, stream, nullptr);
}

int bode_us_variance_dev(const void* workspace, int n, int d, int S, const double* X_grid, int64_t Ng,
                         int64_t idx_offset, double* sigx, double* pred_var, double* best, int64_t* best_idx,
                         void* scratch, void* stream) {
  return score_common(workspace, n, d, S, X_grid, Ng, idx_offset, BODE_FORM_REFERENCE, kModeVariance, nullptr, sigx, pred_var,
                      nullptr, best, best_idx, scratch, stream, nullptr);
}

int bode_ekld_score_host(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                         double nugget_var, double jitter, const double* X_design, int64_t Nd, int form,
                         double* score, double* var, double* kld, int64_t* best_idx, double* mu_sigma, int* info) {
  if (!X || !Y || !hyp || !X_design || !score || !shape_ok(n, d, S) || Nd < 1) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  const size_t wsb = bode_ekld_workspace_bytes(n, d, S);
  DevBuf dX, dY, dH, dXd, dWs, dInfo, dLogk, dScore, dVar, dBest, dBestIdx, dScr, dMs;
  BODE_CUDA(dX.alloc(sizeof(double) * n * d));
  BODE_CUDA(dY.alloc(sizeof(double) * n));
  BODE_CUDA(dH.alloc(sizeof(double) * S * ndim));
  BODE_CUDA(dXd.alloc(sizeof(double) * Nd * d));
  BODE_CUDA(dWs.alloc(wsb));
  BODE_CUDA(dInfo.alloc(sizeof(int) * S));
  BODE_CUDA(dLogk.alloc(sizeof(double) * (size_t)S * Nd));
  BODE_CUDA(dScore.alloc(sizeof(double) * Nd));
  BODE_CUDA(dVar.alloc(sizeof(double) * Nd));
  BODE_CUDA(dBest.alloc(sizeof(double)));
  BODE_CUDA(dBestIdx.alloc(sizeof(int64_t)));
  BODE_CUDA(dScr.alloc(bode_ekld_score_scratch_bytes(Nd, S)));
  BODE_CUDA(dMs.alloc(2 * sizeof(double)));
  BODE_CUDA(cudaMemcpy(dX.p, X, sizeof(double) * n * d, cudaMemcpyHostToDevice));
  BODE_CUDA(cudaMemcpy(dY.p, Y, sizeof(double) * n, cudaMemcpyHostToDevice));
  BODE_CUDA(cudaMemcpy(dH.p, hyp, sizeof(double) * S * ndim, cudaMemcpyHostToDevice));
  BODE_CUDA(cudaMemcpy(dXd.p, X_design, sizeof(double) * Nd * d, cudaMemcpyHostToDevice));
  int rc = bode_ekld_prepare_dev(dX.as<double>(), dY.as<double>(), dH.as<double>(), n, d, S, ndim, nugget_var, jitter,
                                 dWs.p, wsb, dInfo.as<int>(), nullptr);
  if (rc != BODE_OK) return rc;
  rc = bode_ekld_score_dev(dWs.p, n, d, S, dXd.as<double>(), Nd, 0, form, dLogk.as<double>(), dScore.as<double>(),
                           var ? dVar.as<double>() : nullptr, dBest.as<double>(), dBestIdx.as<int64_t>(), dScr.p, nullptr);
  if (rc != BODE_OK) return rc;
  if (mu_sigma) {
    rc = bode_ekld_mu_sigma_dev(dWs.p, n, d, S, dMs.as<double>(), nullptr);
    if (rc != BODE_OK) return rc;
    BODE_CUDA(cudaMemcpy(mu_sigma, dMs.p, 2 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  BODE_CUDA(cudaMemcpy(score, dScore.p, sizeof(double) * Nd, cudaMemcpyDeviceToHost));
  if (var) BODE_CUDA(cudaMemcpy(var, dVar.p, sizeof(double) * Nd, cudaMemcpyDeviceToHost));
  if (best_idx) BODE_CUDA(cudaMemcpy(best_idx, dBestIdx.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (kld) {
    BODE_CUDA(launch_exp_inplace(dLogk.as<double>(), (size_t)S * Nd, nullptr));
    BODE_CUDA(cudaMemcpy(kld, dLogk.p, sizeof(double) * (size_t)S * Nd, cudaMemcpyDeviceToHost));
  }
  std::vector<int> hinfo(S);
  BODE_CUDA(cudaMemcpy(hinfo.data(), dInfo.p, sizeof(int) * S, cudaMemcpyDeviceToHost));
  BODE_CUDA(cudaDeviceSynchronize());
  int bad = 0;
  for (int s = 0; s < S; s++) {
    if (info) info[s] = hinfo[s];
    if (hinfo[s] != 0) bad = 1;
  }
  return bad ? BODE_ERR_NOT_PD : BODE_OK;
}

int bode_gp_loglik_host(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                        double nugget_var, double jitter, double* loglik) {
  if (!X || !Y || !hyp || !loglik || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  const size_t wsb = bode_ekld_workspace_bytes(n, d, S);
  DevBuf dX, dY, dH, dWs, dInfo, dT;
  BODE_CUDA(dX.alloc(sizeof(double) * n * d));
  BODE_CUDA(dY.alloc(sizeof(double) * n));
  BODE_CUDA(dH.alloc(sizeof(double) * S * ndim));
  BODE_CUDA(dWs.alloc(wsb));
  BODE_CUDA(dInfo.alloc(sizeof(int) * S));
  BODE_CUDA(dT.alloc(sizeof(double) * S * kConsts));
  BODE_CUDA(cudaMemcpy(dX.p, X, sizeof(double) * n * d, cudaMemcpyHostToDevice));
  BODE_CUDA(cudaMemcpy(dY.p, Y, sizeof(double) * n, cudaMemcpyHostToDevice));
  BODE_CUDA(cudaMemcpy(dH.p, hyp, sizeof(double) * S * ndim, cudaMemcpyHostToDevice));
  int rc = bode_gp_loglik_dev(dX.as<double>(), dY.as<double>(), dH.as<double>(), n, d, S, ndim, nugget_var, jitter, dWs.p, wsb,
                              dInfo.as<int>(), dT.as<double>(), nullptr);
  if (rc != BODE_OK) return rc;
  BODE_CUDA(cudaMemcpy(loglik, dT.p, sizeof(double) * S, cudaMemcpyDeviceToHost));
  return BODE_OK;
}

int bode_ekld_plan_describe(int n, int d, int S, int64_t Nd, int sm_count, size_t smem_optin, char* out, size_t cap) {
  if (!out || cap < 64 || !shape_ok(n, d, S) || Nd < 1 || sm_count < 1) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  const Layout L = make_layout(n, d, S);
  ScorePlan P;
  if (plan_score(L, Nd, sm_count, smem_optin, &P) != 0) return BODE_ERR_UNSUPPORTED;
  int w = snprintf(out, cap,
                   "{\"variant\": %d, \"row_groups\": %d, \"col_groups\": %d, \"row_tiles_per_warp\": %d, \"col_tiles_per_warp\": %d, "
                   "\"candidates_per_cta\": %d, \"threads\": %d, \"samples_per_item\": %d, \"tiles\": %d, \"items\": %lld, "
                   "\"stages\": %d, \"chunk_doubles\": %d, \"u_global\": %d, \"smem_bytes\": %zu, \"nrt\": %d, \"ng\": %d, \"chunk_g0\": [",
                   P.variant, P.rg, P.cg, P.rt, P.ct, P.C, 32 * P.rg * P.cg, P.sc, P.ntiles, (long long)P.nitems, P.stages, P.chunk_max,
                   P.u_global, P.smem, L.nrt, L.ng);
  for (int q = 0; q <= P.nchunks && w > 0 && (size_t)w < cap; q++)
    w += snprintf(out + w, cap - w, "%s%d", q ? ", " : "", (int)P.chunk_g0[q]);
  if (w > 0 && (size_t)w < cap) w += snprintf(out + w, cap - w, "], \"tile_id\": [");
  for (int g = 0; g < P.rg && w > 0 && (size_t)w < cap; g++) {
    w += snprintf(out + w, cap - w, "%s[", g ? ", " : "");
    for (int t = 0; t < P.rt && (size_t)w < cap; t++) w += snprintf(out + w, cap - w, "%s%d", t ? ", " : "", (int)P.tile_id[g * kMaxRT + t]);
    if ((size_t)w < cap) w += snprintf(out + w, cap - w, "]");
  }
  if (w > 0 && (size_t)w < cap) w += snprintf(out + w, cap - w, "]}");
  return (w > 0 && (size_t)w < cap) ? BODE_OK : BODE_ERR_BAD_ARG;
}

int bode_fp64_peak(double* dmma_tflops, double* dfma_tflops) {
  if (!dmma_tflops || !dfma_tflops) return BODE_ERR_BAD_ARG;
  BODE_CUDA(measure_fp64_peak(dmma_tflops, dfma_tflops));
  return BODE_OK;
}

}  // extern "C"

#ifdef BODE_TIMELINE
#include <stdio.h>
namespace bode { cudaError_t debug_copy_timeline(long long* host); cudaError_t debug_set_phase_a(int v); }
namespace bode { cudaError_t debug_copy_prepare_timeline(long long* host); }
extern "C" int bode_debug_prepare(long long* out) { cudaDeviceSynchronize(); return bode::debug_copy_prepare_timeline(out) == cudaSuccess ? 0 : -4; }
extern "C" int bode_debug_phase_a(int v) { return bode::debug_set_phase_a(v) == cudaSuccess ? 0 : -4; }
extern "C" int bode_debug_dump(const char* path) {
  static long long buf[8 * 16 * 16 * 14];
  if (cudaDeviceSynchronize() != cudaSuccess) return -4;
  if (bode::debug_copy_timeline(buf) != cudaSuccess) return -4;
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  fwrite(buf, sizeof(long long), 8 * 16 * 16 * 14, f);
  fclose(f);
  return 0;
}
#endif
