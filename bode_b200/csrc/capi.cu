// C ABI of bode_b200 (see include/bode_b200.h).  Thin, allocation-free glue around the kernels for the *_dev
// entry points; the *_host entry points own their device scratch for the duration of the call.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "../../include/bode_b200.h"
#include "layout.h"
#include "prepare_args.h"
#include "score_args.h"
#include "tuning.h"

namespace bode {
cudaError_t launch_prepare(const PrepArgs& a, cudaStream_t st);
cudaError_t launch_sample_terms(const void* ws, const Layout& L, double* out, cudaStream_t st);
cudaError_t launch_mu_sigma(const void* ws, const Layout& L, double* out, cudaStream_t st);
cudaError_t launch_loglik_gather(const void* ws, const Layout& L, double* out, cudaStream_t st);
int plan_score(const Layout& L, int64_t Nd, int sm_count, size_t smem_optin, ScorePlan* P);
cudaError_t launch_score(const ScoreArgs& a, const ScorePlan& P, cudaStream_t st);
cudaError_t launch_ekld_reduce(const void* ws, const Layout& L, const double* Xd, const double* xi_ovr, double* sx,
                               const double* wkplane, int S, int64_t Nd, int64_t idx_offset, int form, int keep_logk,
                               double* score, double* var, double* part_val, int64_t* part_idx, unsigned int* done_counter,
                               double* best, int64_t* best_idx, cudaStream_t st);
cudaError_t launch_reduce(const double* logk, int S, int64_t Nd, int64_t idx_offset, double* score, double* var,
                          double* part_val, int64_t* part_idx, unsigned int* done_counter, double* best, int64_t* best_idx,
                          cudaStream_t st);
cudaError_t launch_exp_inplace(double* v, size_t n, cudaStream_t st);
cudaError_t measure_fp64_peak(double* dmma, double* dfma);
cudaError_t launch_rowmean(const double* hyp, int S, int ndim, int d, const double* P, int64_t Np, const double* Xq,
                           const double* wq, int Nq, const double* wsum_ptr, double* out, cudaStream_t st);
cudaError_t launch_weighted_mean(const double* rm, const double* wq, int Nq, const double* wsum_ptr, int S, double* out,
                                 cudaStream_t st);
cudaError_t launch_sum_weights(const double* wq, int Nq, double* out, cudaStream_t st);

// ---- tuning knobs: the environment is read once (and again only on bode_tuning_reload) ----------------------------
static Tuning read_tuning() {
  auto geti = [](const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
  };
  Tuning t;
  t.score_variant = geti("BODE_SCORE_VARIANT", -1);
  t.chunk_kb = geti("BODE_CHUNK_KB", 0);
  t.waves = geti("BODE_WAVES", 0);
  t.stages = geti("BODE_STAGES", 0);
  t.debug_skip = geti("BODE_DEBUG_SKIP", 0);
  t.stagger = geti("BODE_STAGGER", 0);
  t.refill = geti("BODE_REFILL", 0);
  t.xik_series = geti("BODE_XIK_SERIES", 1);
  t.prep_variant = geti("BODE_PREP_VARIANT", -1);
  t.rowmean_variant = geti("BODE_ROWMEAN_VARIANT", -1);
  return t;
}
static Tuning g_tuning = read_tuning();
const Tuning& tuning() { return g_tuning; }
void tuning_reload() { g_tuning = read_tuning(); }
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

// ---- reusable context of the host-buffer entry points -------------------------------------------------------------
// Holds the device buffers (grow-only), a stream and nothing else: after the first call of a given size the host entry
// points make no cudaMalloc / cudaFree (the reference's loop calls them once per outer iteration).
struct bode_ctx {
  std::mutex mu;
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // the design block's host->device copy overlaps the factorisations
  cudaEvent_t copied = nullptr, inputs_in = nullptr;
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  };
  enum { kX, kY, kH, kXd, kWs, kInfo, kPlane, kScore, kVar, kBest, kScr, kMs, kLl, kNumBufs };
  Buf b[kNumBufs];
  std::vector<int> hinfo;
  long long n_alloc = 0;  // cudaMalloc calls made so far (tests: stays constant across repeated calls)
  cudaError_t need(int i, size_t bytes) {
    if (bytes == 0) bytes = 8;
    if (b[i].cap >= bytes) return cudaSuccess;
    if (b[i].p) cudaFree(b[i].p);
    b[i].p = nullptr;
    b[i].cap = 0;
    const size_t cap = align_up(bytes + bytes / 4, 256);
    cudaError_t e = cudaMalloc(&b[i].p, cap);
    if (e == cudaSuccess) { b[i].cap = cap; n_alloc++; }
    return e;
  }
  template <typename T> T* as(int i) { return static_cast<T*>(b[i].p); }
};

static bode_ctx* default_ctx() {
  static bode_ctx ctx;  // process-wide context behind the ctx-less host entry points
  return &ctx;
}

static int ctx_bind(bode_ctx* c) {
  int dev = 0;
  BODE_CUDA(cudaGetDevice(&dev));
  if (c->device < 0) {
    c->device = dev;
    BODE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    BODE_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    BODE_CUDA(cudaEventCreateWithFlags(&c->copied, cudaEventDisableTiming));
    BODE_CUDA(cudaEventCreateWithFlags(&c->inputs_in, cudaEventDisableTiming));
  } else if (c->device != dev) {
    return BODE_ERR_BAD_ARG;  // a context belongs to the device it was first used on
  }
  return BODE_OK;
}

// ---- NCCL through dlopen: the library stays loadable (and single-GPU callers unaffected) without NCCL present ------
// Only the five entry points the argmax exchange needs; types follow nccl.h (ncclUniqueId is 128 opaque bytes,
// ncclResult_t / ncclDataType_t are ints, ncclChar = 0).
struct BodeNcclId {
  char internal[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, BodeNcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // torch's bundled libnccl.so.2 is already mapped in a torch process: the bare SONAME resolves to it
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.handle, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(void**, int, BodeNcclId, int)>(dlsym(api.handle, "ncclCommInitRank"));
    api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(api.handle, "ncclAllGather"));
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.handle, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.handle, "ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) api.handle = nullptr;
  });
  return api.handle ? &api : nullptr;
}
static int nccl_fail(int r, const char* where) {
  NcclApi* api = nccl_api();
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: NCCL error %d (%s)", where, r,
           (api && api->GetErrorString) ? api->GetErrorString(r) : "?");
  return BODE_ERR_NCCL;
}

// np.argmax ordering of (value, index) records: NaN counts as the maximum, ties go to the lowest index
static bool rec_better(double av, int64_t ai, double bv, int64_t bi) {
  const bool an = av != av, bn = bv != bv;
  if (an || bn) {
    if (an && bn) return ai < bi;
    return an;
  }
  if (av > bv) return true;
  if (av < bv) return false;
  return ai < bi;
}


// ---- peer-memory argmax exchange (no NCCL on the data path) -------------------------------------------------------
// Every rank owns a mailbox of 2 (epoch parity) x world slots of 32 bytes {best, best_idx, epoch, pad}, mapped into
// every other rank's address space (CUDA IPC).  One launch of one warp per step: lane r stores this rank's record into
// slot [parity][my rank] of rank r's mailbox (NVLink peer stores, system-scope release of the epoch word), then waits
// for slot [parity][r] of its own mailbox to carry the current epoch and copies it to `gathered`.  The epoch lives in
// device memory (the launch is replayed from CUDA graphs with frozen arguments); slots alternate with its parity, so
// a rank two steps ahead cannot exist (it would need this rank's flag of the step in between).
constexpr int kMaxXchgRanks = 32;  // one lane per rank

struct XchgDev {
  unsigned long long* peer[kMaxXchgRanks];  // mailbox of every rank in this process's address space
  unsigned long long* epoch;               // this rank's step counter
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32) k_argmax_exchange(XchgDev x, const unsigned long long* __restrict__ rec,
                                                        unsigned long long* __restrict__ gathered, unsigned long long timeout_ns) {
  const int lane = threadIdx.x;
  const unsigned long long e = *x.epoch + 1ull;
  const int par = (int)(e & 1ull);
  __syncwarp();
  if (lane < x.world) {
    unsigned long long* dst = x.peer[lane] + (size_t)(par * x.world + x.rank) * 4;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(rec[0]) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(rec[1]) : "memory");
    st_release_sys(dst + 2, e);
    const unsigned long long* src = x.peer[x.rank] + (size_t)(par * x.world + lane) * 4;
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    while (ld_acquire_sys(src + 2) != e) {
      if (globaltimer_ns() - t0 > timeout_ns) { ok = false; break; }
      __nanosleep(64);
    }
    unsigned long long v0, v1;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v0) : "l"(src) : "memory");
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v1) : "l"(src + 1) : "memory");
    // a peer that never arrived: NaN with index -2 (bode_argmax_combine reports it instead of picking a winner)
    gathered[2 * lane] = ok ? v0 : 0x7ff8000000000000ull;
    gathered[2 * lane + 1] = ok ? v1 : (unsigned long long)(-2ll);
  }
  __syncwarp();
  if (lane == 0) *x.epoch = e;
}

struct bode_xchg {
  XchgDev dev;
  unsigned long long* mailbox = nullptr;  // own allocation
  int device = -1;
};

extern "C" {

const char* bode_version(void) { return "bode_b200 0.2.0 (sm_100a)"; }

const char* bode_strerror(int status) {
  switch (status) {
    case BODE_OK: return "ok";
    case BODE_ERR_BAD_ARG: return "bad argument";
    case BODE_ERR_UNSUPPORTED: return "unsupported shape (n > 1024 or d > 16)";
    case BODE_ERR_WORKSPACE: return "workspace too small or misaligned";
    case BODE_ERR_CUDA: return "CUDA error";
    case BODE_ERR_NOT_PD: return "covariance matrix not positive definite";
    case BODE_ERR_NCCL: return "NCCL unavailable or NCCL error (see bode_last_cuda_error)";
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
  a.nugget_var = nugget_var; a.jitter = jitter; a.jitter_s = nullptr; a.x_smem = 0;
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

int bode_ekld_prepare_jitter_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                                 double nugget_var, const double* jitter_per_sample, void* workspace, size_t workspace_bytes,
                                 int* info, void* stream) {
  if (!X || !Y || !hyp || !workspace || !info || !jitter_per_sample || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  PrepArgs a;
  a.X = X; a.Y = Y; a.hyp = hyp; a.n = n; a.d = d; a.S = S; a.ndim = ndim;
  a.nugget_var = nugget_var; a.jitter = 0.0; a.jitter_s = jitter_per_sample; a.x_smem = 0;
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

static int prepare_quad_common(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                               double nugget_var, double jitter, const double* Xq, const double* wq, int Nq, const double* rmq_in,
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
  if (!rmq_in) BODE_CUDA(launch_rowmean(hyp, S, ndim, d, Xq, Nq, Xq, wq, Nq, wsp, rmq, st));  // mean_q' k(x_q, x_q')
  BODE_CUDA(launch_weighted_mean(rmq_in ? rmq_in : rmq, wq, Nq, wsp, S, s0, st));             // sigma0_q
  PrepArgs a;
  a.X = X; a.Y = Y; a.hyp = hyp; a.n = n; a.d = d; a.S = S; a.ndim = ndim;
  a.nugget_var = nugget_var; a.jitter = jitter; a.jitter_s = nullptr; a.x_smem = 0;
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

int bode_ekld_prepare_quad_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                               double nugget_var, double jitter, const double* Xq, const double* wq, int Nq,
                               void* workspace, size_t workspace_bytes, int* info, void* scratch, void* stream) {
  return prepare_quad_common(X, Y, hyp, n, d, S, ndim, nugget_var, jitter, Xq, wq, Nq, nullptr, workspace, workspace_bytes, info,
                             scratch, stream);
}

int bode_ekld_prepare_quad_rm_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                                  double nugget_var, double jitter, const double* Xq, const double* wq, int Nq,
                                  const double* rowmean_q, void* workspace, size_t workspace_bytes, int* info, void* scratch,
                                  void* stream) {
  if (!rowmean_q) return BODE_ERR_BAD_ARG;
  return prepare_quad_common(X, Y, hyp, n, d, S, ndim, nugget_var, jitter, Xq, wq, Nq, rowmean_q, workspace, workspace_bytes, info,
                             scratch, stream);
}

int bode_gp_loglik_dev(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                       double nugget_var, double jitter, void* workspace, size_t workspace_bytes, int* info, double* loglik,
                       void* stream) {
  if (!X || !Y || !hyp || !workspace || !info || !loglik || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  PrepArgs a;
  a.X = X; a.Y = Y; a.hyp = hyp; a.n = n; a.d = d; a.S = S; a.ndim = ndim;
  a.nugget_var = nugget_var; a.jitter = jitter; a.jitter_s = nullptr; a.x_smem = 0;
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

// scratch = [done counter: 256 B] [block argmax values] [block argmax indices] [w.k plane: S*Nd] [sigma_x plane: S*Nd, only
// when the caller passes no logk buffer]
static size_t score_partials_bytes(int64_t Nd) {
  const size_t nb = (size_t)((Nd + 31) / 32);  // an upper bound for both reduction kernels' block counts
  return 256 + align_up(nb * sizeof(double), 256) + align_up(nb * sizeof(int64_t), 256);
}

size_t bode_ekld_score_scratch_bytes(int64_t Nd, int S) {
  if (Nd < 1 || S < 0) return 0;
  return score_partials_bytes(Nd) + align_up(sizeof(double) * (size_t)S * (size_t)Nd, 256);
}

static int score_common(const void* workspace, int n, int d, int S, const double* Xd, int64_t Nd, int64_t idx_offset,
                        int form, int mode, const double* xi_ovr, double* logk, double* score, double* var, double* best,
                        int64_t* best_idx, void* scratch, void* stream, float* kernel_ms, void* ev_start = nullptr,
                        void* ev_stop = nullptr, bool reuse_planes = false) {
  if (!workspace || !Xd || !score || !best || !best_idx || !scratch || !shape_ok(n, d, S) || Nd < 1)
    return BODE_ERR_BAD_ARG;
  if (form != BODE_FORM_REFERENCE && form != BODE_FORM_LOG1P) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  DevProps props;
  int rc = get_props(&props);
  if (rc != BODE_OK) return rc;
  ScorePlan P;
  const Layout L = make_layout(n, d, S);
  if (plan_score(L, Nd, props.sm_count, props.smem_optin, &P) != 0) return BODE_ERR_UNSUPPORTED;
  char* scr = static_cast<char*>(scratch);
  const size_t nb = (size_t)((Nd + 31) / 32);
  unsigned int* done_counter = reinterpret_cast<unsigned int*>(scr);
  double* part_val = reinterpret_cast<double*>(scr + 256);
  int64_t* part_idx = reinterpret_cast<int64_t*>(scr + 256 + align_up(nb * sizeof(double), 256));
  // two S x Nd planes between the kernels: w.k (scratch) and sigma_x (the caller's logk buffer if given -- it is turned
  // into log EKLD in place -- else the second plane of scratch)
  double* wkp = reinterpret_cast<double*>(scr + score_partials_bytes(Nd));
  double* sxp = logk ? logk : wkp + (size_t)S * (size_t)Nd;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ScoreArgs a;
  a.ws = workspace; a.L = L; a.Xd = Xd; a.Nd = Nd; a.S = S;
  a.sx = sxp;
  a.wk = wkp;
  a.done_counter = done_counter;
  a.sc = P.sc; a.ntiles = P.ntiles; a.form = form; a.mode = mode;
  a.n_main = P.n_main; a.n_tail = P.n_tail;
  memcpy(a.tail_s0, P.tail_s0, sizeof(a.tail_s0));
  a.debug_skip = tuning().debug_skip;
  a.stagger = tuning().stagger;
  a.refill_last = tuning().refill;
  a.stages = P.stages; a.chunk_max = P.chunk_max; a.nchunks = P.nchunks; a.u_global = P.u_global;
  memcpy(a.chunk_g0, P.chunk_g0, sizeof(a.chunk_g0));
  memcpy(a.tile_id, P.tile_id, sizeof(a.tile_id));
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (kernel_ms) {
    BODE_CUDA(cudaEventCreate(&ev0));
    BODE_CUDA(cudaEventCreate(&ev1));
    BODE_CUDA(cudaEventRecord(ev0, st));
  }
  if (ev_start) BODE_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_start), st));
  if (!reuse_planes) BODE_CUDA(launch_score(a, P, st));
  if (ev_stop) BODE_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev_stop), st));
  if (kernel_ms) BODE_CUDA(cudaEventRecord(ev1, st));
  if (mode == kModeEkld)
    BODE_CUDA(launch_ekld_reduce(workspace, L, Xd, xi_ovr, sxp, wkp, S, Nd, idx_offset, form, logk != nullptr, score, var, part_val,
                                 part_idx, done_counter, best, best_idx, st));
  else
    BODE_CUDA(launch_reduce(sxp, S, Nd, idx_offset, score, var, part_val, part_idx, done_counter, best, best_idx, st));
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

int bode_ekld_rescore_dev(const void* workspace, int n, int d, int S, const double* X_design, int64_t Nd,
                          int64_t idx_offset, int form, double* score, double* var, double* best, int64_t* best_idx,
                          void* scratch, void* stream) {
  return score_common(workspace, n, d, S, X_design, Nd, idx_offset, form, kModeEkld, nullptr, nullptr, score, var, best, best_idx,
                      scratch, stream, nullptr, nullptr, nullptr, true);
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

int bode_ctx_create(bode_ctx** out) {
  if (!out) return BODE_ERR_BAD_ARG;
  *out = new (std::nothrow) bode_ctx();
  return *out ? BODE_OK : BODE_ERR_BAD_ARG;
}

int bode_ctx_destroy(bode_ctx* c) {
  if (!c || c == default_ctx()) return BODE_ERR_BAD_ARG;
  for (auto& buf : c->b)
    if (buf.p) cudaFree(buf.p);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->copied) cudaEventDestroy(c->copied);
  if (c->inputs_in) cudaEventDestroy(c->inputs_in);
  delete c;
  return BODE_OK;
}

long long bode_ctx_alloc_count(bode_ctx* c) {
  if (!c) c = default_ctx();
  std::lock_guard<std::mutex> lk(c->mu);
  return c->n_alloc;
}

int bode_ekld_score_host_ctx(bode_ctx* c, const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                             double nugget_var, double jitter, const double* X_design, int64_t Nd, int form,
                             double* score, double* var, double* kld, int64_t* best_idx, double* mu_sigma, int* info) {
  if (!X || !Y || !hyp || !X_design || !score || !shape_ok(n, d, S) || Nd < 1) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  if (!c) c = default_ctx();
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = ctx_bind(c);
  if (rc != BODE_OK) return rc;
  typedef bode_ctx B;
  const size_t wsb = bode_ekld_workspace_bytes(n, d, S);
  BODE_CUDA(c->need(B::kX, sizeof(double) * n * d));
  BODE_CUDA(c->need(B::kY, sizeof(double) * n));
  BODE_CUDA(c->need(B::kH, sizeof(double) * S * ndim));
  BODE_CUDA(c->need(B::kXd, sizeof(double) * Nd * d));
  BODE_CUDA(c->need(B::kWs, wsb));
  BODE_CUDA(c->need(B::kInfo, sizeof(int) * S));
  BODE_CUDA(c->need(B::kPlane, sizeof(double) * (size_t)S * Nd));
  BODE_CUDA(c->need(B::kScore, sizeof(double) * Nd));
  BODE_CUDA(c->need(B::kVar, sizeof(double) * Nd));
  BODE_CUDA(c->need(B::kBest, 64));
  BODE_CUDA(c->need(B::kScr, bode_ekld_score_scratch_bytes(Nd, S)));
  BODE_CUDA(c->need(B::kMs, 2 * sizeof(double)));
  cudaStream_t st = c->stream;
  BODE_CUDA(cudaMemcpyAsync(c->b[B::kX].p, X, sizeof(double) * n * d, cudaMemcpyHostToDevice, st));
  BODE_CUDA(cudaMemcpyAsync(c->b[B::kY].p, Y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  BODE_CUDA(cudaMemcpyAsync(c->b[B::kH].p, hyp, sizeof(double) * S * ndim, cudaMemcpyHostToDevice, st));
  BODE_CUDA(cudaEventRecord(c->inputs_in, st));
  rc = bode_ekld_prepare_dev(c->as<double>(B::kX), c->as<double>(B::kY), c->as<double>(B::kH), n, d, S, ndim, nugget_var, jitter,
                             c->b[B::kWs].p, wsb, c->as<int>(B::kInfo), st);
  if (rc != BODE_OK) return rc;
  // the design block goes up on a second stream while the factorisations run (behind the small copies: the host->device
  // DMA queue is served in order; after the launch: a copy from pageable memory blocks the host).  The previous call's
  // scoring kernel, the last reader of the buffer, has finished: every call ends synchronised.
  BODE_CUDA(cudaStreamWaitEvent(c->copy_stream, c->inputs_in, 0));
  BODE_CUDA(cudaMemcpyAsync(c->b[B::kXd].p, X_design, sizeof(double) * Nd * d, cudaMemcpyHostToDevice, c->copy_stream));
  BODE_CUDA(cudaEventRecord(c->copied, c->copy_stream));
  BODE_CUDA(cudaStreamWaitEvent(st, c->copied, 0));
  double* dbest = c->as<double>(B::kBest);
  int64_t* dbest_idx = reinterpret_cast<int64_t*>(dbest + 1);
  rc = bode_ekld_score_dev(c->b[B::kWs].p, n, d, S, c->as<double>(B::kXd), Nd, 0, form, c->as<double>(B::kPlane),
                           c->as<double>(B::kScore), var ? c->as<double>(B::kVar) : nullptr, dbest, dbest_idx, c->b[B::kScr].p, st);
  if (rc != BODE_OK) return rc;
  if (mu_sigma) {
    rc = bode_ekld_mu_sigma_dev(c->b[B::kWs].p, n, d, S, c->as<double>(B::kMs), st);
    if (rc != BODE_OK) return rc;
    BODE_CUDA(cudaMemcpyAsync(mu_sigma, c->b[B::kMs].p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  BODE_CUDA(cudaMemcpyAsync(score, c->b[B::kScore].p, sizeof(double) * Nd, cudaMemcpyDeviceToHost, st));
  if (var) BODE_CUDA(cudaMemcpyAsync(var, c->b[B::kVar].p, sizeof(double) * Nd, cudaMemcpyDeviceToHost, st));
  if (best_idx) BODE_CUDA(cudaMemcpyAsync(best_idx, dbest_idx, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (kld) {
    BODE_CUDA(launch_exp_inplace(c->as<double>(B::kPlane), (size_t)S * Nd, st));
    BODE_CUDA(cudaMemcpyAsync(kld, c->b[B::kPlane].p, sizeof(double) * (size_t)S * Nd, cudaMemcpyDeviceToHost, st));
  }
  c->hinfo.resize(S);
  BODE_CUDA(cudaMemcpyAsync(c->hinfo.data(), c->b[B::kInfo].p, sizeof(int) * S, cudaMemcpyDeviceToHost, st));
  BODE_CUDA(cudaStreamSynchronize(st));
  int bad = 0;
  for (int s = 0; s < S; s++) {
    if (info) info[s] = c->hinfo[s];
    if (c->hinfo[s] != 0) bad = 1;
  }
  return bad ? BODE_ERR_NOT_PD : BODE_OK;
}

int bode_ekld_score_host(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                         double nugget_var, double jitter, const double* X_design, int64_t Nd, int form,
                         double* score, double* var, double* kld, int64_t* best_idx, double* mu_sigma, int* info) {
  return bode_ekld_score_host_ctx(nullptr, X, Y, hyp, n, d, S, ndim, nugget_var, jitter, X_design, Nd, form, score, var, kld,
                                  best_idx, mu_sigma, info);
}

int bode_gp_loglik_host_ctx(bode_ctx* c, const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                            double nugget_var, double jitter, double* loglik) {
  if (!X || !Y || !hyp || !loglik || !shape_ok(n, d, S)) return BODE_ERR_BAD_ARG;
  if (ndim != d + 1 && ndim != d + 2) return BODE_ERR_BAD_ARG;
  if (!shape_supported(n, d)) return BODE_ERR_UNSUPPORTED;
  if (!c) c = default_ctx();
  std::lock_guard<std::mutex> lk(c->mu);
  int rc = ctx_bind(c);
  if (rc != BODE_OK) return rc;
  typedef bode_ctx B;
  const size_t wsb = bode_ekld_workspace_bytes(n, d, S);
  BODE_CUDA(c->need(B::kX, sizeof(double) * n * d));
  BODE_CUDA(c->need(B::kY, sizeof(double) * n));
  BODE_CUDA(c->need(B::kH, sizeof(double) * S * ndim));
  BODE_CUDA(c->need(B::kWs, wsb));
  BODE_CUDA(c->need(B::kInfo, sizeof(int) * S));
  BODE_CUDA(c->need(B::kLl, sizeof(double) * S));
  cudaStream_t st = c->stream;
  BODE_CUDA(cudaMemcpyAsync(c->b[B::kX].p, X, sizeof(double) * n * d, cudaMemcpyHostToDevice, st));
  BODE_CUDA(cudaMemcpyAsync(c->b[B::kY].p, Y, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  BODE_CUDA(cudaMemcpyAsync(c->b[B::kH].p, hyp, sizeof(double) * S * ndim, cudaMemcpyHostToDevice, st));
  rc = bode_gp_loglik_dev(c->as<double>(B::kX), c->as<double>(B::kY), c->as<double>(B::kH), n, d, S, ndim, nugget_var, jitter,
                          c->b[B::kWs].p, wsb, c->as<int>(B::kInfo), c->as<double>(B::kLl), st);
  if (rc != BODE_OK) return rc;
  BODE_CUDA(cudaMemcpyAsync(loglik, c->b[B::kLl].p, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
  BODE_CUDA(cudaStreamSynchronize(st));
  return BODE_OK;
}

int bode_gp_loglik_host(const double* X, const double* Y, const double* hyp, int n, int d, int S, int ndim,
                        double nugget_var, double jitter, double* loglik) {
  return bode_gp_loglik_host_ctx(nullptr, X, Y, hyp, n, d, S, ndim, nugget_var, jitter, loglik);
}

int bode_tuning_reload(void) {
  tuning_reload();
  return BODE_OK;
}

// ---- multi-GPU argmax exchange: one 16-byte all-gather on the compute stream --------------------------
int bode_nccl_unique_id(void* id128) {
  if (!id128) return BODE_ERR_BAD_ARG;
  NcclApi* api = nccl_api();
  if (!api) return BODE_ERR_NCCL;
  const int r = api->GetUniqueId(id128);
  return r == 0 ? BODE_OK : nccl_fail(r, "ncclGetUniqueId");
}

int bode_nccl_comm_create(const void* id128, int world, int rank, void** comm) {
  if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) return BODE_ERR_BAD_ARG;
  NcclApi* api = nccl_api();
  if (!api) return BODE_ERR_NCCL;
  BodeNcclId id;
  memcpy(&id, id128, sizeof(id));
  const int r = api->CommInitRank(comm, world, id, rank);
  return r == 0 ? BODE_OK : nccl_fail(r, "ncclCommInitRank");
}

int bode_nccl_comm_destroy(void* comm) {
  if (!comm) return BODE_ERR_BAD_ARG;
  NcclApi* api = nccl_api();
  if (!api) return BODE_ERR_NCCL;
  const int r = api->CommDestroy(comm);
  return r == 0 ? BODE_OK : nccl_fail(r, "ncclCommDestroy");
}

int bode_ekld_argmax_nccl(void* comm, const void* best_rec, void* gathered, void* stream) {
  if (!comm || !best_rec || !gathered) return BODE_ERR_BAD_ARG;
  NcclApi* api = nccl_api();
  if (!api) return BODE_ERR_NCCL;
  const int r = api->AllGather(best_rec, gathered, 16, /*ncclChar*/ 0, comm, static_cast<cudaStream_t>(stream));
  return r == 0 ? BODE_OK : nccl_fail(r, "ncclAllGather");
}

int bode_xchg_create(int world, int rank, bode_xchg** out, unsigned char handle[64]) {
  if (!out || !handle || world < 1 || world > kMaxXchgRanks || rank < 0 || rank >= world) return BODE_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  bode_xchg* x = new bode_xchg();
  memset(&x->dev, 0, sizeof(x->dev));
  x->dev.world = world;
  x->dev.rank = rank;
  const size_t bytes = (size_t)2 * world * 32 + 64;  // the slots, then the epoch counter
  cudaError_t e = cudaGetDevice(&x->device);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->mailbox), bytes);
  if (e == cudaSuccess) e = cudaMemset(x->mailbox, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, x->mailbox);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (x->mailbox) cudaFree(x->mailbox);
    delete x;
    return cuda_fail(e, "bode_xchg_create");
  }
  memcpy(handle, &h, 64);
  x->dev.peer[rank] = x->mailbox;
  x->dev.epoch = x->mailbox + (size_t)2 * world * 4;
  *out = x;
  return BODE_OK;
}

int bode_xchg_connect(bode_xchg* x, const unsigned char* handles /* world x 64 bytes, rank order */) {
  if (!x || !handles) return BODE_ERR_BAD_ARG;
  for (int r = 0; r < x->dev.world; r++) {
    if (r == x->dev.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)64 * r, 64);
    void* p = nullptr;
    BODE_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->dev.peer[r] = static_cast<unsigned long long*>(p);
  }
  return BODE_OK;
}

int bode_xchg_destroy(bode_xchg* x) {
  if (!x) return BODE_ERR_BAD_ARG;
  for (int r = 0; r < x->dev.world; r++)
    if (r != x->dev.rank && x->dev.peer[r]) cudaIpcCloseMemHandle(x->dev.peer[r]);
  if (x->mailbox) cudaFree(x->mailbox);
  delete x;
  return BODE_OK;
}

int bode_ekld_argmax_p2p(bode_xchg* x, const void* best_rec, void* gathered, void* stream) {
  if (!x || !best_rec || !gathered) return BODE_ERR_BAD_ARG;
  int dev = -1;
  BODE_CUDA(cudaGetDevice(&dev));
  if (dev != x->device) return BODE_ERR_BAD_ARG;  // the mailbox belongs to the device it was created on
  k_argmax_exchange<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(x->dev, static_cast<const unsigned long long*>(best_rec),
                                                                     static_cast<unsigned long long*>(gathered),
                                                                     5000000000ull /* 5 s: a peer that does not show up */);
  BODE_CUDA(cudaGetLastError());
  return BODE_OK;
}

int bode_argmax_combine(const void* records, int count, double* best, int64_t* best_idx) {
  if (!records || count < 1 || !best || !best_idx) return BODE_ERR_BAD_ARG;
  const char* p = static_cast<const char*>(records);
  double bv = 0.0;
  int64_t bi = -1;
  for (int i = 0; i < count; i++) {
    double v;
    int64_t idx;
    memcpy(&v, p + 16 * i, 8);
    memcpy(&idx, p + 16 * i + 8, 8);
    if (idx == -2) return BODE_ERR_NCCL;  // bode_ekld_argmax_p2p: a peer did not deliver its record in time
    if (idx < 0) continue;  // a rank without candidates
    if (bi < 0 || rec_better(v, idx, bv, bi)) { bv = v; bi = idx; }
  }
  *best = bv;
  *best_idx = bi;
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
  if (w > 0 && (size_t)w < cap) w += snprintf(out + w, cap - w, "], \"main_chunks\": %d, \"tail_s0\": [", P.n_main);
  for (int j = 0; j <= P.n_tail && P.n_tail > 0 && w > 0 && (size_t)w < cap; j++)
    w += snprintf(out + w, cap - w, "%s%d", j ? ", " : "", (int)P.tail_s0[j]);
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
namespace bode { cudaError_t debug_copy_prepare_timeline(long long* host); cudaError_t debug_copy_prepare_tiles_timeline(long long* host); }
extern "C" int bode_debug_prepare_tiles(long long* out) { cudaDeviceSynchronize(); return bode::debug_copy_prepare_tiles_timeline(out) == cudaSuccess ? 0 : -4; }
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
