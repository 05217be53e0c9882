// Kernel arguments of the prepare kernel (shared by prepare.cu and capi.cu).
#pragma once
#include "layout.h"

namespace bode {
struct PrepArgs {
  const double* X;    // n*d
  const double* Y;    // n
  const double* hyp;  // S*ndim
  int n, d, S, ndim;
  double nugget_var, jitter;
  const double* jitter_s;    // S (nullable): per-sample diagonal addition replacing `jitter` (GPy's jitchol retries)
  char* ws;
  Layout L;
  int* info;  // S
  // quadrature mode (nullable): kernel mean embedding of the observations and double integral supplied by k_rowmean
  const double* e_ovr;       // S*n   replaces xik(X)
  const double* sigma0_ovr;  // S     replaces bk
  int x_smem;                // (set by the tile kernel's launcher) the scaled observation tiles are staged in shared memory
  int loglik_only;           // 1: stop after the factorisation scalars (host MCMC likelihoods): no w, no L^-1, no tiling
};
}  // namespace bode
