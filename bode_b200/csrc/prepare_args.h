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
  char* ws;
  Layout L;
  int* info;  // S
};
}  // namespace bode
