// Kernel-argument and launch-plan structs of the scoring kernel (shared by score.cu and capi.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "layout.h"

namespace bode {

constexpr int kComputeWarps = 8;
constexpr int kComputeThreads = kComputeWarps * 32;  // 256
constexpr int kScoreThreads = kComputeThreads + 32;  // + one producer warp
constexpr int kMaxStages = 8;
constexpr int kMaxRT = 32;      // row tiles per warp row-group (n <= 1024)
constexpr int kMaxChunks = 256;
constexpr int kXtStride = 16;   // BODE_MAX_D doubles per candidate row in shared memory
constexpr int kModeEkld = 0, kModeVariance = 1;
constexpr int kFormReference = 0, kFormLog1p = 1;

struct ScoreArgs {
  const void* ws;
  Layout L;
  const double* Xd;  // Nd * d
  int64_t Nd;
  double* logk;      // S * Nd
  int S;
  int sc;            // hyper-samples per work item
  int ntiles;        // candidate tiles
  int form, mode;
  int stages, chunk_max /* doubles */, nchunks;
  unsigned short chunk_g0[kMaxChunks + 1];
  short tile_id[4 * kMaxRT];
};

struct ScorePlan {
  int rt, ct, C;
  int sc, ntiles;
  int64_t nitems;
  int stages, chunk_max, nchunks;
  size_t smem;
  unsigned short chunk_g0[kMaxChunks + 1];
  short tile_id[4 * kMaxRT];
};

}  // namespace bode
