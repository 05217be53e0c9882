// Kernel-argument and launch-plan structs of the scoring kernel (shared by score.cu and capi.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "layout.h"

namespace bode {

constexpr int kMaxStages = 8;
constexpr int kMaxRT = 16;      // row tiles per warp row-group
constexpr int kMaxRG = 16;      // row groups
constexpr int kMaxChunks = 256;
constexpr int kMaxTail = 24;    // tapered hyper-sample chunks of the work-item schedule (halving: <= log2(S) + 2 entries)
constexpr int kModeEkld = 0, kModeVariance = 1;
constexpr int kFormReference = 0, kFormLog1p = 1;

struct ScoreArgs {
  const void* ws;
  Layout L;
  const double* Xd;  // Nd * d
  int64_t Nd;
  double* sx;        // S * Nd: sigma_x, the latent predictive variance (every mode)
  double* wk;        // S * Nd: w.k = e_k^T A^-1 k(X, x), the data term of v (EKLD mode only); k_ekld_reduce finishes the EKLD
  unsigned int* done_counter;  // zeroed here for the block-argmax reduction kernel that follows on the same stream
  int S;
  int sc;            // hyper-samples per work item of the n_main equal chunks [i sc, (i+1) sc) ...
  int n_main, n_tail;                    // ... followed by n_tail tapered chunks [tail_s0[j], tail_s0[j+1])
  int tail_s0[kMaxTail + 1];
  int ntiles;        // candidate tiles
  int form, mode;
  int debug_skip;    // developer timing knob (env BODE_DEBUG_SKIP): 1 = skip phase A after the first sample, 2 = skip phase B's DMMAs
  int stages, chunk_max /* doubles */, nchunks;
  int stagger;       // developer knob (env BODE_STAGGER): odd column groups start one phase A late
  int refill_last;   // 1: the last warp to leave a ring stage issues its refill (shared-memory counter, nobody waits); 0: duty warp
  int u_global;      // 1: the tiled observations Ut and the (nun, w) pairs stay in global memory (large n)
  unsigned short chunk_g0[kMaxChunks + 1];
  short tile_id[kMaxRG * kMaxRT];  // [rg][slot], right-aligned ascending, -1 = unused
};

struct ScorePlan {
  int variant;  // 0: n<=256 (RG4 CG4 RT8 CT2), 1: n<=512 (RG4 CG4 RT16 CT1), 2: RG8 CG2 RT16 CT1, 5: RG16 CG1 RT8 CT1 (n<=1024)
  int rg, cg, rt, ct, C;
  int sc, ntiles;
  int n_main, n_tail;
  int tail_s0[kMaxTail + 1];
  int64_t nitems;
  int stages, chunk_max, nchunks, u_global;
  size_t smem;
  unsigned short chunk_g0[kMaxChunks + 1];
  short tile_id[kMaxRG * kMaxRT];
};

}  // namespace bode
