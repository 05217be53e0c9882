// Developer tuning knobs (DESIGN.md "Tuning knobs"): environment variables read ONCE per process, not per call.
// bode_tuning_reload() (C ABI) re-reads them; the tests that force a kernel variant call it after changing the environment.
#pragma once

namespace bode {

struct Tuning {
  int score_variant;  // BODE_SCORE_VARIANT  (-1: planner's choice)
  int chunk_kb;       // BODE_CHUNK_KB       (0: balanced chunks)
  int waves;          // BODE_WAVES          (target CTA waves of the scoring kernel)
  int stages;         // BODE_STAGES         (ring stages, >= 2)
  int debug_skip;     // BODE_DEBUG_SKIP     (timing decomposition only)
  int stagger;        // BODE_STAGGER
  int refill;         // BODE_REFILL         (1: the last warp to leave a ring stage refills it; 0: rotating duty warp)
  int xik_series;     // BODE_XIK_SERIES     (0: evaluate every xik factor with erf)
  int rowmean_variant;  // BODE_ROWMEAN_VARIANT (0: thread-per-point row means instead of the tensor-pipe variant)
  int prep_variant;   // BODE_PREP_VARIANT   (0: global-memory prepare kernel even where the shared-memory tile kernel fits)
};

const Tuning& tuning();
void tuning_reload();

}  // namespace bode
