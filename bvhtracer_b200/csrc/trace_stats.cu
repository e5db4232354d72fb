// Debug instantiation with per-ray work counters (bvht_debug_trace_stats): strict arithmetic + BVHT_STATS.
// Not used by any product path; it exists so that DESIGN.md's per-ray work figures are measured, not guessed.
#define BVHT_STATS 1
#define BVHT_MODE_NS stats
#define BVHT_LAUNCH(name) name##_stats
#include "trace_instantiate.inc"
