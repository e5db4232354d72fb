// Fast instantiation: compiled with --fmad=true (FMA contraction allowed in every culling test); the reported
// (t, u, v) of the winning triangle are re-evaluated without contraction (mt_exact_rn).
#define BVHT_FAST_MODE 1
#define BVHT_MODE_NS fast
#define BVHT_LAUNCH(name) name##_fast
#include "trace_instantiate.inc"
