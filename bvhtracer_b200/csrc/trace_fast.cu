// Fast instantiation: compiled with --fmad=true (FMA contraction allowed).
#define BVHT_MODE_NS fast
#define BVHT_LAUNCH(name) name##_fast
#include "trace_instantiate.inc"
