// Strict instantiation: compiled with --fmad=false -prec-div=true -prec-sqrt=true -ftz=false so that every
// f32 operation is the single IEEE operation the reference's Rust code performs (no FMA contraction).
#define BVHT_MODE_NS strict
#define BVHT_LAUNCH(name) name##_strict
#include "trace_instantiate.inc"
