// psnode_generic_fwd_tb2.cu -- the generic forward integrator compiled with 2 trajectories per CTA (psnode_generic.cuh):
// the fallback of the fallback, for nets whose per-trajectory vectors are too wide for 8 trajectories per CTA
// (neural_01_DAE_02_direct_encode.py:61-121 at hidden = 256: BASELINE configs[4]).
#define PSN_G_TB 2
#define PSN_G_NAME(x) x##_tb2
#include "psnode_generic_fwd.cu"
