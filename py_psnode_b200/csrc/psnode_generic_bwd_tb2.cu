// psnode_generic_bwd_tb2.cu -- the generic reverse sweep compiled with 2 trajectories per CTA (see psnode_generic_fwd_tb2.cu).
#define PSN_G_TB 2
#define PSN_G_NAME(x) x##_tb2
#include "psnode_generic_bwd.cu"
