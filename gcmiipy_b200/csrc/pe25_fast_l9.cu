// fused 2.5-D half step for 9 layers (see pe25_fast_impl.h)
#include "pe25_fast_impl.h"
GCM_PF_INSTANTIATE(9)
