// Translation unit of aff_fast_kernel<K, G, BT> (aff_fast_kernels.cuh).
#define POYB200_DEFINE_AFF_FAST
#include "launch.h"
#include "aff_fast_kernels.cuh"
