// Translation unit of aff_fast_kernel<K, G, BT> (aff_fast_kernels.cuh) and aff_x2_kernel<K, G> (aff_x2_kernels.cuh).
#define POYB200_DEFINE_AFF_FAST
#include "launch.h"
#include "aff_fast_kernels.cuh"
#include "aff_x2_kernels.cuh"
