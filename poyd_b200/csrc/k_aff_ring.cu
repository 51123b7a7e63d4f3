// Translation unit of aff_ring_kernel<K, G, BT, EBF> (aff_ring_kernels.cuh).
#define POYB200_DEFINE_AFF_RING
#include "launch.h"
#include "aff_ring_kernels.cuh"
