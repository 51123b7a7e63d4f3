// Translation unit of aff_stripe_kernel<K, G, BT> (stripe_kernels.cuh).
#define POYB200_DEFINE_AFF_STRIPE
#include "launch.h"
#include "stripe_kernels.cuh"
