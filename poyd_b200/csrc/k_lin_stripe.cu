// Translation unit of lin_stripe_kernel<K, G, BT> (lin_stripe_kernels.cuh).
#define POYB200_DEFINE_LIN_STRIPE
#include "launch.h"
#include "lin_stripe_kernels.cuh"
