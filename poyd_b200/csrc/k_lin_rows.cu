// Translation unit of lin_rows_kernel<C, G, BT> (lin_rows_kernels.cuh).
#define POYB200_DEFINE_LIN_ROWS
#include "launch.h"
#include "lin_rows_kernels.cuh"
