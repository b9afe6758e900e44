// Tensor-product kinds with lmax_filter = 3, lmax_out = 3 (see conv_dispatch.cuh).
#include "conv_dispatch.cuh"
S7B_DEFINE_CONV_GROUP(3, 3)
