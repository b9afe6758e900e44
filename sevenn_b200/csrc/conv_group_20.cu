// Tensor-product kinds with lmax_filter = 2, lmax_out = 0 (see conv_dispatch.cuh).
#include "conv_dispatch.cuh"
S7B_DEFINE_CONV_GROUP(2, 0)
