// explicit instantiations of the slice kernels for dim_head = 32
#include "slice_v2.cuh"
namespace tbns {
TBNS_SLICE_INSTANTIATE_D(32)
}
