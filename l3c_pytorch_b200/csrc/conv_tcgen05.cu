// conv_tcgen05.cu -- tensor-core (tcgen05 + TMA + TMEM) implicit-GEMM convolution.  Placeholder
// until the kernel lands: the dispatcher fails loudly instead of silently falling back.
#include "common.cuh"

namespace l3c {
int conv2d_tcgen05(const l3c_conv_t &p, cudaStream_t) {
    set_error("l3c_conv2d: precision mode %d (tcgen05) is not built yet", p.precision);
    return L3C_EINVAL;
}
}  // namespace l3c
