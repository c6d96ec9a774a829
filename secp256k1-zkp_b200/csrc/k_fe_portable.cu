// k_fe_portable.cu -- the same kernel through the portable (no inline PTX) limb code, for the
// on-device A/B self test of the PTX carry chains.  Built with -DB200_PORTABLE_BN -DB200_NS=b200_portable.
#ifndef B200_PORTABLE_BN
#error "compile with -DB200_PORTABLE_BN -DB200_NS=b200_portable"
#endif
#include "k_fe_impl.cuh"
namespace b200 {
cudaError_t launch_fe_op_portable(int op, uint8_t* out, const uint8_t* a, const uint8_t* b, size_t n, int iters, cudaStream_t s, uint64_t* launches) {
    return b200_portable::launch_fe_op_impl(op, out, a, b, n, iters, s, launches);
}
}
