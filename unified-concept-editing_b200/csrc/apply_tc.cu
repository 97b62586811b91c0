// tcgen05 3xTF32 implementation of the low-rank apply (placeholder until the kernel lands).
#include "uce_ws.h"
namespace uce {
bool apply_tc_available(const uce_ws*) { return false; }
int apply_tc_lowrank(uce_ws*, const LayerRef*, const LayerRef*, int, int, cudaStream_t, int*) {
    set_error("tcgen05 apply not built");
    return UCE_E_STATE;
}
}  // namespace uce
