// tu_ring_tully2.cu -- ring-polymer FSSH / Ehrenfest kernels for one model (see ring_select.cuh).
#include "ring_select.cuh"

namespace nq {
bool select_ring_tully2(const nqcb200_config& c, KernelSet& out) {
    t_device = c.device;
    return pick_beads<ModelT<NQCB200_MODEL_TULLY_TWO>>(c.method, c.nbeads, c.ntraj, out, "rp_tully2");
}
}  // namespace nq
