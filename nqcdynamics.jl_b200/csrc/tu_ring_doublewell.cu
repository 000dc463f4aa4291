// tu_ring_doublewell.cu -- ring-polymer FSSH / Ehrenfest kernels for one model (see ring_select.cuh).
#include "ring_select.cuh"

namespace nq {
bool select_ring_doublewell(const nqcb200_config& c, KernelSet& out) {
    t_device = c.device;
    return pick_beads<ModelT<NQCB200_MODEL_DOUBLE_WELL>>(c.method, c.nbeads, c.ntraj, out, "rp_doublewell");
}
}  // namespace nq
