// tu_ring_tully1.cu -- ring-polymer FSSH / Ehrenfest kernels for one model (see ring_select.cuh).
#include "ring_select.cuh"

namespace nq {
bool select_ring_tully1(const nqcb200_config& c, KernelSet& out) {
    t_device = c.device;
    return pick_beads<ModelT<NQCB200_MODEL_TULLY_ONE>>(c.method, c.nbeads, c.ntraj, out, "rp_tully1");
}
}  // namespace nq
