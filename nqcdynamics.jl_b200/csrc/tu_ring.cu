// tu_ring.cu -- RPSH / RP-Ehrenfest: dispatch to the per-model translation units (tu_ring_<model>.cu, ring_select.cuh).
#include <string>

#include "kernels.h"

namespace nq {
bool select_ring_tully1(const nqcb200_config& c, KernelSet& out);
bool select_ring_tully2(const nqcb200_config& c, KernelSet& out);
bool select_ring_doublewell(const nqcb200_config& c, KernelSet& out);
bool select_ring_morse3(const nqcb200_config& c, KernelSet& out);

bool select_ring_density(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.ndofs != 1) { why = "ring-polymer FSSH/Ehrenfest kernels are instantiated for ndofs == 1"; return false; }
    bool ok = false;
    switch (c.model) {
        case NQCB200_MODEL_TULLY_ONE: ok = select_ring_tully1(c, out); break;
        case NQCB200_MODEL_TULLY_TWO: ok = select_ring_tully2(c, out); break;
        case NQCB200_MODEL_DOUBLE_WELL: ok = select_ring_doublewell(c, out); break;
        case NQCB200_MODEL_THREE_STATE_MORSE: ok = select_ring_morse3(c, out); break;
        default: break;
    }
    if (!ok) why = "ring-polymer kernel: unsupported model, or the beads do not fit in shared memory";
    return ok;
}
}  // namespace nq
