// tu_density_1d.cu -- FSSH / Ehrenfest kernels for the one-dimensional models (thread per trajectory).
#include "kernel_density.cuh"

namespace nq {
namespace {
template <class M>
bool pick(int method, KernelSet& out, const char* name) {
    if (method == NQCB200_METHOD_FSSH) {
        out.step = density_step_kernel<M, 1, 1, NQCB200_METHOD_FSSH>;
        out.step_term = density_step_kernel<M, 1, 1, NQCB200_METHOD_FSSH, true>;
        out.init = density_init_kernel<M, 1, 1, NQCB200_METHOD_FSSH>;
    } else if (method == NQCB200_METHOD_EHRENFEST) {
        out.step = density_step_kernel<M, 1, 1, NQCB200_METHOD_EHRENFEST>;
        out.step_term = density_step_kernel<M, 1, 1, NQCB200_METHOD_EHRENFEST, true>;
        out.init = density_init_kernel<M, 1, 1, NQCB200_METHOD_EHRENFEST>;
    } else return false;
    out.L = 1; out.DPL = 1; out.name = name;
    return true;
}
}  // namespace

bool select_density_1d(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.ndofs != 1) { why = "this model's FSSH/Ehrenfest kernels are instantiated for ndofs == 1"; return false; }
    switch (c.model) {
        case NQCB200_MODEL_TULLY_ONE: return pick<ModelT<NQCB200_MODEL_TULLY_ONE>>(c.method, out, "tully1");
        case NQCB200_MODEL_TULLY_TWO: return pick<ModelT<NQCB200_MODEL_TULLY_TWO>>(c.method, out, "tully2");
        case NQCB200_MODEL_TULLY_THREE: return pick<ModelT<NQCB200_MODEL_TULLY_THREE>>(c.method, out, "tully3");
        case NQCB200_MODEL_DOUBLE_WELL: return pick<ModelT<NQCB200_MODEL_DOUBLE_WELL>>(c.method, out, "doublewell");
        case NQCB200_MODEL_THREE_STATE_MORSE: return pick<ModelT<NQCB200_MODEL_THREE_STATE_MORSE>>(c.method, out, "morse3");
        default: break;
    }
    why = "no FSSH/Ehrenfest kernel for this model";
    return false;
}
}  // namespace nq
