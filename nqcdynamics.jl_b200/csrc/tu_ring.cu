// tu_ring.cu -- RPSH / RP-Ehrenfest kernels (beads on lanes).
#include "kernel_ring.cuh"

namespace nq {
namespace {
template <class M, int NB>
bool pick(int method, KernelSet& out, const char* name) {
    if (method == NQCB200_METHOD_FSSH) {
        out.step = ring_step_kernel<M, NB, NQCB200_METHOD_FSSH>;
        out.init = ring_init_kernel<M, NB, NQCB200_METHOD_FSSH>;
    } else if (method == NQCB200_METHOD_EHRENFEST) {
        out.step = ring_step_kernel<M, NB, NQCB200_METHOD_EHRENFEST>;
        out.init = ring_init_kernel<M, NB, NQCB200_METHOD_EHRENFEST>;
    } else return false;
    out.L = NB; out.DPL = 1; out.name = name;
    return true;
}
template <class M>
bool pick_beads(int method, int B, KernelSet& out, const char* name) {
    switch (B) {
        case 2: return pick<M, 2>(method, out, name);
        case 4: return pick<M, 4>(method, out, name);
        case 8: return pick<M, 8>(method, out, name);
        case 16: return pick<M, 16>(method, out, name);
        case 32: return pick<M, 32>(method, out, name);
    }
    return false;
}
}  // namespace

bool select_ring_density(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.ndofs != 1) { why = "ring-polymer FSSH/Ehrenfest kernels are instantiated for ndofs == 1"; return false; }
    bool ok = false;
    switch (c.model) {
        case NQCB200_MODEL_TULLY_ONE: ok = pick_beads<ModelT<NQCB200_MODEL_TULLY_ONE>>(c.method, c.nbeads, out, "rp_tully1"); break;
        case NQCB200_MODEL_TULLY_TWO: ok = pick_beads<ModelT<NQCB200_MODEL_TULLY_TWO>>(c.method, c.nbeads, out, "rp_tully2"); break;
        case NQCB200_MODEL_DOUBLE_WELL: ok = pick_beads<ModelT<NQCB200_MODEL_DOUBLE_WELL>>(c.method, c.nbeads, out, "rp_doublewell"); break;
        case NQCB200_MODEL_THREE_STATE_MORSE: ok = pick_beads<ModelT<NQCB200_MODEL_THREE_STATE_MORSE>>(c.method, c.nbeads, out, "rp_morse3"); break;
        default: break;
    }
    if (!ok) why = "ring-polymer kernel: unsupported model or nbeads not in {2,4,8,16,32}";
    return ok;
}
}  // namespace nq
