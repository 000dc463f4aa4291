// tu_ring.cu -- RPSH / RP-Ehrenfest kernels (beads on lanes).
#include <cstdlib>

#include "kernel_ring.cuh"
#include "kernel_ring_tpt.cuh"

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
    // step kernel: one thread per trajectory with the beads in shared memory (kernel_ring_tpt.cuh) when they fit;
    // the beads-on-lanes kernel stays as the fallback (and as an A/B switch: NQCB200_RING_TPT=0)
    const char* env = getenv("NQCB200_RING_TPT");
    const bool want = !(env && atoi(env) == 0);
    const bool ehr = (method == NQCB200_METHOD_EHRENFEST);
    const size_t bytes = ring_tpt_smem_bytes(M::NS, NB, ehr, true);
    if (want && bytes <= 200 * 1024) {
        if (ehr) out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_EHRENFEST>;
        else out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_FSSH>;
        out.step_L = 1; out.step_block = kRtThreads; out.step_smem = bytes;
    }
    return true;
}
// any other nbeads: thread-per-trajectory init + step with the dense normal-mode product
template <class M>
bool pick_generic(int method, int B, KernelSet& out, const char* name) {
    const bool ehr = (method == NQCB200_METHOD_EHRENFEST);
    if (method != NQCB200_METHOD_FSSH && !ehr) return false;
    const size_t bytes = ring_tpt_smem_bytes(M::NS, B, ehr, false);
    if (B < 2 || bytes > 200 * 1024) return false;
    if (ehr) { out.step = ring_tpt_step_kernel<M, 0, NQCB200_METHOD_EHRENFEST>; out.init = ring_tpt_init_kernel<M, NQCB200_METHOD_EHRENFEST>; }
    else { out.step = ring_tpt_step_kernel<M, 0, NQCB200_METHOD_FSSH>; out.init = ring_tpt_init_kernel<M, NQCB200_METHOD_FSSH>; }
    out.L = 1; out.DPL = 1; out.name = name;
    out.step_L = 1; out.step_block = kRtThreads; out.step_smem = bytes;
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
    return pick_generic<M>(method, B, out, name);
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
    if (!ok) why = "ring-polymer kernel: unsupported model, or the beads do not fit in shared memory";
    return ok;
}
}  // namespace nq
