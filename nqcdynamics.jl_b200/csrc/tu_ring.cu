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
    if (want && method == NQCB200_METHOD_FSSH && ring_tpt_smem_bytes<M::NS, NB, NQCB200_METHOD_FSSH>() <= 200 * 1024) {
        out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_FSSH>;
        out.step_L = 1; out.step_block = kRtThreads; out.step_smem = ring_tpt_smem_bytes<M::NS, NB, NQCB200_METHOD_FSSH>();
    } else if (want && method == NQCB200_METHOD_EHRENFEST && ring_tpt_smem_bytes<M::NS, NB, NQCB200_METHOD_EHRENFEST>() <= 200 * 1024) {
        out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_EHRENFEST>;
        out.step_L = 1; out.step_block = kRtThreads; out.step_smem = ring_tpt_smem_bytes<M::NS, NB, NQCB200_METHOD_EHRENFEST>();
    }
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
