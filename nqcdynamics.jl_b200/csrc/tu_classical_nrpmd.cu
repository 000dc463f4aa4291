// tu_classical_nrpmd.cu -- classical MD / RPMD and NRPMD kernels (beads on lanes).
#include <cstdlib>

#include "kernel_nrpmd.cuh"
#include "kernel_ring_tpt.cuh"

namespace nq {
namespace {
template <class M, int NB>
void set_classical(KernelSet& k, const char* name) {
    k.step = classical_ring_step_kernel<M, NB>;
    k.init = classical_ring_init_kernel<M, NB>;
    k.L = NB; k.DPL = 1; k.name = name;
    if constexpr (NB >= 2) {
        // step kernel: the whole ring polymer in one thread's registers (kernel_ring_tpt.cuh); NQCB200_RPMD_TPT=0
        // keeps the beads-on-lanes kernel (A/B switch)
        const char* env = getenv("NQCB200_RPMD_TPT");
        if (!(env && atoi(env) == 0)) {
            k.step = classical_tpt_fft_kernel<M, NB>;
            k.step_L = 1; k.step_block = kRtThreads; k.step_smem = 0;
        }
    }
}
template <class M>
void k_generic(KernelSet& k, const char* name, size_t bytes) {
    k.step = classical_tpt_step_kernel<M>;
    k.init = classical_tpt_init_kernel<M>;
    k.L = 1; k.DPL = 1; k.name = name;
    k.step_L = 1; k.step_block = kRtThreads; k.step_smem = bytes;
}
template <class M>
bool pick_classical(int B, KernelSet& out, const char* name) {
    switch (B) {
        case 1: set_classical<M, 1>(out, name); return true;
        case 2: set_classical<M, 2>(out, name); return true;
        case 4: set_classical<M, 4>(out, name); return true;
        case 8: set_classical<M, 8>(out, name); return true;
        case 16: set_classical<M, 16>(out, name); return true;
        case 32: set_classical<M, 32>(out, name); return true;
    }
    // any other nbeads: thread per trajectory, dense normal-mode product (kernel_ring_tpt.cuh)
    const size_t bytes = ((size_t)5 * B * kRtThreads + (size_t)B * B + 4 * B) * sizeof(double);
    if (B < 2 || bytes > 200 * 1024) return false;
    k_generic<M>(out, name, bytes);
    return true;
}
template <class M, int NB>
void set_nrpmd(KernelSet& k, const char* name) {
    k.step = nrpmd_step_kernel<M, NB>;
    k.init = nrpmd_init_kernel<M, NB>;
    k.L = NB; k.DPL = 1; k.name = name;
}
template <class M>
bool pick_nrpmd(int B, KernelSet& out, const char* name) {
    switch (B) {
        case 1: set_nrpmd<M, 1>(out, name); return true;
        case 2: set_nrpmd<M, 2>(out, name); return true;
        case 4: set_nrpmd<M, 4>(out, name); return true;
        case 8: set_nrpmd<M, 8>(out, name); return true;
        case 16: set_nrpmd<M, 16>(out, name); return true;
        case 32: set_nrpmd<M, 32>(out, name); return true;
    }
    return false;
}
}  // namespace

bool select_classical(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.ndofs != 1) { why = "classical/RPMD kernels are instantiated for ndofs == 1"; return false; }
    bool ok = false;
    if (c.model == NQCB200_MODEL_HARMONIC) ok = pick_classical<ModelT<NQCB200_MODEL_HARMONIC>>(c.nbeads, out, "rpmd_harmonic");
    else if (c.model == NQCB200_MODEL_FREE) ok = pick_classical<ModelT<NQCB200_MODEL_FREE>>(c.nbeads, out, "rpmd_free");
    if (!ok) why = "classical method needs a classical model (and beads that fit in shared memory)";
    return ok;
}

namespace {
template <class M, int NB>
void set_langevin(KernelSet& k, const char* name) {
    k.step = langevin_tpt_fft_kernel<M, NB>;
    k.init = classical_ring_init_kernel<M, NB>;
    k.L = NB; k.DPL = 1; k.name = name;
    k.step_L = 1; k.step_block = kRtThreads; k.step_smem = 0;
}
template <class M>
bool pick_langevin(int B, KernelSet& out, const char* name) {
    switch (B) {
        case 2: set_langevin<M, 2>(out, name); return true;
        case 4: set_langevin<M, 4>(out, name); return true;
        case 8: set_langevin<M, 8>(out, name); return true;
        case 16: set_langevin<M, 16>(out, name); return true;
        case 32: set_langevin<M, 32>(out, name); return true;
    }
    // any other bead count >= 2: thread per trajectory, dense normal-mode product
    const size_t bytes = ((size_t)5 * B * kRtThreads + (size_t)B * B + 4 * B) * sizeof(double);
    if (B < 2 || bytes > 200 * 1024) return false;
    out.step = classical_tpt_step_kernel<M, true>;
    out.init = classical_tpt_init_kernel<M>;
    out.L = 1; out.DPL = 1; out.name = name;
    out.step_L = 1; out.step_block = kRtThreads; out.step_smem = bytes;
    return true;
}
}  // namespace

bool select_langevin(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.ndofs != 1) { why = "ThermalLangevin kernels are instantiated for ndofs == 1"; return false; }
    bool ok = false;
    if (c.model == NQCB200_MODEL_HARMONIC) ok = pick_langevin<ModelT<NQCB200_MODEL_HARMONIC>>(c.nbeads, out, "langevin_harmonic");
    else if (c.model == NQCB200_MODEL_FREE) ok = pick_langevin<ModelT<NQCB200_MODEL_FREE>>(c.nbeads, out, "langevin_free");
    if (!ok) why = "ThermalLangevin (BCOCB) needs a classical model and nbeads >= 2 (beads must fit in shared memory)";
    return ok;
}

bool select_nrpmd(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.ndofs != 1) { why = "NRPMD kernels are instantiated for ndofs == 1"; return false; }
    bool ok = false;
    switch (c.model) {
        case NQCB200_MODEL_TULLY_ONE: ok = pick_nrpmd<ModelT<NQCB200_MODEL_TULLY_ONE>>(c.nbeads, out, "nrpmd_tully1"); break;
        case NQCB200_MODEL_TULLY_TWO: ok = pick_nrpmd<ModelT<NQCB200_MODEL_TULLY_TWO>>(c.nbeads, out, "nrpmd_tully2"); break;
        case NQCB200_MODEL_DOUBLE_WELL: ok = pick_nrpmd<ModelT<NQCB200_MODEL_DOUBLE_WELL>>(c.nbeads, out, "nrpmd_doublewell"); break;
        case NQCB200_MODEL_THREE_STATE_MORSE: ok = pick_nrpmd<ModelT<NQCB200_MODEL_THREE_STATE_MORSE>>(c.nbeads, out, "nrpmd_morse3"); break;
        default: break;
    }
    if (!ok) why = "NRPMD kernel: unsupported model or nbeads not in {1,2,4,8,16,32}";
    return ok;
}
}  // namespace nq
