// tu_density_spinboson.cu -- FSSH / Ehrenfest kernels for SpinBoson: the shared-memory-resident kernel of
// kernel_spinboson.cuh (default) and the generic lanes-over-modes kernels (init kernel, baths beyond the shared-memory budget).
#include <cstdlib>

#include "kernel_density.cuh"
#include "kernel_spinboson.cuh"
#include "kernel_spinboson_epoch.cuh"

namespace nq {
namespace {
template <int DPL, int L>
bool pick(int method, KernelSet& out) {
    using M = ModelT<NQCB200_MODEL_SPIN_BOSON>;
    if (method == NQCB200_METHOD_FSSH) {
        out.step = density_step_kernel<M, DPL, L, NQCB200_METHOD_FSSH>;
        out.init = density_init_kernel<M, DPL, L, NQCB200_METHOD_FSSH>;
        out.name = "spinboson_fssh";
    } else if (method == NQCB200_METHOD_EHRENFEST) {
        out.step = density_step_kernel<M, DPL, L, NQCB200_METHOD_EHRENFEST>;
        out.init = density_init_kernel<M, DPL, L, NQCB200_METHOD_EHRENFEST>;
        out.name = "spinboson_ehrenfest";
    } else return false;
    out.L = L; out.DPL = DPL;
    return true;
}
}  // namespace

bool select_density_spinboson(const nqcb200_config& c, KernelSet& out, std::string& why) {
    const int D = c.ndofs, m = c.method;
    if (c.nbath != D) { why = "SpinBoson needs nbath == ndofs"; return false; }
    int lanes = 0;
    if (const char* env = getenv("NQCB200_SPINBOSON_LANES")) lanes = atoi(env);
    // default: thread per trajectory with the bath resident in shared memory (kernel_spinboson.cuh); the lane-split
    // generic kernels below remain for D beyond the shared-memory budget (and as an A/B switch through the env var)
    if (lanes == 0 && D >= 2 && sb_smem_bytes(D) <= 220 * 1024 && (m == NQCB200_METHOD_FSSH || m == NQCB200_METHOD_EHRENFEST)) {
        bool ok = (D <= 8) ? pick<8, 1>(m, out) : (D <= 104) ? pick<13, 8>(m, out) : pick<4, 32>(m, out);   // init kernel
        if (!ok) return false;
        // two lanes per trajectory (two warps per scheduler) once the sweep is long enough to split
        int lpt = (D >= 16) ? 2 : 1;
        if (const char* env = getenv("NQCB200_SPINBOSON_LPT")) { const int v = atoi(env); if (v == 1 || v == 2) lpt = v; }
        if (lpt == 2) out.step = (m == NQCB200_METHOD_FSSH) ? spinboson_step_kernel<NQCB200_METHOD_FSSH, 2>
                                                            : spinboson_step_kernel<NQCB200_METHOD_EHRENFEST, 2>;
        else out.step = (m == NQCB200_METHOD_FSSH) ? spinboson_step_kernel<NQCB200_METHOD_FSSH, 1>
                                                   : spinboson_step_kernel<NQCB200_METHOD_EHRENFEST, 1>;
        out.step_L = lpt; out.step_block = kSbTraj * lpt; out.step_smem = sb_smem_bytes(D);
        out.fused_init = true;
        out.needs_sb_carry = true;
        out.name = (m == NQCB200_METHOD_FSSH) ? "spinboson_fssh_tpt" : "spinboson_ehrenfest_tpt";
        // E steps per (bath pass, electronic kernel) pair (kernel_spinboson_epoch.cuh) when no output needs bath
        // coordinates at the save points.  NQCB200_SPINBOSON_EPOCH=0 keeps the step-by-step kernel, =8 selects the
        // shorter epoch (A/B switches, documented in DESIGN.md section 3).
        const uint32_t electronic_only = (1u << NQCB200_OBS_ADIABATIC_POP) | (1u << NQCB200_OBS_DIABATIC_POP) |
                                         (1u << NQCB200_OBS_POPCORR_DIABATIC) | (1u << NQCB200_OBS_POPCORR_ADIABATIC) |
                                         (1u << NQCB200_OBS_DISCRETE_STATE) | (1u << NQCB200_OBS_SIGMA);
        int epoch = 16;
        if (const char* env = getenv("NQCB200_SPINBOSON_EPOCH")) epoch = atoi(env);
        if (epoch != 0 && D >= 8 && !c.diagnostics && (c.observables & ~electronic_only) == 0) {
            const bool f = (m == NQCB200_METHOD_FSSH);
#define NQ_SB_EPOCH(E)                                                                                                   \
            do {                                                                                                             \
                out.sb_epoch = E;                                                                                            \
                out.sb_prep = f ? sb_prep_kernel<NQCB200_METHOD_FSSH, E> : sb_prep_kernel<NQCB200_METHOD_EHRENFEST, E>;      \
                out.sb_bath = sb_bath_kernel<E>;                                                                             \
                out.sb_elec = f ? sb_elec_kernel<NQCB200_METHOD_FSSH, E> : sb_elec_kernel<NQCB200_METHOD_EHRENFEST, E>;      \
            } while (0)
            if (epoch == 8) NQ_SB_EPOCH(8); else NQ_SB_EPOCH(16);
#undef NQ_SB_EPOCH
            out.sb_init = f ? sb_init_kernel<NQCB200_METHOD_FSSH> : sb_init_kernel<NQCB200_METHOD_EHRENFEST>;
            out.fused_init = true;       // nqcb200_run_from_host: chunked upload overlapped with the epochs of the previous chunk
            out.name = f ? "spinboson_fssh_epoch" : "spinboson_ehrenfest_epoch";
        }
        return true;
    }
    if (D <= 4 && (lanes == 0 || lanes == 1)) return pick<4, 1>(m, out);
    if (D <= 8 && (lanes == 0 || lanes == 1)) return pick<8, 1>(m, out);
    if (D <= 104 && (lanes == 0 || lanes == 8)) return pick<13, 8>(m, out);
    if (D <= 112 && lanes == 16) return pick<7, 16>(m, out);
    if (D <= 128) return pick<4, 32>(m, out);
    if (D <= 512) return pick<16, 32>(m, out);
    why = "SpinBoson kernels cover ndofs <= 512";
    return false;
}
}  // namespace nq
