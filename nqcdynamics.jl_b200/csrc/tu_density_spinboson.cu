// tu_density_spinboson.cu -- FSSH / Ehrenfest kernels for SpinBoson: the shared-memory-resident kernel of
// kernel_spinboson.cuh (default) and the generic lanes-over-modes kernels (init kernel, baths beyond the shared-memory budget).
#include <cstdlib>

#include "kernel_density.cuh"
#include "kernel_spinboson.cuh"

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
