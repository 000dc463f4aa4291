// ring_select.cuh -- kernel selection for the ring-polymer FSSH / Ehrenfest family, shared by the per-model translation
// units tu_ring_<model>.cu (one model each, so that `make -j` compiles the 200-odd instantiations in parallel: a single
// translation unit took 7 minutes).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "kernel_ring.cuh"
#include "kernel_ring_tpt.cuh"

namespace nq {
namespace {

thread_local int t_device = 0;   // device of the handle being created (select_ring_density)
// block size of ring_tpt_step_kernel (0: the beads of even one warp do not fit in shared memory)
int tpt_threads(int N, int NB, bool ehr, bool fft, int64_t ntraj) {
    int max_threads = 0;
    for (int b = 32; b <= kRpshMaxThreads; b += 32)
        if (ring_tpt_smem_bytes(N, NB, ehr, fft, b) <= 200 * 1024) max_threads = b;
    if (max_threads == 0) return 0;
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t_device) != cudaSuccess || sms <= 0) sms = 148;
    return ring_tpt_block_threads(ntraj > 0 ? ntraj : 1, sms, max_threads);
}
template <class M, int NB>
bool pick(int method, int64_t ntraj, KernelSet& out, const char* name) {
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
    // Shards smaller than one wave of threads (strong scaling: BASELINE config 5 is 12 500 trajectories per GPU on 8 GPUs):
    // warp-specialised phases, LPT members per trajectory (kernel_ring_tpt.cuh).  NQCB200_RING_LPT=1|2|4 overrides (A/B).
    if constexpr (NB >= 8) {
        if (want && !ehr) {
            int sms = 148;
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t_device) != cudaSuccess || sms <= 0) sms = 148;
            const char* force = getenv("NQCB200_RING_LPT");
            const int64_t per_sm = (std::max<int64_t>(ntraj, 1) + sms - 1) / sms;
            const int lpt = force ? atoi(force) : (per_sm <= 96 ? 4 : (per_sm <= 192 ? 2 : 1));
            int ks = 0;
            if (lpt == 4 || lpt == 2) {
                ks = (int)std::min<int64_t>(kRpshMaxThreads / lpt, 32 * ((per_sm + 31) / 32));     // owners fill whole warps
                while (ks > 32 && ring_tpt_smem_bytes(M::NS, NB, false, true, ks, true) > 200 * 1024) ks -= 32;   // 32 beads x 192 owners do not fit
                if (ring_tpt_smem_bytes(M::NS, NB, false, true, ks, true) > 200 * 1024) ks = 0;
            }
            if (ks > 0) {
                if (lpt == 4) out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_FSSH, false, 4>;
                else out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_FSSH, false, 2>;
                out.step_L = lpt; out.step_block = ks * lpt;
                out.step_smem = ring_tpt_smem_bytes(M::NS, NB, false, true, ks, true);
                // TerminatingCallback: the thread-per-trajectory TERM instantiation with its own launch shape
                const int tthreads = tpt_threads(M::NS, NB, false, true, ntraj);
                if (tthreads > 0) {
                    out.step_term = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_FSSH, true>;
                    out.term_L = 1; out.term_block = tthreads; out.term_smem = ring_tpt_smem_bytes(M::NS, NB, false, true, tthreads);
                    out.step_term_step_shape = true;
                }
                return true;
            }
        }
    }
    const int threads = tpt_threads(M::NS, NB, ehr, true, ntraj);
    if (want && threads > 0) {
        if (ehr) { out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_EHRENFEST>; out.step_term = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_EHRENFEST, true>; }
        else { out.step = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_FSSH>; out.step_term = ring_tpt_step_kernel<M, NB, NQCB200_METHOD_FSSH, true>; }
        out.step_term_step_shape = true;
        out.step_L = 1; out.step_block = threads; out.step_smem = ring_tpt_smem_bytes(M::NS, NB, ehr, true, threads);
    }
    return true;
}
// any other nbeads: thread-per-trajectory init + step with the dense normal-mode product
template <class M>
bool pick_generic(int method, int B, int64_t ntraj, KernelSet& out, const char* name) {
    const bool ehr = (method == NQCB200_METHOD_EHRENFEST);
    if (method != NQCB200_METHOD_FSSH && !ehr) return false;
    const int threads = tpt_threads(M::NS, B, ehr, false, ntraj);
    if (B < 2 || threads == 0) return false;
    if (ehr) { out.step = ring_tpt_step_kernel<M, 0, NQCB200_METHOD_EHRENFEST>; out.step_term = ring_tpt_step_kernel<M, 0, NQCB200_METHOD_EHRENFEST, true>; out.init = ring_tpt_init_kernel<M, NQCB200_METHOD_EHRENFEST>; }
    else { out.step = ring_tpt_step_kernel<M, 0, NQCB200_METHOD_FSSH>; out.step_term = ring_tpt_step_kernel<M, 0, NQCB200_METHOD_FSSH, true>; out.init = ring_tpt_init_kernel<M, NQCB200_METHOD_FSSH>; }
    out.step_term_step_shape = true;
    out.L = 1; out.DPL = 1; out.name = name;
    out.step_L = 1; out.step_block = threads; out.step_smem = ring_tpt_smem_bytes(M::NS, B, ehr, false, threads);
    return true;
}
template <class M>
bool pick_beads(int method, int B, int64_t ntraj, KernelSet& out, const char* name) {
    switch (B) {
        case 2: return pick<M, 2>(method, ntraj, out, name);
        case 4: return pick<M, 4>(method, ntraj, out, name);
        case 8: return pick<M, 8>(method, ntraj, out, name);
        case 16: return pick<M, 16>(method, ntraj, out, name);
        case 32: return pick<M, 32>(method, ntraj, out, name);
    }
    return pick_generic<M>(method, B, ntraj, out, name);
}
}  // namespace
}  // namespace nq
