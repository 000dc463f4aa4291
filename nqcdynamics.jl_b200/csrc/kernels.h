// kernels.h -- host-visible kernel table.  Each tu_*.cu instantiates one kernel family and exports a
// selector; engine.cu only sees function pointers, so the families compile in parallel.
#pragma once
#include <string>

#include "common.cuh"

namespace nq {

using StepFn = void (*)(const KParams);
using InitFn = void (*)(const KParams, int, int, const double*);

struct KernelSet {
    StepFn step = nullptr;
    InitFn init = nullptr;
    StepFn step_term = nullptr;   // TerminatingCallback instantiation of `step` (nqcb200_set_termination), if any
    bool step_term_step_shape = false;   // step_term launches with the STEP kernel's shape (step_block / step_smem) instead of the init kernel's
    int term_L = 0, term_block = 0;       // ... or with its own shape when the step kernel's differs (warp-specialised ring-polymer variant)
    size_t term_smem = 0;
    int L = 1;              // lanes (threads) per trajectory
    int DPL = 1;            // nuclear dofs per lane
    int block = kBlockThreads;
    size_t dyn_smem = 0;
    // launch shape of the STEP kernel when it differs from the init kernel's (0 = same as above)
    int step_L = 0;
    int step_block = 0;
    size_t step_smem = 0;
    bool needs_sb_carry = false;   // kernel_spinboson.cuh: two force scalars per trajectory carried between launches
    bool fused_init = false;       // the step kernel can initialise from KParams.r_aos / v_aos (see common.cuh)
    // kernel_spinboson_epoch.cuh: E nuclear steps per (bath pass, electronic kernel) pair instead of one `step` kernel
    int sb_epoch = 0;
    StepFn sb_init = nullptr, sb_prep = nullptr, sb_bath = nullptr, sb_elec = nullptr;
    bool cta_per_trajectory = false;
    IeshLayout iesh = {};   // AdiabaticIESH tile / shared-memory plan (kernel_iesh.cuh)
    const char* name = "";
};

constexpr int kObsReplicas = 16;  // accumulator copies, folded after each launch (atomic contention)

bool select_density_1d(const nqcb200_config& c, KernelSet& out, std::string& why);
bool select_density_spinboson(const nqcb200_config& c, KernelSet& out, std::string& why);
bool select_ring_density(const nqcb200_config& c, KernelSet& out, std::string& why);
bool select_classical(const nqcb200_config& c, KernelSet& out, std::string& why);
bool select_nrpmd(const nqcb200_config& c, KernelSet& out, std::string& why);
bool select_langevin(const nqcb200_config& c, KernelSet& out, std::string& why);
bool select_iesh(const nqcb200_config& c, KernelSet& out, std::string& why);

}  // namespace nq
