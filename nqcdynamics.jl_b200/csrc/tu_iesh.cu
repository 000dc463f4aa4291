// tu_iesh.cu -- AdiabaticIESH kernels (CTA per trajectory) and their shared-memory / tile plan.
#include "kernel_iesh.cuh"

namespace nq {
namespace {

// Tile / shared-memory plan for (n states, ne electrons); false when nothing fits.
bool iesh_plan(int n, int ne, size_t smem_max, IeshLayout& L) {
    L = IeshLayout{};
    L.threads = 384;
    L.nrt = (n + 7) / 8;
    L.ldg = 8 * L.nrt;
    if (L.nrt > L.threads) return false;
    const int nct_need = (ne + 1) / 2;
    const int nct_cap = L.threads / L.nrt;
    L.lds = ne | 1;
    const long small = iesh_small_doubles(n);
    const long hop = 2L * ne * L.lds + 7L * ne + n + (2 * ne + 1) / 2 + 4;
    int lr = 1;
    while (lr < 32 && (long)n * lr * 2 <= L.threads) lr *= 2;
    L.lr = lr;
    // (a) G resident in shared memory
    {
        const int nct = std::min(nct_need, nct_cap);
        const long B = (long)L.ldg * 4 * nct;
        const long work = (long)L.ldg * L.ldg + std::max(B, hop);
        if ((size_t)(small + work) * 8 <= smem_max) {
            L.resident = 1; L.nct = nct; L.ldb = 4 * nct; L.kb = L.ldg; L.nslab = 1;
            L.off_b = L.ldg * L.ldg; L.off_hop = L.off_b; L.work_doubles = (int)work;
        }
    }
    // (b) G streamed from global memory (L2) in slabs of kb columns, double buffered
    if (!L.resident) {
        L.kb = 16;
        L.nslab = (n + L.kb - 1) / L.kb;
        const long slabs = 2L * L.ldg * L.kb;
        int nct = std::min(nct_need, nct_cap);
        while (nct >= 1) {
            const long work = std::max(slabs + (long)L.ldg * 4 * nct, hop);
            if ((size_t)(small + work) * 8 <= smem_max) break;
            --nct;
        }
        if (nct < 1) return false;
        L.nct = nct; L.ldb = 4 * nct; L.off_b = (int)slabs; L.off_hop = 0;
        L.work_doubles = (int)std::max(slabs + (long)L.ldg * 4 * nct, hop);
    }
    L.nchunks = (nct_need + L.nct - 1) / L.nct;
    L.smem_bytes = (int)((small + L.work_doubles) * 8);
    return true;
}

}  // namespace

bool select_iesh(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.model != NQCB200_MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK) {
        why = "AdiabaticIESH is built for the AndersonHolstein (Newns-Anderson) model"; return false;
    }
    if (c.ndofs != 1 || c.nbeads != 1) { why = "AdiabaticIESH kernel: ndofs == 1 and nbeads == 1"; return false; }
    const int n = c.nstates, ne = c.nelectrons;
    if (n < 3 || ne < 1 || ne >= n) { why = "AdiabaticIESH needs nstates >= 3 and 1 <= nelectrons < nstates"; return false; }
    if (c.nbath != n - 1 || !c.bath_a || !c.bath_b) { why = "AndersonHolstein needs nstates-1 bath energies and couplings"; return false; }
    for (int k = 0; k < n - 1; ++k) {
        // the secular-equation eigensolver relies on strict interlacing: distinct ascending bath energies, no zero coupling
        if (c.bath_b[k] == 0.0 || (k > 0 && !(c.bath_a[k] > c.bath_a[k - 1]))) {
            why = "AndersonHolstein bath must have strictly ascending energies and non-zero couplings"; return false;
        }
    }
    IeshLayout L;
    if (!iesh_plan(n, ne, 227 * 1024, L)) { why = "AdiabaticIESH: system too large for one CTA's shared memory"; return false; }
    out.step = iesh_step_kernel;
    out.init = iesh_init_kernel;
    out.L = 1; out.DPL = 1;
    out.block = L.threads;
    out.dyn_smem = (size_t)L.smem_bytes;
    out.cta_per_trajectory = true;
    out.iesh = L;
    out.name = "iesh_anderson_holstein";
    return true;
}
}  // namespace nq
