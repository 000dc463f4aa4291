// tu_iesh.cu -- AdiabaticIESH kernels (CTA per trajectory) and their shared-memory / tile plan.
#include "kernel_iesh.cuh"

namespace nq {
namespace {

// Tile / shared-memory plan for (n states, ne electrons); false when nothing fits.
int pad16_4(int x) { int y = x; while (y % 16 != 4) ++y; return y; }

bool iesh_plan(int n, int ne, size_t smem_max, IeshLayout& L) {
    L = IeshLayout{};
    L.threads = 384;
    const int nwarps = L.threads / 32;
    L.nrt = (n + 7) / 8;
    L.rounds = L.nrt / nwarps;                        // full rows per warp; the remaining rows are dealt out tile by tile
    if (L.rounds > 2) return false;
    const int rem_rows = L.nrt - L.rounds * nwarps;
    L.ldg = pad16_4(8 * L.nrt);
    const int n4 = (n + 3) & ~3;
    const int nt_need = (ne + 3) / 4;                 // column tiles for all electrons
    int nt_max = (L.rounds <= 1) ? 14 : 8;            // accumulator tiles per warp (register budget)
    if (rem_rows > 0) nt_max = std::min(nt_max, (2 * nwarps) / rem_rows);   // at most 2 extra tiles per warp
    if (nt_max < 1) return false;
    L.lds = ne | 1;
    const int nep = (ne + 3) & ~3;
    const long small = iesh_small_doubles(n);
    // hop phase: Gauss-Jordan fallback (S, S^-1 work vectors) or the LU buffers (+ S itself when ne > 64)
    const long lu = 8L * nep + (ne > 64 ? 2L * ne * L.lds : 0L);
    const long hop = std::max(2L * ne * L.lds + 7L * ne + n + (2 * ne + 1) / 2 + 4, lu);
    int lr = 1;
    while (lr < 32 && (long)n * lr * 2 <= L.threads) lr *= 2;
    L.lr = lr;
    // (a) G resident in shared memory
    {
        const int nct = std::min(nt_need, nt_max);
        const int ldb = pad16_4(8 * nct);
        const long work = (long)L.ldg * n4 + std::max((long)n4 * ldb, hop);
        if ((size_t)(small + work) * 8 <= smem_max) {
            L.resident = 1; L.nct = nct; L.ldb = ldb; L.kb = n4; L.nslab = 1;
            L.off_b = L.ldg * n4; L.off_hop = L.off_b; L.work_doubles = (int)work;
        }
    }
    // (b) G streamed from global memory (L2) in slabs of kb columns, double buffered
    if (!L.resident) {
        L.kb = 16;
        L.nslab = (n4 + L.kb - 1) / L.kb;
        const long slabs = 2L * L.ldg * L.kb;
        int nct = std::min(nt_need, nt_max);
        while (nct >= 1) {
            const long work = std::max(slabs + (long)n4 * pad16_4(8 * nct), hop);
            if ((size_t)(small + work) * 8 <= smem_max) break;
            --nct;
        }
        if (nct < 1) return false;
        L.nct = nct; L.ldb = pad16_4(8 * nct); L.off_b = (int)slabs; L.off_hop = 0;
        L.work_doubles = (int)std::max(slabs + (long)n4 * L.ldb, hop);
    }
    L.nchunks = (nt_need + L.nct - 1) / L.nct;
    L.nct = (nt_need + L.nchunks - 1) / L.nchunks;      // balance the chunks
    L.ldb = pad16_4(8 * L.nct);
    L.smem_bytes = (int)((small + L.work_doubles) * 8);
    return true;
}

}  // namespace

bool select_iesh(const nqcb200_config& c, KernelSet& out, std::string& why) {
    if (c.model != NQCB200_MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK && c.model != NQCB200_MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS) {
        why = "AdiabaticIESH is built for the AndersonHolstein (Newns-Anderson) model"; return false;
    }
    if (c.ndofs != 1) { why = "AdiabaticIESH kernel: ndofs == 1"; return false; }
    if (c.nbeads < 1 || c.nbeads > 32) { why = "AdiabaticIESH / EhrenfestNA ring polymers: 1 <= nbeads <= 32"; return false; }
    if (c.nbeads > 1 && c.edc_C > 0.0) { why = "EDC decoherence is built for nbeads == 1"; return false; }
    const int n = c.nstates, ne = c.nelectrons;
    if (ne > 112) { why = "AdiabaticIESH kernel: at most 112 electrons"; return false; }
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
    if (c.nbeads > 1) { out.step = iesh_step_kernel<true>; out.init = iesh_init_kernel<true>; }       // RPIESH / RP-EhrenfestNA (BCBWavefunction)
    else { out.step = iesh_step_kernel<false>; out.init = iesh_init_kernel<false>; }
    out.L = 1; out.DPL = 1;
    out.block = L.threads;
    out.dyn_smem = (size_t)L.smem_bytes;
    out.cta_per_trajectory = true;
    out.iesh = L;
    out.name = (c.nbeads > 1) ? "rpiesh_anderson_holstein" : "iesh_anderson_holstein";
    return true;
}
}  // namespace nq
