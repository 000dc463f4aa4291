"""Condense `ncu --page raw --csv` exports (one kernel each) into one markdown table.
usage: ncu_summary.py TAG=FILE_raw.csv [TAG=FILE_raw.csv ...]"""
import csv, sys

KEYS = [
    ("time ms", "gpu__time_duration.sum", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("block", "launch__block_size", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("smem/blk KB", "launch__shared_mem_per_block_dynamic", 1.0),
    ("occ % ach", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("issue %", "smsp__issue_active.avg.pct", 1.0),
    ("fp64 pipe %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    ("tensor(dmma) %", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("lsu wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("dram rd", "dram__bytes_read.sum", 1.0),
    ("dram wr", "dram__bytes_write.sum", 1.0),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("local ld", "smsp__inst_executed_op_local_ld.sum", 1.0),
    ("local st", "smsp__inst_executed_op_local_st.sum", 1.0),
]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    cols = []
    for arg in sys.argv[1:]:
        tag, path = arg.split("=", 1)
        cols.append((tag, load(path)))
    print("| metric | " + " | ".join(t for t, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    print("| kernel | " + " | ".join(d.get("Kernel Name", ("?", ""))[0][:48].replace("|", "/") for _, d in cols) + " |")
    for name, key, _ in KEYS:
        cells = []
        for _, d in cols:
            v, u = d.get(key, ("-", ""))
            cells.append(f"{v} {u}".strip())
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
