"""Aggregate tools/instmix_job.sh captures into profiles/rNN/executed_flops.json (read by bench.py).
usage: executed_flops.py OUT.json WORKLOAD=CSV:T:NSTEPS:KERNEL_REGEX [...]"""
import collections, csv, json, re, sys

out = {}
for spec in sys.argv[2:]:
    wl, rest = spec.split("=", 1)
    path, T, ns, rx = rest.split(":", 3)
    T, ns = int(T), int(ns)
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    count = collections.Counter()
    seen = set()
    for r in csv.reader(open(path)):
        if len(r) < 15 or not r[0].isdigit():
            continue
        name = re.sub(r"\(.*", "", r[4])
        per[name][r[12]] += float(r[14].replace(",", ""))
        if (r[0], name) not in seen:
            seen.add((r[0], name)); count[name] += 1
    sel = [k for k in per if re.search(rx, k)]
    tot = collections.defaultdict(float)
    for k in sel:
        for m, v in per[k].items():
            tot[m] += v
    units = float(T) * ns
    flops = (2 * tot["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + tot["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
             + tot["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]) / units
    out[wl] = {"flops_per_traj_step": flops,
               "dram_bytes_per_traj_step": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / units,
               "kernels": {k: {"launches": count[k], "ms": per[k]["gpu__time_duration.sum"] / 1e6,
                               "fp64_pipe_pct": per[k]["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"] / max(1, count[k]),
                               "regs": per[k]["launch__registers_per_thread"] / max(1, count[k]),
                               "local_ld": per[k]["smsp__inst_executed_op_local_ld.sum"], "local_st": per[k]["smsp__inst_executed_op_local_st.sum"]}
                           for k in sel},
               "source": f"profiles/{path.split('/')[-1]} ({T} trajectories x {ns} steps, kernels /{rx}/)"}
    print(wl, "flops/traj-step %.1f" % flops, "dram B/traj-step %.2f" % out[wl]["dram_bytes_per_traj_step"],
          {k: round(v["ms"], 3) for k, v in out[wl]["kernels"].items()})
json.dump(out, open(sys.argv[1], "w"), indent=1)
