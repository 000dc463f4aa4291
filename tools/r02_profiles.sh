#!/bin/bash
# Round-2 ncu captures at HEAD (run under gpurun): --set full of the dominant kernel of each config + launch lists.
tools/profile.sh spinboson_debye100_fssh sb_bath 1000000 32 r02_sb_bath > /dev/null 2>&1
tools/profile.sh spinboson_debye100_fssh sb_elec 1000000 32 r02_sb_elec > /dev/null 2>&1
tools/profile.sh iesh_anderson_holstein_m100 iesh_step 1480 100 r02_iesh_m100 > /dev/null 2>&1
tools/profile.sh iesh_anderson_holstein_m200 iesh_step 592 30 r02_iesh_m200 > /dev/null 2>&1
tools/profile.sh rpsh_morse3_16 ring_tpt_step 113664 300 r02_rpsh > /dev/null 2>&1
tools/profile.sh tully1_fssh density_step 1048576 600 r02_tully1 > /dev/null 2>&1
tools/profile.sh nrpmd_morse3_16 nrpmd_step 100000 300 r02_nrpmd > /dev/null 2>&1
for t in sb_bath sb_elec iesh_m100 iesh_m200 rpsh tully1 nrpmd; do python tools/ncu_summary.py $t=gpurun_out/prof_r02_${t}_raw.csv | grep -E "kernel|time|regs|occ|fp64|tensor|lsu|dram %|local"; done
ls gpurun_out | wc -l
