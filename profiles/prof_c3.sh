# ncu captures for the cut-off path (run under gpurun from the repo root)
set -e
python profiles/prof_force.py c3 fp32 3 > gpurun_out/plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_nbr_force|k_nbr_build' -c 4 -o gpurun_out/prof_c3 python profiles/prof_force.py c3 fp32 3 > gpurun_out/ncu_c3.log 2>&1
