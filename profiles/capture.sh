# Run under gpurun from the repo root (1 GPU).  Writes raw ncu outputs to gpurun_out/; profiles/summarize.py
# turns them into the tracked summaries.  Each ncu run is preceded by the same command without ncu (&&).
set -x
C3="python bench.py --steps 2 --warmup 3 --md-steps 10 --no-cpu-baseline --no-extra"
C2="python bench.py --workload c2 --steps 2 --warmup 3 --md-steps 20 --no-cpu-baseline"
$C3 > gpurun_out/plain_c3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c3.csv $C3 > gpurun_out/ncu_l_c3.log 2>&1
$C2 > gpurun_out/plain_c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c2.csv $C2 > gpurun_out/ncu_l_c2.log 2>&1
F3="python profiles/ab_cutoff.py fp32 20"
$F3 > gpurun_out/plain_f3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_nbr_force|k_nbr_build|k_cut_kick|k_bin_' -s 10 -c 8 -o gpurun_out/prof_c3 $F3 > gpurun_out/ncu_f3.log 2>&1
F3d="python profiles/ab_cutoff.py fp64 20"
$F3d > gpurun_out/plain_f3d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_nbr_force' -s 4 -c 1 -o gpurun_out/prof_c3_f64 $F3d > gpurun_out/ncu_f3d.log 2>&1
# rebuild kernels (list build + binning) after 200 warm-up steps, and the all-pairs kernels of configs[1]
RB="python profiles/ab_cutoff.py fp32 20 c3 200"
$RB > gpurun_out/plain_rb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_nbr_build|k_bin_' -s 6 -c 4 -o gpurun_out/prof_c3_rebuild $RB > gpurun_out/ncu_rb.log 2>&1
C2s="python bench.py --workload c2 --steps 2 --warmup 3 --md-steps 20 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_allpairs' -s 20 -c 1 -o gpurun_out/prof_c2_n3 $C2s > gpurun_out/ncu_c2n3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_allpairs' -s 20 -c 1 -o gpurun_out/prof_c2_f32 $C2s --precision fp32 > gpurun_out/ncu_c2f32.log 2>&1
ls -la gpurun_out | tail -12
# then, here:  python profiles/summarize.py launches gpurun_out/launches_c3.csv profiles/r01_launches_c3.txt   (same for c2)
#              python profiles/summarize.py full gpurun_out/prof_c3.ncu-rep profiles/r01_ncu_c3_kernels.txt
#              ... prof_c3_f64 -> r01_ncu_c3_force_f64.txt, prof_c3_rebuild -> r01_ncu_c3_rebuild.txt,
#              prof_c2_n3 -> r01_ncu_c2_allpairs_f64.txt, prof_c2_f32 -> r01_ncu_c2_allpairs_f32.txt
