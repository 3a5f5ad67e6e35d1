# r02 captures.  Run under gpurun from the repo root (1 GPU).  The .ncu-rep files stay in /tmp on the box (together they exceed
# what gpurun copies back); the text summaries (profiles/summarize.py, profiles/hot_lines.py) are written to gpurun_out/ there
# and copied into profiles/r02/ here.  Every ncu run is preceded by the same command without ncu.
set -x
mkdir -p gpurun_out /tmp/ncu
C3="python bench.py --steps 2 --warmup 3 --md-steps 10 --no-cpu-baseline --no-extra --no-strong --no-parity"
timeout 300 $C3 > gpurun_out/r02_plain_c3.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_c3.csv $C3 > gpurun_out/r02_ncu_l_c3.log 2>&1
C2="python bench.py --workload c2 --steps 2 --warmup 3 --md-steps 20 --no-cpu-baseline --no-parity"
timeout 300 $C2 > gpurun_out/r02_plain_c2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_c2.csv $C2 > gpurun_out/r02_ncu_l_c2.log 2>&1
# tile kernels + finish + binning in the MELTED state (200 warm-up steps = ~520 matching launches are skipped)
RB="python profiles/ab_cutoff.py fp32 20 c3 200"
timeout 300 $RB > gpurun_out/r02_plain_rb.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tile_force|k_tile_build|k_cut_kick|k_bin_' -s 560 -c 24 -o /tmp/ncu/r02_prof_c3 $RB > gpurun_out/r02_ncu_rb.log 2>&1
python profiles/summarize.py full /tmp/ncu/r02_prof_c3.ncu-rep gpurun_out/r02_ncu_c3_kernels.txt > /dev/null
python profiles/hot_lines.py /tmp/ncu/r02_prof_c3.ncu-rep k_tile_build mdsea_b200/_lib/obj/cells.o 45 > gpurun_out/r02_hot_lines_tile_build.txt 2>&1
python profiles/hot_lines.py /tmp/ncu/r02_prof_c3.ncu-rep k_tile_force mdsea_b200/_lib/obj/cells.o 45 > gpurun_out/r02_hot_lines_tile_force.txt 2>&1
# all-pairs kernels of configs[1], both precisions
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_allpairs' -s 20 -c 1 -o /tmp/ncu/r02_prof_c2_f64 $C2 > gpurun_out/r02_ncu_c2f64.log 2>&1
python profiles/summarize.py full /tmp/ncu/r02_prof_c2_f64.ncu-rep gpurun_out/r02_ncu_c2_allpairs_f64.txt > /dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_allpairs' -s 20 -c 1 -o /tmp/ncu/r02_prof_c2_f32 $C2 --precision fp32 > gpurun_out/r02_ncu_c2f32.log 2>&1
python profiles/summarize.py full /tmp/ncu/r02_prof_c2_f32.ncu-rep gpurun_out/r02_ncu_c2_allpairs_f32.txt > /dev/null
python profiles/hot_lines.py /tmp/ncu/r02_prof_c2_f64.ncu-rep k_allpairs_n3_f64 mdsea_b200/_lib/obj/allpairs.o 30 > gpurun_out/r02_hot_lines_allpairs_n3_f64.txt 2>&1
ls -la gpurun_out /tmp/ncu | tail -30
