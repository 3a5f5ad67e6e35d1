# r02, final tile force kernel (window copy with cp.async): one ncu --set full capture in the melted state + hot lines
set -x
mkdir -p gpurun_out /tmp/ncu
RB="python profiles/ab_cutoff.py fp32 20 c3 200"
timeout 200 $RB > gpurun_out/r02b_plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tile_force' -s 205 -c 2 -o /tmp/ncu/r02b_force $RB > gpurun_out/r02b_ncu.log 2>&1
python profiles/summarize.py full /tmp/ncu/r02b_force.ncu-rep gpurun_out/r02b_ncu_tile_force.txt > /dev/null
python profiles/hot_lines.py /tmp/ncu/r02b_force.ncu-rep k_tile_force mdsea_b200/_lib/obj/cells.o 30 > gpurun_out/r02b_hot_lines_tile_force.txt 2>&1
cat gpurun_out/r02b_plain.log
