A="python profiles/ab_cutoff.py fp32 20"
$A > gpurun_out/plain_ab.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_nbr_build' -s 1 -c 1 -o gpurun_out/prof_b5 $A > gpurun_out/ncu_b5.log 2>&1
