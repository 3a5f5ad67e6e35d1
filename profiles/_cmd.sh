set -x
timeout 600 python -m pytest tests/test_gpu_cutoff.py tests/test_gpu_multirank.py -x -q 2>&1 | tail -5
A="python profiles/ab_cutoff.py fp32 20"
$A > gpurun_out/plain_ab.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_notex.csv $A > gpurun_out/ncu_ab.log 2>&1
export MDSEA_FORCE_TEX=1
$A > gpurun_out/plain_ab2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_tex.csv $A > gpurun_out/ncu_ab2.log 2>&1
cat gpurun_out/plain_ab.log gpurun_out/plain_ab2.log
