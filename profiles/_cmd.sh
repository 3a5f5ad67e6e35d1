set -x
timeout 600 python -m pytest tests/test_gpu_cutoff.py tests/test_gpu_multirank.py -x -q 2>&1 | tail -5
MDSEA_BINNING=radix timeout 600 python -m pytest tests/test_gpu_cutoff.py tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
A="python profiles/ab_cutoff.py fp32 20"
$A > gpurun_out/plain_ab.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_count.csv $A > gpurun_out/ncu_ab.log 2>&1
cat gpurun_out/plain_ab.log
