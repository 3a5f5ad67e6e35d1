python profiles/probe_8m_per_gpu.py 200 40 2>&1 | tail -3 | tee gpurun_out/probe_8m.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
