set -x
python -m pytest tests/test_gpu_kalman_assoc.py tests/test_gpu_tdloop.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_launches_new.csv python profiles/probe_c3_launches.py > gpurun_out/c3_new.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches_new.csv python profiles/probe_c2_launches.py > gpurun_out/c2_new.log 2>&1
python bench.py --config C2 2>/dev/null | tail -1 | cut -c1-1500
python bench.py --config C3 2>/dev/null | tail -1 | cut -c1-1200
