python -m pytest tests/test_gpu_tdloop.py tests/test_gpu_kalman_assoc.py -x -q -m gpu 2>&1 | tail -3
python bench.py --config C2 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['details'],indent=0)[:1500])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches_new.csv python profiles/probe_c2_launches.py > gpurun_out/c2_new.log 2>&1
tail -3 gpurun_out/c2_launches_new.csv
