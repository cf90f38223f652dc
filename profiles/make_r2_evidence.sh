#!/bin/bash
# Round-2 evidence, run on the GPU box (gpurun -- bash profiles/make_r2_evidence.sh): bench lines, ncu launch list of the bench
# command, `ncu --set full` captures of the fixed-size fused kernels and of the any-size kernel (two window sizes), any-size probe.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/bench_r2_b200.json 2> $O/bench_r2_b200.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_r2_reference.json 2>> $O/bench_r2_b200.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kcf_fused -s 2 -c 2 -o $O/r2_fused -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-loop > $O/ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kcf_any -s 3 -c 2 -o $O/r2_any_30x40 -f python profiles/probe_any_ncu.py 120x160 1184 > $O/ncu_any1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kcf_any -s 3 -c 2 -o $O/r2_any_25x15 -f python profiles/probe_any_ncu.py 100x60 2368 > $O/ncu_any2.log 2>&1
python profiles/probe_anysize.py > $O/r2_probe_anysize.jsonl 2> $O/probe.err
ls -la $O | tail -20
