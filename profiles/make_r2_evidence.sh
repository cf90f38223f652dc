#!/bin/bash
# Round-2 evidence, run on the GPU box (gpurun -- bash profiles/make_r2_evidence.sh): bench lines, ncu launch list of the bench
# command, `ncu --set full` captures of the fixed-size fused kernels and of the any-size kernel (two window sizes) summarised on
# the box (the reports themselves are too large to carry back), any-size probe, config lines, launch lists of the single-stream
# frame loops, SASS instruction counts.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/bench_r2_b200.json 2> $O/bench_r2_b200.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_r2_reference.json 2>> $O/bench_r2_b200.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kcf_fused -s 2 -c 2 -o /tmp/r2_fused -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-loop > $O/ncu_fused.log 2>&1
{ python profiles/analyze.py /tmp/r2_fused.ncu-rep; python profiles/phase_detail.py /tmp/r2_fused.ncu-rep 1; python profiles/phase_detail.py /tmp/r2_fused.ncu-rep 0; } > $O/r2_kcf_fused_ncu_summary.txt 2>&1
if [ "$1" != "quick" ]; then
for sz in 120x160:1184 100x60:2368; do
  px=${sz%%:*}; n=${sz##*:}
  ncu --set full --clock-control none --import-source on -k regex:kcf_any -s 3 -c 2 -o /tmp/r2_any_$px -f python profiles/probe_any_ncu.py $px $n > $O/ncu_any_$px.log 2>&1
  { ncu -i /tmp/r2_any_$px.ncu-rep --page raw --csv | python profiles/raw_headline.py; python profiles/any_profile.py /tmp/r2_any_$px.ncu-rep 0 25 512; python profiles/any_profile.py /tmp/r2_any_$px.ncu-rep 1 12 512; } > $O/r2_kcf_any_${px}_ncu_summary.txt 2>&1
done
python profiles/probe_anysize.py > $O/r2_probe_anysize.jsonl 2> $O/probe.err
fi
for c in C1 C2 C3 C5; do python bench.py --config $c 2>> $O/r2_bench_configs.err | tail -1; done > $O/r2_bench_configs.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_c3_frame.csv python profiles/probe_c3_launches.py > $O/c3_new.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_c2_frame.csv python profiles/probe_c2_launches.py > $O/c2_new.log 2>&1
cuobjdump -sass multiple-object-tracking_b200/libmot_b200.so | python profiles/sass_counts.py > $O/r2_sass_counts.txt
ls -la $O | tail -20
