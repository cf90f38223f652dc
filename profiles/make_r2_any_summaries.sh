#!/bin/bash
# `ncu --set full` of the any-size kernel at two window sizes, summarised on the box (headline metrics, stall split, per-phase and per-line shares)
O=gpurun_out
for sz in 120x160:1184 100x60:2368; do
  px=${sz%%:*}; n=${sz##*:}
  ncu --set full --clock-control none --import-source on -k regex:kcf_any -s 3 -c 2 -o /tmp/r2_any_$px -f python profiles/probe_any_ncu.py $px $n > $O/ncu_any_$px.log 2>&1
  { ncu -i /tmp/r2_any_$px.ncu-rep --page raw --csv | python profiles/raw_headline.py; python profiles/any_profile.py /tmp/r2_any_$px.ncu-rep 0 25 512; python profiles/any_profile.py /tmp/r2_any_$px.ncu-rep 1 10 512; } > $O/r2_kcf_any_${px}_ncu_summary.txt 2>&1
done
