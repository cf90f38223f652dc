#!/bin/bash
# `ncu --set full` of the association-side kernels of round 2: the Munkres solver on a 256x256 tracking-shaped problem (config 3's frame)
# and the one-launch Kalman frame kernel (config 2's frame); the small reports travel back, the summaries are made with
# profiles/line_samples.py (stall samples per source line) on the build that was profiled.
O=gpurun_out
ncu --set full --import-source on --clock-control none --warp-sampling-interval 1 -k regex:munkres -s 3 -c 1 -o $O/prof_munkres_c3 -f python profiles/probe_c3_launches.py > $O/munk_ncu.log 2>&1
ncu --set full --import-source on --clock-control none --warp-sampling-interval 1 -k regex:td_frame -s 20 -c 1 -o $O/prof_tdframe_c2 -f python profiles/probe_c2_launches.py > $O/tdf_ncu.log 2>&1
ls -la $O/prof_munkres_c3.ncu-rep $O/prof_tdframe_c2.ncu-rep
