ncu --set full --import-source on --clock-control none --warp-sampling-interval 1 -k regex:td_frame -s 20 -c 1 -o gpurun_out/prof_tdframe_c2 -f python profiles/probe_c2_launches.py > gpurun_out/tdf_ncu.log 2>&1
tail -3 gpurun_out/tdf_ncu.log
