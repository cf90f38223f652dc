ncu --set full --import-source on --clock-control none -k regex:munkres -s 3 -c 1 -o gpurun_out/prof_munkres_c3 -f python profiles/probe_c3_launches.py > gpurun_out/munk_ncu.log 2>&1
tail -3 gpurun_out/munk_ncu.log
