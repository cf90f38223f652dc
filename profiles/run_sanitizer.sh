# compute-sanitizer over the round-2 kernels: memcheck on the association / frame-loop / golden tests, racecheck (shared-memory hazards:
# the warp-level phase hand-offs and named barriers of the fused KCF kernel, the one-launch frame kernel) on the golden + Kalman-loop tests
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_golden.py tests/test_gpu_kalman_assoc.py "tests/test_gpu_tdloop.py::test_device_resident_loop_trace" -x -q -m gpu > $O/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_golden.py "tests/test_gpu_tdloop.py::test_device_resident_loop_trace" -x -q -m gpu > $O/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/r2_sanitizer_racecheck.txt
tail -n 5 $O/r2_sanitizer_memcheck.txt; tail -n 5 $O/r2_sanitizer_racecheck.txt
