# same-box A/B of two builds of the library: bash profiles/ab_lib.sh [rounds]; A = libmot_base.so, B = libmot_b200.so
for i in $(seq ${1:-2}); do
  for L in libmot_base.so libmot_b200.so; do
    MOT_B200_LIB=$PWD/multiple-object-tracking_b200/$L python bench.py --no-e2e --no-cpu --no-loop --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$L', round(d['value']), d['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('kernel_ms'))"
  done
done
