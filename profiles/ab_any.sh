for cfg in "1024 1 0" "512 1 1" "512 2 0" "256 2 1" "256 4 0" "384 1 1" "256 1 1"; do
  set -- $cfg
  echo "== threads $1 ctas $2 r128 $3"
  MOT_ANY_THREADS=$1 MOT_ANY_CTAS=$2 MOT_ANY_R128=$3 python profiles/probe_anysize.py 100x60 120x160 52x36 80x48 2>&1 | tail -4
done
