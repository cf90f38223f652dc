# config 5 with the number of concurrently resident Munkres CTAs capped (working copies L2-resident or not)
for c in 0 24 32 48 64 96; do echo "== MOT_MUNKRES_CTAS=$c"; MOT_MUNKRES_CTAS=$c python bench.py --config C5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())['details']; print({k: round(d[k]['gpu_matrices_per_s'],1) for k in ('ref_centroid','iou_clamped')})"; done
