for cfg in "16 128 512" "14 128 512" "13 128 512" "12 128 512" "14 64 512" "14 256 512" "14 128 256" "14 128 128" "13 128 128" "13 64 64" "15 128 128" "14 256 256"; do
  set -- $cfg
  for a in "16 uniform 20" "20 uniform 10" "24 uniform 3"; do
    H2B_MSM_REDUCE_BLOCK_MAX_LOG=$1 H2B_MSM_REDUCE_BLOCK=$2 H2B_MSM_REDUCE_TAIL=$3 python tools/msm_tune.py $a | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(json.dumps({'cfg': '$cfg', 'k': d['k'], 'ms': d['ms'], 'reduce': d['phases']['7']}))"
  done
done
