#!/bin/bash
# scripts/gpu_multi.sh <N> [parts...] — multi-GPU measurements of round 2 on N GPUs of one box (run under gpurun --gpus N)
#   tests : pytest tests/test_gpu_dist.py (the 2-GPU NCCL exchange test)
#   fine  : bench.py --gpus N (default: colour-grid exchange started inside backward) and with ESR_ALLREDUCE_OVERLAP=0,
#           each REPS times
#   lts   : bench.py --stage lts --gpus N (touched-block exchange) and --dense-allreduce
#   eval  : bench.py --stage eval --gpus N with and without the gather
cd "$(dirname "$0")/.." || exit 1
N=$1; shift
PARTS=${*:-tests fine lts eval}
REPS=${REPS:-1}
O=gpurun_out/multi_n$N
mkdir -p "$O"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29600
run() {
  local name=$1 t=$2; shift 2
  port=$((port + 1))
  timeout "$t" "$@" > "$O/$name.json" 2> "$O/$name.err"
  echo "$name rc=$?" | tee -a "$O/summary.txt"
}
for part in $PARTS; do
  case $part in
    tests) timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > "$O/pytest_dist.log" 2>&1; echo "pytest_dist rc=$?" | tee -a "$O/summary.txt"; tail -3 "$O/pytest_dist.log" ;;
    fine) for i in $(seq 1 "$REPS"); do
            run fine_default_$i 400 $TR --master-port $((port + 1)) bench.py --gpus "$N" --steps 20 --no-cpu-baseline
            ESR_ALLREDUCE_OVERLAP=0 run fine_nooverlap_$i 400 $TR --master-port $((port + 1)) bench.py --gpus "$N" --steps 20 --no-cpu-baseline
          done ;;
    lts) run lts_blocks 500 $TR --master-port $((port + 1)) bench.py --gpus "$N" --stage lts --steps 10
         run lts_dense 500 $TR --master-port $((port + 1)) bench.py --gpus "$N" --stage lts --steps 10 --dense-allreduce ;;
    eval) run eval_gather 400 $TR --master-port $((port + 1)) bench.py --gpus "$N" --stage eval --steps 5
          run eval_nogather 400 $TR --master-port $((port + 1)) bench.py --gpus "$N" --stage eval --steps 5 --no-gather ;;
  esac
done
python scripts/bench_summary.py "$O"/*.json 2>/dev/null | grep -v "^    " 
for f in "$O"/*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "exchange_check", d.get("exchange_check"), (d.get("exchange_check_detail") or {}).get("worst_rel_to_max"),
          "parallelism", d["config"].get("parallelism"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -n 3 "$O"/*.err | tail -n 40
