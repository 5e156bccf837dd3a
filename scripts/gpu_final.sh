#!/bin/bash
# scripts/gpu_final.sh <part> — the round-2 closing measurements (one gpurun call per part; results in gpurun_out/final/)
#   single : pytest -m gpu, smoke(), the three N = 1 bench lines (fine = the contract's default command, lts, eval) + the reference arm
#   ncu    : launch list + --set full capture of exactly the timed steps of the default command (not bench values)
#   multi  : on the GPUs of the box (gpurun --gpus N): the 2-GPU NCCL exchange test, fine weak + strong (2^16 global rays), lts
cd "$(dirname "$0")/.." || exit 1
PART=${1:-single}
O=gpurun_out/final
mkdir -p "$O"
NGPU=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $NGPU"
run() {  # run <name> <timeout_s> <command...>
  local name=$1 t=$2; shift 2
  timeout "$t" "$@" > "$O/$name.json" 2> "$O/$name.err"
  echo "$name rc=$?" | tee -a "$O/summary_$PART.txt"
}
case $PART in
  single)
    timeout 600 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$O/summary_$PART.txt"; tail -2 "$O/pytest_gpu.log"
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$O/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$O/summary_$PART.txt"; tail -2 "$O/smoke.log"
    run fine_n1 400 python bench.py --steps 20
    run ref_n1 400 python bench.py --impl reference --steps 2 --warmup 1
    run lts_n1 300 python bench.py --stage lts --steps 10 --no-cpu-baseline
    run eval_n1 300 python bench.py --stage eval --steps 5 --no-cpu-baseline
    run fine_n1_bf16 300 python bench.py --steps 10 --mlp-mode bf16 --no-cpu-baseline ;;
  ncu)
    B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profiler-range"
    run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file "$O/r02b_launches.csv" $B
    run ncu_full 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o "$O/r02b_full" -f $B ;;
  multi)
    [ "$NGPU" -eq 2 ] && { timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > "$O/pytest_dist.log" 2>&1; echo "pytest_dist rc=$?" | tee -a "$O/summary_$PART.txt"; tail -2 "$O/pytest_dist.log"; }
    [ "$NGPU" -eq 2 ] && run fine_weak_n$NGPU 400 $TR --master-port 29611 bench.py --gpus "$NGPU" --steps 20 --no-cpu-baseline
    run fine_strong_n$NGPU 400 $TR --master-port 29612 bench.py --gpus "$NGPU" --steps 20 --no-cpu-baseline --global-rays 65536
    run lts_n$NGPU 500 $TR --master-port 29613 bench.py --gpus "$NGPU" --stage lts --steps 10 --no-cpu-baseline ;;
esac
python scripts/bench_summary.py "$O"/*.json 2>/dev/null | grep -v "^    "
