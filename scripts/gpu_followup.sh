#!/bin/bash
# scripts/gpu_followup.sh — ONE gpurun call that measures everything round 1 wrote after its GPU budget was spent
# (DESIGN.md §7 "Next, ranked"):
#
#   /usr/local/graft/bin/gpurun           --timeout 900  -- 'bash scripts/gpu_followup.sh single'   # ~10 GPU-minutes
#   /usr/local/graft/bin/gpurun           --timeout 1500 -- 'bash scripts/gpu_followup.sh ncu'      # ~6 GPU-minutes
#   /usr/local/graft/bin/gpurun --gpus 2  --timeout 1200 -- 'bash scripts/gpu_followup.sh pair'     # ~15 min x 2 GPUs
#   /usr/local/graft/bin/gpurun --gpus 4  --timeout 900  -- 'bash scripts/gpu_followup.sh quad'     # ~12 min x 4 GPUs
#
# (box time is charged per GPU: run the parts separately; with no argument every part the box has GPUs for runs.)
#
#  1. default build: pytest -m gpu + the bench line (the committed MLP chains never ran on a GPU in their present form:
#     a mechanical revert of the tile overlap on top of GPU-verified epilogue trims)
#  2. ESR_MLP_TILE_OVERLAP=1 (the safe tile overlap, own mbarrier for the layer-0 commit): MLP / fine / LTS parity, bench
#  3. esr_grad_block_flags against its torch restatement; N = 2: fine bench; LTS bench dense vs --block-exchange
#     (dist.TouchedBlockCompactor, gloo-checked only)
#  4. N = 4: the hang of round 1's last call — three runs each without / with ESR_ALLREDUCE_OVERLAP=1
# Every step runs under its own `timeout`, so a wedged kernel costs minutes, not the call; results land in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/followup
mkdir -p "$O"
NGPU=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # run <name> <timeout_s> <command...>: stdout -> $O/name.json|log, stderr -> $O/name.err, exit code recorded
  local name=$1 t=$2; shift 2
  timeout "$t" "$@" > "$O/$name.out" 2> "$O/$name.err"
  echo "$name rc=$?" | tee -a "$O/summary.txt"
}

PART=${1:-all}
want() { { [ "$PART" = all ] && [ "$1" != ncu ]; } || [ "$PART" = "$1" ]; }   # (ncu only when asked for)

if want single; then
run pytest_default 900 python -m pytest tests -m gpu -x -q
run bench_default 400 python bench.py --steps 20
ESR_TEST_UNVERIFIED=1 run pytest_block_flags 300 python -m pytest tests/test_gpu_native_ops.py -m gpu -x -q -k block_flags
ESR_TEST_UNVERIFIED=1 run pytest_fib_envmaps 300 python -m pytest tests/test_gpu_esrnerf.py -m gpu -x -q -k other_samplers
ESR_MLP_TILE_OVERLAP=1 run pytest_tile_overlap 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_voxurff.py tests/test_gpu_esrnerf.py -m gpu -x -q
ESR_MLP_TILE_OVERLAP=1 run bench_tile_overlap 400 python bench.py --steps 20 --no-cpu-baseline
ESR_MLP_TILE_OVERLAP=1 run bench_tile_overlap_eval 400 python bench.py --stage eval --steps 5
fi

if want ncu; then   # ~6 GPU-minutes: launch list + full capture of exactly the timed steps (numbers printed here are not bench values)
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profiler-range"
run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file "$O/r02_launches.csv" $B
run ncu_full 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o "$O/r02_full" -f $B
# then, here: python scripts/ncu_summary.py launches gpurun_out/followup/r02_launches.csv profiles/r02_launches_summary.csv "..."
#             python scripts/ncu_summary.py full gpurun_out/followup/r02_full.ncu-rep profiles/r02_top_kernels_ncu_full.json "..."
fi

if want pair && [ "$NGPU" -ge 2 ]; then
  run pytest_dist 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q
  run bench_n2 400 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2
  run bench_lts_n2_dense 500 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --stage lts
  run bench_lts_n2_blocks 500 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --stage lts --block-exchange
  run bench_fine_n2_blocks 400 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --block-exchange
  run bench_eval_n2_gather 400 $TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --stage eval --steps 5
  run bench_eval_n2_nogather 400 $TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --stage eval --steps 5 --no-gather
fi
if want quad && [ "$NGPU" -ge 4 ]; then
  for i in 1 2 3; do
    run bench_n4_$i 400 $TR --nproc-per-node 4 --master-port $((29520 + i)) bench.py --gpus 4
    ESR_ALLREDUCE_OVERLAP=1 run bench_n4_overlap_$i 400 $TR --nproc-per-node 4 --master-port $((29530 + i)) bench.py --gpus 4
  done
fi
cat "$O/summary.txt"
