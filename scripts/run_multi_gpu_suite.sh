#!/bin/bash
# The multi-GPU evidence of a round in one lease (gpurun --gpus 8): parity tests at 4 and 8 ranks on both transports and
# through the single-process front end, then the strong-scaling points.  Every step has its own timeout and stray GPU
# processes are killed (by exact PID) between steps, so that one hang cannot pollute or starve the rest.
# usage: bash scripts/run_multi_gpu_suite.sh <tag>
tag=${1:-multi}
out=gpurun_out
mkdir -p $out
cleanup() {
  for p in $(nvidia-smi --query-compute-apps=pid --format=csv,noheader 2>/dev/null | sort -u); do kill -9 "$p" 2>/dev/null; done
  sleep 1
}
run_bench() {  # name, ngpu, timeout, extra args...
  local name=$1 n=$2 t=$3; shift 3
  timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus "$n" "$@" > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  echo "$name rc=$?"
  cleanup
}
T="tests/test_gpu_multi.py"
timeout 560 python -m pytest -q \
  "$T::test_slab_solve_multi_gpu[4-23,8,6-1]" "$T::test_slab_solve_multi_gpu[8-40,6,6-1]" "$T::test_slab_solve_multi_gpu[8-40,6,6-0]" \
  "$T::test_slab_solve_multi_gpu[8-9,6,5-1]" "$T::test_single_process_multi_gpu_regular[4-23,8,6]" \
  "$T::test_single_process_multi_gpu_regular[8-40,6,6]" "$T::test_single_process_multi_gpu_regular[8-9,6,5]" \
  "$T::test_single_process_multi_gpu_irregular[4]" "$T::test_c_abi_multi_demo[8]" > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $out/${tag}_pytest.log | cut -c1-250
cleanup
if grep -q " passed" $out/${tag}_pytest.log && ! grep -q "failed" $out/${tag}_pytest.log; then export FVB_BENCH_ALT=1; echo "distributed multigrid leg enabled"; fi
run_bench n8_512 8 400 --steps 3 --warmup 3 --no-cpu-baseline
run_bench n8_1024 8 400 --grid 1024 --implicit --steps 1 --warmup 1 --skip-e2e --no-extras --no-cpu-baseline --parity-n 0
CUDA_VISIBLE_DEVICES=0,1,2,3 run_bench n4_1024 4 500 --grid 1024 --implicit --steps 1 --warmup 1 --skip-e2e --no-extras --no-cpu-baseline --parity-n 0
unset FVB_BENCH_ALT
FVB_FUSED_HALO_OFF=1 run_bench n8_512_unfused 8 300 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --parity-n 0 --skip-e2e
