#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full captures of the blend kernels.
# Every step has its own timeout so one stall cannot eat the visit.  Usage: bash scripts/gpu_round.sh [tag] [what...]
TAG=${1:-r01}; shift
WHAT=${@:-tests smoke bench ncu ref}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/gpu_${TAG}.txt
echo "host cores: $(nproc)" >> gpurun_out/gpu_${TAG}.txt
for w in $WHAT; do
case $w in
tests)
  timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=20 --tb=short > gpurun_out/tests_${TAG}.log 2>&1
  tail -12 gpurun_out/tests_${TAG}.log ;;
smoke)
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
bench)
  timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  tail -c 2500 gpurun_out/bench_${TAG}.json; tail -12 gpurun_out/bench_${TAG}.err ;;
benchstaged)
  timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --staged > gpurun_out/bench_staged_${TAG}.json 2> gpurun_out/bench_staged_${TAG}.err
  tail -c 2500 gpurun_out/bench_staged_${TAG}.json; tail -4 gpurun_out/bench_staged_${TAG}.err ;;
benchfull)
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full_${TAG}.json 2> gpurun_out/bench_full_${TAG}.err
  tail -c 3000 gpurun_out/bench_full_${TAG}.json; tail -4 gpurun_out/bench_full_${TAG}.err ;;
ref)
  timeout 700 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
  tail -c 1500 gpurun_out/bench_ref_${TAG}.json; tail -5 gpurun_out/bench_ref_${TAG}.err ;;
ncu)
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --steps 2 --warmup 1 --profile-mode > gpurun_out/ncu_launches_${TAG}.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend_rec_bwd -s 2 -c 2 -f -o gpurun_out/prof_bwd_${TAG} \
      python bench.py --steps 2 --warmup 1 --profile-mode > gpurun_out/ncu_bwd_${TAG}.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend_rec_fwd -s 2 -c 2 -f -o gpurun_out/prof_fwd_${TAG} \
      python bench.py --steps 2 --warmup 1 --profile-mode > gpurun_out/ncu_fwd_${TAG}.log 2>&1
  tail -3 gpurun_out/ncu_bwd_${TAG}.log ;;
esac
done
ls -la gpurun_out | tail -25
