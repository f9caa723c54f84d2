#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full captures of the blend kernels.
# Every step has its own timeout so one stall cannot eat the visit.  Usage: bash scripts/gpu_round.sh [tag] [what...]
TAG=${1:-r01}; shift
WHAT=${@:-tests smoke bench ncu ref}   # also: benchstaged benchfull trace experimental multi2 multi4 multi8
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
multi[248])
  N=${w#multi}   # needs gpurun --gpus N
  T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
  timeout 200 $T 29533 scripts/check_exchange_multi.py 2>&1 | grep -E "step|ok|rror" | tail -5
  timeout 300 $T 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${TAG}.json 2> gpurun_out/bench_${N}gpu_${TAG}.err
  grep "train (res" gpurun_out/bench_${N}gpu_${TAG}.err | tail -1
  timeout 300 $T 29518 bench.py --gpus $N --steps 10 --warmup 3 --trace gpurun_out/trace_${N}gpu_${TAG}.json 2>&1 | grep "us x" | head -14 ;;
experimental)
  # kernel variants of the backward blend: parity, then the in-step time of each
  timeout 300 python -m pytest tests/test_frame_gpu.py -m gpu -q -p no:cacheprovider -k "variants or loss" --tb=short 2>&1 | tail -5
  for v in ${VARIANTS:-"SPV_BWD_VARIANT=0" "SPV_BWD_VARIANT=1"}; do
    env $v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --quick 2> gpurun_out/bench_variant_${TAG}.err | python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value'], 1), 'it/s', d['kernels_in_step_ms'])"
  done ;;
ncusmall)
  timeout 500 ncu --set full --clock-control none --import-source on -k "regex:${NCU_SMALL:-cull_count_emit|tile_sort_small|tile_scan|pack_records|deform_fwd2|frame_geometry|sh_fwd|sh_bwd|unpack_frame|select_pass|rgb_row|track_fused|adam_lazy}" -s ${NCU_SKIP:-30} -c ${NCU_COUNT:-24} -f -o gpurun_out/prof_small_${TAG} \
      python bench.py --steps 2 --warmup 1 --profile-mode --no-graph > gpurun_out/ncu_small_${TAG}.log 2>&1
  tail -3 gpurun_out/ncu_small_${TAG}.log ;;
ncufwd)
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend_rec_fwd -s 2 -c 2 -f -o gpurun_out/prof_fwd_${TAG} \
      python bench.py --steps 2 --warmup 1 --profile-mode > gpurun_out/ncu_fwd_${TAG}.log 2>&1
  tail -3 gpurun_out/ncu_fwd_${TAG}.log ;;
ncubwd)
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend_rec_bwd -s 2 -c 2 -f -o gpurun_out/prof_bwd_${TAG} \
      python bench.py --steps 2 --warmup 1 --profile-mode > gpurun_out/ncu_bwd_${TAG}.log 2>&1
  tail -3 gpurun_out/ncu_bwd_${TAG}.log ;;
bigcfg)
  # BASELINE configs[2] / configs[3] on ONE GPU (dry run of the workloads the 4- and 8-GPU lines carry as extra_configs)
  for c in cfg3_480p_500k cfg4_1080p_2m; do
    timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --quick 2> gpurun_out/bench_${c}_${TAG}.err | tee gpurun_out/bench_${c}_${TAG}.json | cut -c1-600
    tail -2 gpurun_out/bench_${c}_${TAG}.err
  done ;;
trace)
  timeout 200 python bench.py --steps 10 --warmup 3 --trace gpurun_out/trace_1gpu_${TAG}.json 2>&1 | grep "us x" | head -30 ;;
esac
done
ls -la gpurun_out | tail -25
