#!/bin/bash
# One gpurun call's worth of evidence for the current tree (run from the repo root on a B200 box):
#   parity suite, the default bench line, the ncu launch list of the same command, and one
#   `--set full` capture of the dominant search kernel per workload.  Outputs -> gpurun_out/<tag>_*.
# Usage: tools/gpu_round.sh <tag> [what...]   what in {tests bench launches full_cfg3 full_cfg2 ref exp}
tag=${1:-run}; shift
what=${*:-tests bench launches full_cfg3 full_cfg2}
mkdir -p gpurun_out
OURS='regex:pack_seed|seed_packed|count_kmers|constrain_ranges'
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for w in $what; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
      echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log ;;
    bench)
      timeout 1200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      echo "bench rc=$?"; head -c 600 gpurun_out/${tag}_bench.json; echo ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
      echo "ref rc=$?"; head -c 600 gpurun_out/${tag}_bench_ref.json; echo ;;
    launches)
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 250 --csv \
        --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_launches.log 2>&1
      echo "launches rc=$?" ;;
    full_cfg3)
      timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:count_kmers_(quad|oct)" -s 3 -c 1 \
        -f -o gpurun_out/${tag}_quad_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 3 > gpurun_out/${tag}_full_cfg3.log 2>&1
      echo "full_cfg3 rc=$?" ;;
    full_cfg2)
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:count_kmers_(quad|oct)" -s 3 -c 1 \
        -f -o gpurun_out/${tag}_quad_cfg2 python bench.py --workload cfg2 --steps 1 --warmup 3 > gpurun_out/${tag}_full_cfg2.log 2>&1
      echo "full_cfg2 rc=$?" ;;
    exp)
      # what round 1 left unverified (queue): run FIRST next round.  The final-step parts need
      # `tools/build_variant.sh finalstep -DMSBWT_FINAL_STEP` built before the snapshot is taken.
      MSBWT_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_u64_kmers.py -m gpu -q > gpurun_out/${tag}_exp_u64.log 2>&1
      echo "exp u64 rc=$?"; tail -3 gpurun_out/${tag}_exp_u64.log
      for k in 15 63 101; do
        timeout 120 python tools/pack_ab.py --k $k > gpurun_out/${tag}_exp_pack_k${k}_runtime.json 2>> gpurun_out/${tag}_exp.err
        MSBWT_PACK_FIXED_K=all timeout 120 python tools/pack_ab.py --k $k > gpurun_out/${tag}_exp_pack_k${k}_fixed.json 2>> gpurun_out/${tag}_exp.err
        cat gpurun_out/${tag}_exp_pack_k${k}_runtime.json gpurun_out/${tag}_exp_pack_k${k}_fixed.json
      done
      if [ -f build/variants/lib_finalstep.so ]; then
        MSBWT_LIBRARY_PATH=build/variants/lib_finalstep.so MSBWT_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_final_step.py -m gpu -q -x > gpurun_out/${tag}_exp_final.log 2>&1
        echo "exp final rc=$?"; tail -3 gpurun_out/${tag}_exp_final.log
        for w in cfg2 cfg3; do
          MSBWT_LIBRARY_PATH=build/variants/lib_finalstep.so timeout 300 python tools/pack_ab.py --workload $w > gpurun_out/${tag}_exp_final_${w}_off.json 2>> gpurun_out/${tag}_exp.err
          MSBWT_LIBRARY_PATH=build/variants/lib_finalstep.so MSBWT_FINAL_INDEX=1 timeout 300 python tools/pack_ab.py --workload $w > gpurun_out/${tag}_exp_final_${w}_on.json 2>> gpurun_out/${tag}_exp.err
          cat gpurun_out/${tag}_exp_final_${w}_off.json gpurun_out/${tag}_exp_final_${w}_on.json
        done
      fi ;;
  esac
done
ls -la gpurun_out | tail -20
