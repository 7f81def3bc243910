#!/bin/bash
# One gpurun call = several measurements, each under its own timeout, everything logged under gpurun_out/.
#   tools/gpu_session.sh <tag> <step> [<step> ...]     steps: pytest kb:<which...> exp bench bench_ref launches
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
for step in "$@"; do
  case "$step" in
    pytest)      timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_$TAG.log ;;
    pytestall)   timeout 1200 python -m pytest tests -m gpu -q -s > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_$TAG.log ;;
    pytest:*)    timeout 900 python -m pytest tests -m gpu -x -q -s -k "${step#pytest:}" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_$TAG.log ;;
    kb:*)        for w in $(echo "${step#kb:}" | tr ',' ' '); do timeout 600 python tools/kernel_bench.py $w >> $OUT/kb_$TAG.jsonl 2>> $OUT/kb_$TAG.err; done; cat $OUT/kb_$TAG.jsonl ;;
    exp)         timeout 600 python tools/exp_two_microbatches.py > $OUT/exp2mb_$TAG.jsonl 2> $OUT/exp2mb_$TAG.err; cat $OUT/exp2mb_$TAG.jsonl; tail -3 $OUT/exp2mb_$TAG.err ;;
    exp8)        CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 600 python tools/exp_two_microbatches.py baseline stag:full:12:32 > $OUT/exp2mb_${TAG}_conn8.jsonl 2> $OUT/exp2mb_${TAG}_conn8.err; cat $OUT/exp2mb_${TAG}_conn8.jsonl; tail -3 $OUT/exp2mb_${TAG}_conn8.err ;;
    exp:*)       timeout 600 python tools/exp_two_microbatches.py $(echo "${step#exp:}" | tr ',' ' ') > $OUT/exp2mb_$TAG.jsonl 2> $OUT/exp2mb_$TAG.err; cat $OUT/exp2mb_$TAG.jsonl; tail -3 $OUT/exp2mb_$TAG.err ;;
    bench)       timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err ;;
    bench_ref)   timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err; cat $OUT/bench_${TAG}_reference.json ;;
    launches)    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1; echo "launches rc=$?" ;;
    lat)         timeout 300 python tools/time_mpe.py > $OUT/lat_$TAG.jsonl 2> $OUT/lat_$TAG.err; cat $OUT/lat_$TAG.jsonl; tail -3 $OUT/lat_$TAG.err ;;
    phases)      timeout 300 python tools/step_phases.py > $OUT/phases_$TAG.json 2> $OUT/phases_$TAG.err; cat $OUT/phases_$TAG.json; tail -3 $OUT/phases_$TAG.err ;;
    ncu_lat)     timeout 600 ncu --set full --clock-control none --import-source on -k regex:lat_ -c 6 -o $OUT/lat_$TAG -f python tools/time_mpe.py 1500,1200,900,600 > $OUT/ncu_lat_$TAG.log 2>&1; echo "ncu_lat rc=$?" ;;
    ncu_fbank)   timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank -c 2 -o $OUT/fbank_$TAG -f python tools/kernel_bench.py fbank > $OUT/ncu_fbank_$TAG.log 2>&1; echo "ncu_fbank rc=$?" ;;
    timeline)    timeout 300 python tools/timeline.py > $OUT/timeline_$TAG.csv 2> $OUT/timeline_$TAG.err; wc -l $OUT/timeline_$TAG.csv; tail -3 $OUT/timeline_$TAG.err ;;
    c3)          timeout 600 python tools/bench_c3.py > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; timeout 600 python tools/bench_c3.py --criterion smbr >> $OUT/bench_c3_$TAG.json 2>> $OUT/bench_c3_$TAG.err; cat $OUT/bench_c3_$TAG.json; tail -3 $OUT/bench_c3_$TAG.err ;;
    emul)        timeout 200 python bench.py --emulate-shard 0/8 --steps 3 --no-cpu-baseline --watchdog 150 > $OUT/bench_${TAG}_shard0of8.json 2> $OUT/bench_${TAG}_shard0of8.err; cut -c1-300 $OUT/bench_${TAG}_shard0of8.json; tail -4 $OUT/bench_${TAG}_shard0of8.err ;;
    ncu_den)     timeout 900 ncu --set full --clock-control none --import-source on -k regex:den_ -s 5 -c 5 -o $OUT/den_$TAG -f python tools/kernel_bench.py den1 > $OUT/ncu_den_$TAG.log 2>&1; echo "ncu_den rc=$?" ;;
    py:*)        timeout 900 python ${step#py:} > $OUT/py_$TAG.log 2>&1; echo "py rc=$?"; tail -30 $OUT/py_$TAG.log ;;
    *)           echo "unknown step $step" ;;
  esac
done
