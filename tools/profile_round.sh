#!/bin/bash
# Round profile run (GPU box, one GPU): benches of every workload + reference arm, the ncu
# launch list of the default bench command, one `ncu --set full` capture per hot kernel of
# cfg2 / cfg4 / cfg5 / cfg1.  Everything lands in gpurun_out/; copy what is judged into profiles/.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh v4'
TAG=${1:-v4}
OUT=gpurun_out
mkdir -p $OUT
for w in cfg2 cfg1 cfg3 cfg4 cfg5; do
    timeout 300 python bench.py --workload $w > $OUT/r01_bench_${w}_${TAG}.json 2> $OUT/bench_${w}.err
done
timeout 300 python bench.py --impl reference > $OUT/r01_bench_reference_${TAG}.json 2> $OUT/bench_reference.err
# launch list of the bench command itself (times under ncu are cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/r01_launches_cfg2_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-graph \
    > $OUT/bench_under_ncu.log 2>&1
for w in cfg2 cfg4 cfg5 cfg1; do
    timeout 600 ncu --set full --clock-control none --import-source on \
        -k regex:'fold|snr_moments|fbe_features' -s 12 -c 4 -f -o $OUT/r01_ncu_${w}_${TAG} \
        python tools/stage_run.py $w 6 > $OUT/ncu_${w}.log 2>&1
    # gpurun brings back at most 64 MiB: keep the condensed counters, drop the reports (cfg2's stays)
    ncu -i $OUT/r01_ncu_${w}_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py \
        > $OUT/r01_ncu_${w}_${TAG}_summary.csv
    if [ $w != cfg2 ]; then rm -f $OUT/r01_ncu_${w}_${TAG}.ncu-rep; fi
done
timeout 200 python tools/vs_torch_gpu.py > $OUT/r01_vs_torch_gpu_cfg2_${TAG}.log 2>&1
ls -la $OUT | tail -30
