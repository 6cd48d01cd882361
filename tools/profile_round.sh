#!/bin/bash
# Round profile run (GPU box, one GPU): benches of every workload + reference arm, the ncu
# launch list of the default bench command, one `ncu --set full` capture per hot kernel of
# cfg2 / cfg4 / cfg5 / cfg1.  Everything lands in gpurun_out/; copy what is judged into profiles/.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r02'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for w in cfg2 cfg1 cfg3 cfg4 cfg5; do
    timeout 300 python bench.py --workload $w > $OUT/${TAG}_bench_${w}.json 2> $OUT/bench_${w}.err
done
timeout 300 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/bench_default.err
timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/bench_reference.err
# launch list of the bench command itself (times under ncu are cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-graph --no-extras \
    > $OUT/bench_under_ncu.log 2>&1
for w in cfg2 cfg4 cfg5 cfg1; do
    timeout 600 ncu --set full --clock-control none --import-source on \
        -k regex:'fold|_t_kernel|snr_moments|fbe_features|channel_mean' -s 12 -c 6 -f -o $OUT/${TAG}_ncu_${w} \
        python tools/stage_run.py $w 6 > $OUT/ncu_${w}.log 2>&1
    # gpurun brings back at most 64 MiB: keep the condensed counters, drop the reports (cfg2's stays)
    ncu -i $OUT/${TAG}_ncu_${w}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py \
        > $OUT/${TAG}_ncu_${w}_summary.csv
    python tools/ncu_roofline_rows.py $OUT/${TAG}_ncu_${w}.ncu-rep > $OUT/${TAG}_ncu_${w}_details.txt 2>/dev/null
    if [ $w != cfg2 ]; then rm -f $OUT/${TAG}_ncu_${w}.ncu-rep; fi
done
timeout 200 python tools/vs_torch_gpu.py > $OUT/${TAG}_vs_torch_gpu_cfg2.log 2>&1
ls -la $OUT | tail -30
