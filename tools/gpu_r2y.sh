#!/bin/bash
# Round 2, call Y: ncu launch list of the bench command on the final build (four frames in flight), and the refit kernels on the 10 M-triangle city.
out=gpurun_out
mkdir -p $out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/r2y_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep > $out/r2y_launches_bench.log 2>&1
python tools/launch_summary.py $out/r2y_launches.csv > $out/r2y_launch_summary.txt 2>&1; tail -14 $out/r2y_launch_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_refit" --csv --log-file $out/r2y_refit_kernels.csv python tools/overlap_refit.py refit 10000000 > $out/r2y_refit_ncu.log 2>&1
grep -v "^==" $out/r2y_refit_kernels.csv | cut -d, -f5,13- | head -30
# bench lines of the final build again with a warm-up that sizes every frame lane (call X's device-resident figures included the fourth lane's cudaMalloc)
timeout 900 python bench.py --steps 20 --warmup 3 > $out/r2y_bench_n1.json 2> $out/r2y_bench_n1.err; echo "bench exit $?"; cut -c1-200 $out/r2y_bench_n1.json
for c in cornell_closed grating layered city; do timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > $out/r2y_bench_$c.json 2> $out/r2y_bench_$c.err; echo "$c exit $?"; cut -c1-200 $out/r2y_bench_$c.json; tail -2 $out/r2y_bench_$c.err; done
