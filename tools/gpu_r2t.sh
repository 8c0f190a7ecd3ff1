#!/bin/bash
# Round 2, call T: ncu evidence for the final build -- launch list of the bench command, full captures of the stage kernels on the Cornell
# box (traffic / issue / lane utilisation per kernel), and of the traversal and manifold-walk kernels on the 10 M-triangle city.
out=gpurun_out
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep > $out/r2t_launches_bench.log 2>&1
bash tools/ncu_capture.sh r2t
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/r2t_$1 python tools/kernel_times.py city 1920 1080 1 > $out/r2t_ncu_$1.log 2>&1
  ncu -i /tmp/r2t_$1.ncu-rep --page raw --csv > $out/r2t_ncu_$1_raw.csv 2>>$out/r2t_ncu_$1.log
  ncu -i /tmp/r2t_$1.ncu-rep --page source --csv 2>>$out/r2t_ncu_$1.log | gzip -9 > $out/r2t_ncu_$1_source.csv.gz
}
cap city_trace "k_trace" 18 2
cap city_mnee "k_mnee_walk" 2 1
du -sh $out; ls $out | grep r2t | head -40
