#!/bin/bash
# Full ncu captures of the top kernels of one 1080p Cornell-box frame (second frame; the first is warm-up).
# Reports stay in /tmp on the GPU box; CSV pages (and reports small enough) are copied to gpurun_out/.
# usage: tools/ncu_capture.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/${tag}_$1 python tools/profile_frame.py 2 > $out/${tag}_$1.log 2>&1
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv > $out/${tag}_$1_raw.csv 2>>$out/${tag}_$1.log
  ncu -i /tmp/${tag}_$1.ncu-rep --page source --csv 2>>$out/${tag}_$1.log | gzip -9 > $out/${tag}_$1_source.csv.gz
  ncu -i /tmp/${tag}_$1.ncu-rep --page source --print-source cuda,sass --csv 2>>$out/${tag}_$1.log | gzip -9 > $out/${tag}_$1_cuda.csv.gz
}
cap unwind "k_unwind_step" 9 9   # all nine replay steps of the second frame: per-launch averages comparable with bench.py's live timers
cap sample "k_shade|k_trace" 18 4
cap connect "k_connect_a|k_shadow_trace|k_connect_b" 27 6   # three kernels per bounce (a, shadow, b_slow): bounces 0 and 1 of the second frame
# gpurun only merges gpurun_out/ back when it is below 64 MiB in total: the .ncu-rep files stay on the box
du -sh $out; ls -la $out /tmp/*.ncu-rep
