#!/bin/bash
# Full ncu capture of the BSDF sweep kernel (BASELINE configs[0]): conductor and dielectric launches.
out=gpurun_out
mkdir -p $out
cat > /tmp/sweep_once.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
from pltfalcor_b200 import binding, scene
sb, ids = scene.sweep_scene()
q = binding.Pltpt(); q.set_scene(sb)
print(bench.bsdf_sweep(torch, q, ids, 6543.7))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bsdf_eval_mueller" -s 2 -c 1 -f -o /tmp/sweep python /tmp/sweep_once.py > $out/r1_sweep_ncu.log 2>&1
ncu -i /tmp/sweep.ncu-rep --page raw --csv > $out/r1_sweep_raw.csv 2>>$out/r1_sweep_ncu.log
ncu -i /tmp/sweep.ncu-rep --page source --csv 2>>$out/r1_sweep_ncu.log | gzip -9 > $out/r1_sweep_source.csv.gz
tail -2 $out/r1_sweep_ncu.log
