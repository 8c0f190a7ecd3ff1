#!/bin/bash
# Round 2, call A: full GPU parity suite (no -x: every measured figure is wanted), kernel variants on Cornell / city / grating,
# per-round MNEE profile, ncu captures of the traversal and manifold-walk kernels on the 10 M-triangle city.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/r2a_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --deselect tests/test_gpu_trace.py::test_closest_hit_bit_exact_10m_triangles > $out/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/r2a_pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_trace.py::test_closest_hit_bit_exact_10m_triangles -q > $out/r2a_pytest_10m.log 2>&1; echo "pytest exit $?" >> $out/r2a_pytest_10m.log
for v in base nostage dedup; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 16; done > $out/r2a_variants.log 2>&1
for v in base dedup; do for s in grating city; do PLTPT_LIB=$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4; done; done > $out/r2a_kernel_times.log 2>&1
for s in grating city; do timeout 300 python tools/mnee_profile.py $s 1920 1080 2; done > $out/r2a_mnee_profile.log 2>&1
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/r2a_$1 python tools/kernel_times.py city 1920 1080 1 > $out/r2a_ncu_$1.log 2>&1
  ncu -i /tmp/r2a_$1.ncu-rep --page raw --csv > $out/r2a_ncu_$1_raw.csv 2>>$out/r2a_ncu_$1.log
  ncu -i /tmp/r2a_$1.ncu-rep --page source --csv 2>>$out/r2a_ncu_$1.log | gzip -9 > $out/r2a_ncu_$1_source.csv.gz
}
cap city_trace "k_trace" 18 2
cap city_shadow "k_shadow_trace" 18 2
cap city_mnee "k_mnee_walk" 2 1
tail -5 $out/r2a_pytest_gpu.log; tail -3 $out/r2a_pytest_10m.log; cat $out/r2a_variants.log | cut -c1-700; cat $out/r2a_kernel_times.log | cut -c1-600
