#!/bin/bash
# Round 2, call F: unrolled wavelength loops (variant) on Cornell / sweep / grating / city / layered.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
for v in base lam; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 16; done > $out/r2f_variants.log 2>&1
cut -c1-700 $out/r2f_variants.log
for v in base lam; do for s in city grating layered; do PLTPT_LIB=$V/$v/libpltpt.so timeout 300 python tools/kernel_times.py $s 1920 1080 4 | sed "s|^|$v |"; done; done > $out/r2f_kernel_times.log 2>&1
cut -c1-500 $out/r2f_kernel_times.log
