#!/bin/bash
# Round 2, call H: replay step with approximate division / square root / rsqrt normalisation (variants), parity figures per variant,
# fresh source-level ncu capture of k_unwind_step.
out=gpurun_out
mkdir -p $out
V=pltfalcor_b200/build/variants
for v in base fm_flags fm_u fm_uc ufm_u; do timeout 300 python tools/variants.py one $V/$v/libpltpt.so 16; done > $out/r2h_variants.log 2>&1
cut -c1-700 $out/r2h_variants.log
for v in base fm_u fm_uc ufm_u; do
  PLTPT_LIB=$PWD/$V/$v/libpltpt.so PLTPT_PARITY_TAG=r2h_${v}_ timeout 900 python -m pytest tests/test_gpu_integrator.py -q -k "single_frame or option_variants" > $out/r2h_pytest_$v.log 2>&1
  tail -3 $out/r2h_pytest_$v.log
done
tag=r2h
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_unwind_step" -s 9 -c 9 -f -o /tmp/${tag}_unwind python tools/profile_frame.py 2 > $out/${tag}_unwind.log 2>&1
ncu -i /tmp/${tag}_unwind.ncu-rep --page raw --csv > $out/${tag}_unwind_raw.csv 2>>$out/${tag}_unwind.log
ncu -i /tmp/${tag}_unwind.ncu-rep --page source --print-source cuda,sass --csv 2>>$out/${tag}_unwind.log | gzip -9 > $out/${tag}_unwind_cuda.csv.gz
ls -la $out | tail -12
