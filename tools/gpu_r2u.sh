#!/bin/bash
# Round 2, call U: fp16 -> fp32 by the hardware conversion; Cornell per-kernel times, sample-stage / integrator / texture / BSDF parity.
out=gpurun_out
mkdir -p $out
timeout 300 python tools/variants.py one pltfalcor_b200/libpltpt.so 16 > $out/r2u_variants.log 2>&1; cut -c1-700 $out/r2u_variants.log
timeout 1500 python -m pytest tests/test_gpu_integrator.py tests/test_gpu_textures.py tests/test_gpu_bsdf.py tests/test_ingest.py -q > $out/r2u_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2u_pytest.log; tail -6 $out/r2u_pytest.log
cp $out/parity_metrics.json $out/r2u_parity_metrics.json
