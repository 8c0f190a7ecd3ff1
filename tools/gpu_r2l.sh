#!/bin/bash
# Round 2, call L: kernels instantiated without the textured path for untextured scenes, tangent decode skipped on meshes without tangents;
# texture tests + integrator parity + per-kernel times on Cornell / grating / city / layered.
out=gpurun_out
mkdir -p $out
timeout 300 python tools/variants.py one pltfalcor_b200/libpltpt.so 16 > $out/r2l_variants.log 2>&1; cut -c1-700 $out/r2l_variants.log
timeout 900 python -m pytest tests/test_gpu_textures.py tests/test_gpu_integrator.py tests/test_gpu_trace.py -q -x --deselect tests/test_gpu_trace.py::test_closest_hit_bit_exact_10m_triangles > $out/r2l_pytest.log 2>&1; echo "pytest exit $?" >> $out/r2l_pytest.log; tail -8 $out/r2l_pytest.log
for s in grating layered city; do timeout 300 python tools/kernel_times.py $s 1920 1080 4; done > $out/r2l_kernel_times.log 2>&1; cut -c1-600 $out/r2l_kernel_times.log
