// ORACLE — TEST INFRASTRUCTURE ONLY (see plt_math.h header).
// Colour maps of the PLTPT debug views (reference: ccMagma / ccInferno, 256-entry tables with linear interpolation,
// Source/RenderPasses/PLTPT/ColourCoding.slang:275-540, 542-807).  Restated as piecewise polynomials fitted to those tables by
// tools/make_colormap_fit.py (four degree-5 segments per channel; the fit error is printed below and checked against
// tests/golden/colormaps.json).  A debug overlay is for the eye; it is not part of the radiometric path.
// ccMagma: four degree-5 least-squares segments of the 256-entry table, max abs error 0.00522
// ccInferno: four degree-5 least-squares segments of the 256-entry table, max abs error 0.00835
#pragma once
#include "plt_math.h"

namespace orc {
static const float kMagmaFit[4][3][6] = {{{2.2381228e-03f, 2.5308679e-03f, 8.8498560e-01f, -1.5483508e+00f, 1.6596858e+00f, -6.8589240e-01f}, {1.8331593e-03f, -3.1460768e-03f, 7.7789785e-01f, -1.2332009e+00f, 2.5476926e-01f, 2.7635312e-01f}, {8.6443273e-03f, 5.0550900e-01f, -3.2727577e-01f, 1.5635110e+00f, -1.8350667e+00f, 5.6721593e-01f}},
 {{3.1667111e-01f, 3.9967019e-01f, -5.1595437e-02f, 1.0025436e-01f, -5.7399439e-02f, 8.8014214e-03f}, {7.1523151e-02f, 1.3888676e-01f, 1.2934028e-01f, -3.1442361e-01f, 2.6546082e-01f, -7.5906015e-02f}, {4.8546961e-01f, 1.3025467e-01f, -2.8385012e-01f, 3.5034100e-01f, -2.9899712e-01f, 9.2130909e-02f}},
 {{7.1645920e-01f, 4.0965803e-01f, -1.9838030e-02f, 3.5942110e-02f, -3.7691684e-01f, 2.2175325e-01f}, {2.1458490e-01f, 1.5532146e-01f, -5.9076385e-02f, 4.2516500e-01f, -1.4437000e-01f, -5.6625566e-02f}, {4.7579476e-01f, -1.4653719e-01f, 9.1668551e-02f, -8.1922285e-01f, 1.2974737e+00f, -5.1690015e-01f}},
 {{9.8670590e-01f, 6.1922047e-02f, -1.1865591e-01f, 1.0290488e-01f, -8.1433424e-02f, 3.5443290e-02f}, {5.3579133e-01f, 4.6858379e-01f, 5.1195364e-02f, -2.1823834e-01f, 2.8588493e-01f, -1.2631353e-01f}, {3.8250142e-01f, 1.7661507e-01f, 4.1630379e-01f, -5.1914803e-01f, 4.9503995e-01f, -1.9601760e-01f}}};
static const float kInfernoFit[4][3][6] = {{{1.4787414e-03f, 3.6751310e-02f, 5.4161886e-01f, -1.1409110e-01f, -2.2761360e-01f, 1.0268070e-01f}, {4.1063421e-03f, -9.8045243e-02f, 1.6312149e+00f, -4.2751363e+00f, 4.1486880e+00f, -1.3465005e+00f}, {9.7320047e-03f, 4.9678165e-01f, -1.9210330e-01f, 1.9033041e+00f, -3.2395578e+00f, 1.4526437e+00f}},
 {{3.4150525e-01f, 4.0144167e-01f, -2.0836375e-02f, 5.2216894e-02f, -4.5371508e-02f, 6.7278249e-03f}, {6.2270997e-02f, 1.4863648e-01f, 1.6030348e-02f, -9.9373704e-02f, 1.2174894e-01f, -3.3424794e-02f}, {4.2946393e-01f, 5.1887038e-02f, -1.9573989e-01f, 1.1610300e-01f, -1.1672732e-01f, 4.5288348e-02f}},
 {{7.3569958e-01f, 3.6764049e-01f, -6.2249737e-02f, -9.3448598e-02f, 3.3710621e-02f, -2.9206021e-03f}, {2.1589449e-01f, 2.0405012e-01f, 1.1168130e-01f, 1.1148865e-01f, -1.1858878e-01f, 3.3410856e-02f}, {3.2992707e-01f, -2.2107073e-01f, -2.1314454e-01f, 3.9623957e-01f, -4.4381064e-01f, 1.8591876e-01f}},
 {{9.7918506e-01f, 5.9351113e-02f, -4.2589003e-02f, -1.7763508e-01f, -2.3464711e-01f, 4.1358195e-01f}, {5.5765378e-01f, 4.6408360e-01f, 2.0993310e-03f, 1.0345698e-01f, -4.2376193e-02f, -8.4870893e-02f}, {4.0354673e-02f, -5.0338033e-01f, 3.8974506e+00f, -7.7197422e+00f, 8.1019429e+00f, -3.1619149e+00f}}};
inline float3 colourMapEval(const float (*fit)[3][6], float val) {
    // the table rule on a non-finite argument: lerp(data[l], data[h], frac(val)) with frac(NaN) = frac(+-inf) = NaN is NaN, which writeDebugOutput then
    // rejects (pltpt_debug.slangh:21) -- found by rendering the views with the reference's own text (tests/golden/ref_debug_views.npz)
    if (!std::isfinite(val)) return float3(std::numeric_limits<float>::quiet_NaN());
    val = hlsl_min(hlsl_max(val, 0.f), 255.f / 256.f + 0.99f / 256.f);
    const int s = hlsl_min(int(val * 4.f), 3);
    const float u = val * 4.f - float(s);
    float c[3];
    for (int k = 0; k < 3; ++k) {
        const float* a = fit[s][k];
        c[k] = a[0] + u * (a[1] + u * (a[2] + u * (a[3] + u * (a[4] + u * a[5]))));
    }
    return float3(c[0], c[1], c[2]);
}
inline float3 ccMagma(float val) { return colourMapEval(kMagmaFit, val); }
inline float3 ccInferno(float val) { return colourMapEval(kInfernoFit, val); }
enum : uint32_t { DV_None = 0, DV_NaNs, DV_Inf, DV_PathLength, DV_HwssComps, DV_Normals, DV_Albedo, DV_Roughness, DV_UVs, DV_CoherenceArea, DV_CoherenceAnisotropy,
                  DV_Polarization, DV_MneeIterations, DV_MneeBounces, DV_Count };
}  // namespace orc
