// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library; built only where /root/reference lies).
//
// ref_spectrum.cpp — the REFERENCE'S OWN C++ for the host half of the spectral tables, compiled from the sources where they lie:
//   Utils/Color/Spectrum.{h,cpp}        PiecewiseLinearSpectrum (.spd reader `fromFile`, `scale`, `eval`)
//   Utils/Color/SampledSpectrum.h       the resampling constructor SampledSpectrum<float>(PiecewiseLinearSpectrum) and `eval`
//   Utils/Color/SpectrumUtils.{h,cpp}   toRGB_D65 (CIE 1931 tables, D65 white)
//   Scene/SceneBuilder.cpp:911-958      genSpectralProfile -- that translation unit needs the whole engine, so the one function is cut out by
//                                       name into a temporary header (oracle/Makefile; text extraction only, no rewriting)
// with g++ directly (oracle/Makefile `_ref/libref_spectrum.so`; glm and fmt are vendored under /root/reference/external).  Two of the reference's
// headers spell fmt formatter specialisations as `struct ::fmt::formatter<...>`, which GCC rejects (MSVC / clang accept): the Makefile writes
// copies with the leading `::` removed to a temporary directory that is first on the include path and is deleted after the compile.  Hand-written here: the
// SpectralProfile struct's C layout is the C ABI's PltptSpectralProfile (BasicMaterialData.slang:43-54, checked by static_assert), the one
// constructor of Falcor::Exception the two sources call (Core/Errors.cpp:34 without the stack trace), and the extern "C" entry points.
// Used by tools/make_golden_ref.py to produce tests/golden/ref_spectral_profiles.npz, to which oracle_gen_spectral_profile and the product's
// pltpt_spd_load / pltpt_gen_spectral_profile are held (tests/test_oracle_ref_frames.py, tests/test_ingest.py).
#ifndef FMT_HEADER_ONLY
#define FMT_HEADER_ONLY 1
#endif
#include <cstring>

#include "Utils/Color/SampledSpectrum.h"
#include "Utils/Color/Spectrum.h"
#include "Utils/Color/SpectrumUtils.h"
#include "pltpt.h"

namespace Falcor {
Exception::Exception(const char* what) : mpWhat(std::make_shared<std::string>(what)) {}   // Core/Errors.cpp:34, without stack trace / error report

// Scene/Material/BasicMaterialData.slang:43-54 (a host / device shared header that needs the shader tool chain's include machinery)
struct SpectralProfile {
    static const uint32_t kMaxBins = 128;
    float3 rgb;
    float intensity;
    uint32_t bins;
    float minWavelength;
    float maxWavelength;
    float pdf[kMaxBins];
    float icdf[kMaxBins];
};
static_assert(sizeof(SpectralProfile) == sizeof(PltptSpectralProfile) && sizeof(SpectralProfile) == 1052, "the C ABI's profile record is the reference's");

#include "ref_spectral_gen.h"   // genSpectralProfile, cut from Scene/SceneBuilder.cpp into the build's temporary directory (oracle/Makefile)
}  // namespace Falcor

extern "C" {

// wavelengths / values: a piecewise-linear spectrum (what an .spd file holds); scale as addSpectralProfileForEmitterType applies it
// (SceneBuilder.cpp:985-990); computeICdf as addSpectralProfile passes it (false for the IOR tables, :972)
int ref_gen_spectral_profile(uint32_t n, const float* wavelengths, const float* values, float scale, int computeICdf, PltptSpectralProfile* out) {
    try {
        Falcor::PiecewiseLinearSpectrum spectrum(fstd::span<const float>(wavelengths, n), fstd::span<const float>(values, n));
        spectrum.scale(scale);
        const Falcor::SpectralProfile sp = Falcor::genSpectralProfile(Falcor::SampledSpectrum<float>{spectrum}, computeICdf != 0);
        std::memcpy(out, &sp, sizeof sp);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// The whole chain the scene builder runs for a table on disk (SceneBuilder.cpp:962-990): PiecewiseLinearSpectrum::fromFile (the .spd reader of
// Spectrum.cpp) -> scale -> SampledSpectrum<float>{spectrum} -> genSpectralProfile.
int ref_profile_from_spd_file(const char* path, float scale, int computeICdf, PltptSpectralProfile* out) {
    try {
        Falcor::PiecewiseLinearSpectrum spectrum = Falcor::PiecewiseLinearSpectrum::fromFile(path);
        spectrum.scale(scale);
        const Falcor::SpectralProfile sp = Falcor::genSpectralProfile(Falcor::SampledSpectrum<float>{spectrum}, computeICdf != 0);
        std::memcpy(out, &sp, sizeof sp);
        return 0;
    } catch (const std::exception&) { return -1; }
}

}  // extern "C"

#ifdef REF_SPECTRUM_TOOL
// ref_spectrum_tool <file.spd> <scale> <computeICdf>: the 1052-byte profile record to stdout.  (Python calls the .spd chain through this program: inside
// the interpreter process the reference's reader faults in libstdc++'s filesystem / stream code, which the same library called from a C++ main does not.)
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
    if (argc != 4) return 2;
    static PltptSpectralProfile p;
    if (ref_profile_from_spd_file(argv[1], float(std::atof(argv[2])), std::atoi(argv[3]), &p) != 0) return 1;
    return std::fwrite(&p, sizeof p, 1, stdout) == 1 ? 0 : 1;
}
#endif
