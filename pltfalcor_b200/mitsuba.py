"""Mitsuba-style BSDF descriptions -> PLT materials of the procedural scene builder.

Follows the PLT branch of the reference's importer, Source/plugins/importers/MitsubaImporter/MitsubaImporter.cpp:286-527 (`buildBSDF`):
same plugin names, parameter names and defaults, the named-IOR table of Tables.h:13-39 (index of refraction, Abbe number), the
roughness conversion `convertGSSigma2` (:182-184) and the constructor defaults of the PLT*Material classes
(Source/Falcor/Rendering/Materials/PLT/*.cpp).  A description is a plain dict, e.g.

    {"type": "roughconductor", "material": "Au", "roughness": 0.3}
    {"type": "coated_conductor", "thickness": 2.0, "int_ior": 1.5, "bsdf": {"type": "conductor", "material": "Cu"}}
    {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": (0.5, 0.4, 0.3)}}

Textures, `normalmap` and spectra given as curves are not supported (the CUDA path takes uniform parameters only); they raise.
Harness code: the product is the C ABI, which takes the packed material blobs this module ends up producing through scene.SceneBuilder."""
import math

INF = float("inf")
# Tables.h:13-39: (IOR, Abbe number); Abbe number of infinity = no dispersion
IOR_TABLE = {
    "vacuum": (1.0, INF), "helium": (1.000036, INF), "hydrogen": (1.000132, INF), "air": (1.000277, INF), "carbon dioxide": (1.00045, INF),
    "water": (1.3330, 55.0), "acetone": (1.36, 54.0), "ethanol": (1.361, 59.0), "carbon tetrachloride": (1.461, 49.0), "glycerol": (1.4729, INF),
    "benzene": (1.501, 30.3), "silicone oil": (1.52045, INF), "bromine": (1.661, 31.7),
    "water ice": (1.31, 64.0), "fused quartz": (1.458, 67.8), "pyrex": (1.470, 60.0), "acrylic glass": (1.49, INF), "polypropylene": (1.49, INF),
    "bk7": (1.5046, 64.0), "sodium chloride": (1.544, 42.9), "amber": (1.55, INF), "pet": (1.5750, INF), "diamond": (2.419, 55.3),
}
GRATING_TYPES = {"sinusoidal": 0, "rectangular": 1, "linear": 2, "radial_sinusoidal": 0 | 0x10, "radial_rectangular": 1 | 0x10}  # DiffractionGrating.slangh:33-40


class UnsupportedMitsubaFeature(ValueError):
    pass


def lookup_ior(props, name, default):
    """MitsubaImporter.cpp:170-180 + Tables.h:41-54: a float is taken as is (no dispersion); a name goes through the table; an unknown
    name falls back to air."""
    v = props.get(name, default)
    if isinstance(v, (int, float)):
        return float(v), INF
    return IOR_TABLE.get(str(v).lower(), IOR_TABLE["air"])


def convert_gs_sigma2(sigma2):
    """MitsubaImporter.cpp:182-184."""
    return min(max(2.0 * (1.0 - max(0.0, sigma2 / 1000.0) ** (1.0 / 32.0)), 0.0), 1.0)


def _color(props, name, default):
    v = props.get(name, default)
    if isinstance(v, dict) or isinstance(v, str):
        raise UnsupportedMitsubaFeature(f"'{name}': textures are not supported by the CUDA path (uniform parameters only)")
    if isinstance(v, (int, float)):
        return (float(v),) * 3
    return tuple(float(x) for x in v)[:3]


def _roughness(props):
    return float(props["roughness"]) if "roughness" in props else convert_gs_sigma2(float(props.get("sigma2", 100.0)))


def _conductor_profiles(builder, props, what):
    if "material" in props:
        return builder.add_ior_profiles(props["material"])          # addSpectralProfileFromMaterial
    if "eta" in props and "k" in props:
        raise UnsupportedMitsubaFeature(f"'{what}': eta / k given as spectra are not supported; name a tabulated 'material'")
    raise ValueError(f"'{what}' BSDF must specify either 'eta' and 'k' IOR values or 'material' name.")


def add_bsdf(builder, desc, _two_sided=False):
    """buildBSDF: returns the material id added to `builder` (scene.SceneBuilder)."""
    t = desc["type"]
    ext_ior = lookup_ior(desc, "ext_ior", "air")[0]
    if t == "diffuse":
        return builder.add_diffuse(_color(desc, "reflectance", 0.5), double_sided=_two_sided)
    if t in ("dielectric", "roughdielectric"):
        ior, abbe = lookup_ior(desc, "int_ior", "bk7")
        # constructor: abbe 100 at the default IoR; the importer then sets the IoR and the table's Abbe number (setAbbeNumber: B = .52 (n-1) / abbe)
        return builder.add_dielectric(ior, _roughness(desc), ext_eta=ext_ior, dispersion_b=0.52 * (ior - 1.0) / abbe, specular=_color(desc, "specular_reflectance", 1.0),
                                      transmission=_color(desc, "specular_transmittance", 1.0))
    if t == "thindielectric":
        for k in ("specular_reflectance", "specular_transmittance"):
            if k in desc:
                raise UnsupportedMitsubaFeature(f"'{k}' is not supported by the thin dielectric (MitsubaImporter.cpp:347-348)")
        ior, _ = lookup_ior(desc, "int_ior", "bk7")
        ax = [float(x) for x in desc.get("optic_axis", (0.0, 0.0, 1.0))]
        n = math.sqrt(sum(x * x for x in ax))
        ax = [x / n for x in ax]
        if ax[2] < 0:
            ax = [-x for x in ax]                                    # setBirefringenceOpticAxis
        # sic (MitsubaImporter.cpp:355-360): the material's birefringence slot (baseColor.x) receives the `birefringence_scale` property
        # and its scale slot (specular.x) the `birefringence` property
        return builder.add_thin_dielectric(ior, float(desc.get("thickness", 0.001)), birefringence=float(desc.get("birefringence_scale", 1.0)), optic_axis_xy=(ax[0], ax[1]),
                                           birefringence_scale=float(_color(desc, "birefringence", 0.0)[0]))
    if t in ("conductor", "roughconductor"):
        return builder.add_conductor(_conductor_profiles(builder, desc, "conductor"), _roughness(desc), gamma=3.0, specular=_color(desc, "specular_reflectance", 1.0), ext_eta=ext_ior)
    if t in ("coated_conductor", "coated_roughconductor"):
        nested = [v for v in desc.values() if isinstance(v, dict) and "type" in v]
        if len(nested) != 1 or nested[0]["type"] not in ("conductor", "roughconductor"):
            raise ValueError("'coatedConductor' BSDF must contain only a nested 'conductor' BSDF.")
        inner = nested[0]
        coat_ior, _ = lookup_ior(desc, "int_ior", "bk7")
        return builder.add_coated_conductor(_conductor_profiles(builder, inner, "conductor"), _roughness(inner), tau_um=float(desc.get("thickness", 50.0)), coat_ior=coat_ior,
                                            specular=_color(inner, "specular_reflectance", 1.0), ext_eta=lookup_ior(inner, "ext_ior", "air")[0])
    if t == "grating":
        gname = desc.get("type_name", desc.get("grating_type", "sinusoidal"))
        if gname not in GRATING_TYPES:
            raise ValueError("Unsupported 'type' name.")
        return builder.add_grated_conductor(_conductor_profiles(builder, desc, "grating"), (float(desc.get("pitch", 5.0)), 0.0), q=1.0, roughness=0.0,
                                            direction=float(desc.get("grating_direction", 0.0)) * math.pi / 180.0, q_scale=float(desc.get("q", 0.5)),
                                            power_mult=float(desc.get("amplify", 1.0)), gtype=GRATING_TYPES[gname], lobes=int(desc.get("lobes", 3)), ext_eta=ext_ior)
    if t == "twosided":
        nested = [v for v in desc.values() if isinstance(v, dict) and "type" in v]
        if len(nested) != 1:
            raise ValueError("'twosided' BSDF must contain a single nested BSDF.")
        mid = add_bsdf(builder, nested[0])
        builder.materials[mid]["header"][0] |= 1 << 28                # setDoubleSided(true)
        return mid
    if t == "normalmap":
        raise UnsupportedMitsubaFeature("'normalmap' needs a texture; the CUDA path takes uniform parameters only")
    # unsupported plugin: the importer logs it and substitutes a red diffuse material (MitsubaImporter.cpp:517-523)
    return builder.add_diffuse((1.0, 0.0, 0.0))
