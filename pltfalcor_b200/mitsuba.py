"""Mitsuba-style BSDF descriptions -> PLT materials of the procedural scene builder.

Follows the PLT branch of the reference's importer, Source/plugins/importers/MitsubaImporter/MitsubaImporter.cpp:286-527 (`buildBSDF`):
same plugin names, parameter names and defaults, the named-IOR table of Tables.h:13-39 (index of refraction, Abbe number), the
roughness conversion `convertGSSigma2` (:182-184) and the constructor defaults of the PLT*Material classes
(Source/Falcor/Rendering/Materials/PLT/*.cpp).  A description is a plain dict, e.g.

    {"type": "roughconductor", "material": "Au", "roughness": 0.3}
    {"type": "coated_conductor", "thickness": 2.0, "int_ior": 1.5, "bsdf": {"type": "conductor", "material": "Cu"}}
    {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": (0.5, 0.4, 0.3)}}

A colour parameter may be a texture description instead (buildTexture, :212-263): {"type": "checkerboard", "color0": .., "color1": ..}
(the importer's 512 x 512 RGBA32F pattern) or {"type": "bitmap", "texels": array (H, W, 3 or 4), "raw": bool} -- the texels themselves,
since there are no image files offline: uint8 texels are sRGB unless `raw` (Texture::createFromFile(.., loadAsSrgb = !raw)), float
texels are linear.  It goes to the slot the importer puts it in (diffuse reflectance, specular_reflectance -> BaseColor,
specular_transmittance -> Transmission, thin-dielectric birefringence -> Specular); `to_uv` is parsed and never applied there, so it is
ignored here too.  {"type": "normalmap", "bsdf": {...}, "texture": {...}} sets the nested material's normal map.  Conductor `eta` / `k`
may be curves, (wavelengths_nm, values) pairs (getSpectrum -> SceneBuilder::addSpectralProfile); RGB-valued eta / k raise.
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


def build_texture(builder, desc):
    """buildTexture (MitsubaImporter.cpp:212-263): the texture ID added to `builder`."""
    import numpy as np
    t = desc.get("type")
    if t == "checkerboard":
        c0, c1 = _rgb(desc.get("color0", 0.4)), _rgb(desc.get("color1", 0.2))
        k = 512
        half = np.arange(k) < k // 2
        pick1 = half[None, :] ^ half[:, None]                      # (x < kSize / 2) ^ (y < kSize / 2) ? color1 : color0
        px = np.ones((k, k, 4), np.float32)
        px[..., :3] = np.where(pick1[..., None], np.float32(c1), np.float32(c0))
        return builder.add_texture(px, "rgba32f")
    if t == "bitmap":
        if "texels" not in desc:
            raise UnsupportedMitsubaFeature("'bitmap': image files cannot be read here; pass the decoded image as 'texels'")
        for k in ("filter_type", "wrap_mode"):                     # ctx.unsupportedParameter: logged and ignored by the importer
            desc.get(k)
        px = np.asarray(desc["texels"])
        if px.ndim != 3 or px.shape[2] not in (3, 4):
            raise ValueError("'bitmap' texels must have shape (H, W, 3) or (H, W, 4)")
        if px.shape[2] == 3:
            one = 255 if px.dtype == np.uint8 else 1
            px = np.concatenate([px, np.full(px.shape[:2] + (1,), one, px.dtype)], axis=2)
        if px.dtype == np.uint8:
            return builder.add_texture(px, "rgba8" if desc.get("raw", False) else "rgba8_srgb")
        return builder.add_texture(px.astype(np.float32), "rgba32f")
    raise UnsupportedMitsubaFeature(f"texture type '{t}' is not supported (MitsubaImporter.cpp:258-261)")


def _rgb(v):
    return (float(v),) * 3 if isinstance(v, (int, float)) else tuple(float(x) for x in v)[:3]


def _color_or_texture(builder, props, name, default):
    """lookupTexture (MitsubaImporter.cpp:265-284): (uniform colour, texture ID or None)."""
    v = props.get(name, default)
    if isinstance(v, dict):
        return _rgb(default), build_texture(builder, v)
    if isinstance(v, str):
        raise ValueError(f"Parameter '{name}' needs to be a color or texture.")
    return _rgb(v), None


def _color(props, name, default):
    v = props.get(name, default)
    if isinstance(v, dict) or isinstance(v, str):
        raise UnsupportedMitsubaFeature(f"'{name}': a texture is not accepted for this parameter")
    if isinstance(v, (int, float)):
        return (float(v),) * 3
    return tuple(float(x) for x in v)[:3]


def _roughness(props):
    return float(props["roughness"]) if "roughness" in props else convert_gs_sigma2(float(props.get("sigma2", 100.0)))


def _conductor_profiles(builder, props, what):
    if "material" in props:
        return builder.add_ior_profiles(props["material"])          # addSpectralProfileFromMaterial
    if "eta" in props and "k" in props:
        ids = []
        for k in ("eta", "k"):
            v = props[k]
            if not (isinstance(v, (tuple, list)) and len(v) == 2 and hasattr(v[0], "__len__")):
                raise UnsupportedMitsubaFeature(f"'{what}': '{k}' must be a (wavelengths_nm, values) curve; RGB-valued IORs are not supported")
            ids.append(builder.add_profile(v[0], v[1]))              # addSpectralProfile(spectrum): computeICdf defaults to true (SceneBuilder.h:544)
        return tuple(ids)
    raise ValueError(f"'{what}' BSDF must specify either 'eta' and 'k' IOR values or 'material' name.")


def add_bsdf(builder, desc, _two_sided=False):
    """buildBSDF: returns the material id added to `builder` (scene.SceneBuilder)."""
    t = desc["type"]
    ext_ior = lookup_ior(desc, "ext_ior", "air")[0]
    if t == "diffuse":
        color, tex = _color_or_texture(builder, desc, "reflectance", 0.5)
        mid = builder.add_diffuse(color, double_sided=_two_sided)
        if tex is not None:
            builder.set_texture(mid, "baseColor", tex)
        return mid
    if t in ("dielectric", "roughdielectric"):
        ior, abbe = lookup_ior(desc, "int_ior", "bk7")
        # constructor: abbe 100 at the default IoR; the importer then sets the IoR and the table's Abbe number (setAbbeNumber: B = .52 (n-1) / abbe)
        spec, spec_tex = _color_or_texture(builder, desc, "specular_reflectance", 1.0)
        trans, trans_tex = _color_or_texture(builder, desc, "specular_transmittance", 1.0)
        mid = builder.add_dielectric(ior, _roughness(desc), ext_eta=ext_ior, dispersion_b=0.52 * (ior - 1.0) / abbe, specular=spec, transmission=trans)
        if spec_tex is not None:
            builder.set_texture(mid, "baseColor", spec_tex)
        if trans_tex is not None:
            builder.set_texture(mid, "transmission", trans_tex)
        return mid
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
        bi, bi_tex = _color_or_texture(builder, desc, "birefringence", 0.0)
        mid = builder.add_thin_dielectric(ior, float(desc.get("thickness", 0.001)), birefringence=float(desc.get("birefringence_scale", 1.0)), optic_axis_xy=(ax[0], ax[1]),
                                          birefringence_scale=float(bi[0]))
        if bi_tex is not None:
            builder.set_texture(mid, "specular", bi_tex)
        return mid
    if t in ("conductor", "roughconductor"):
        spec, spec_tex = _color_or_texture(builder, desc, "specular_reflectance", 1.0)
        mid = builder.add_conductor(_conductor_profiles(builder, desc, "conductor"), _roughness(desc), gamma=3.0, specular=spec, ext_eta=ext_ior)
        if spec_tex is not None:
            builder.set_texture(mid, "baseColor", spec_tex)
        return mid
    if t in ("coated_conductor", "coated_roughconductor"):
        nested = [v for v in desc.values() if isinstance(v, dict) and "type" in v]
        if len(nested) != 1 or nested[0]["type"] not in ("conductor", "roughconductor"):
            raise ValueError("'coatedConductor' BSDF must contain only a nested 'conductor' BSDF.")
        inner = nested[0]
        coat_ior, _ = lookup_ior(desc, "int_ior", "bk7")
        spec, spec_tex = _color_or_texture(builder, inner, "specular_reflectance", 1.0)
        mid = builder.add_coated_conductor(_conductor_profiles(builder, inner, "conductor"), _roughness(inner), tau_um=float(desc.get("thickness", 50.0)), coat_ior=coat_ior,
                                           specular=spec, ext_eta=lookup_ior(inner, "ext_ior", "air")[0])
        if spec_tex is not None:
            builder.set_texture(mid, "baseColor", spec_tex)       # setBaseColorTexture(conductor->getBaseColorTexture())
        return mid
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
        nested = [v for v in desc.values() if isinstance(v, dict) and v.get("type") not in ("bitmap", "checkerboard")]
        maps = [v for v in desc.values() if isinstance(v, dict) and v.get("type") in ("bitmap", "checkerboard")]
        if len(nested) > 1:
            raise ValueError("'normalmap' BSDF can only have one nested BSDF.")
        if len(nested) != 1 or not maps:
            raise ValueError("'normalmap' BSDF must contain a single nested BSDF and a bitmap.")
        mid = add_bsdf(builder, nested[0])
        builder.set_texture(mid, "normalMap", build_texture(builder, maps[-1]), normal_map_type=1)   # setNormalMap: three-channel map -> NormalMapType::RGB
        return mid
    # unsupported plugin: the importer logs it and substitutes a red diffuse material (MitsubaImporter.cpp:517-523)
    return builder.add_diffuse((1.0, 0.0, 0.0))
