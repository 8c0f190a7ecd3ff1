"""The C++ host shell: the drop-in `PLTPT` RenderPass plugin (pltfalcor_b200/host/PLTPT.cpp -> libPLTPT.so) driven by the Mogwai
stand-in (host/pltpt_host.cpp).  The CPU tests read like checks of the reference pass's host behaviour (PLTPT.cpp:43-46,102-214);
the GPU test renders through the plugin and compares with the C-ABI call made directly."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# PLTPT.h:74-108 (EmissiveLightSamplerType::Power == 2, EmissiveLightSamplerType.slangh:37-44)
REFERENCE_DEFAULTS = dict(maxBounces=32, tileSize=512, debugView=0, debugViewIntensity=1.0, sourcingAreaFromEmissiveGeometry=1.0, sourcingMaxBeamOmega=0.025, sampleGenerator=1,
                          HWSS=4, HWSSDoMIS=True, useDirectLights=True, useEnvLights=True, useEmissiveLights=True, useAnalyticLights=True, alphaMasking=True, doNEE=True,
                          NEEUsePerTileSG=False, doImportanceSampleEmitters=True, doMNEE=True, doMIS=True, emissiveSampler=2, doRussianRoulette=True, MNEEMaxOccluders=2,
                          MNEEMaxIterations=60, MNEESolverThreshold=5e-5)


@pytest.fixture(scope="module")
def host():
    from pltfalcor_b200 import build
    build.build()
    plugin, exe = build.build_host()
    return plugin, exe


def _describe(host, *props):
    plugin, exe = host
    args = [exe, "--plugin", plugin, "--describe"]
    for p in props:
        args += ["--prop", p]
    r = subprocess.run(args, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout), r.stderr


def test_plugin_exports_register_plugin(host):
    out = subprocess.run(["nm", "-D", "--defined-only", host[0]], capture_output=True, text=True).stdout
    assert " T registerPlugin" in out  # Core/Plugin.h:86-97: the one entry point a plugin library exports


def test_registers_type_pltpt_with_reference_channels(host):  # PLTPT.h:46, PLTPT.cpp:65-78
    d, _ = _describe(host)
    assert d["type"] == "PLTPT" and d["registered"] == ["PLTPT"]
    assert d["desc"] == "Physical Light Transport Path Tracer."
    ch = {c["name"]: c for c in d["channels"]}
    assert list(ch) == ["vbuffer", "viewW", "normal", "albedo", "color"]
    assert not ch["vbuffer"]["output"] and not ch["vbuffer"]["optional"]
    assert not ch["viewW"]["output"] and ch["viewW"]["optional"]
    for name in ("normal", "albedo", "color"):
        assert ch[name]["output"] and ch[name]["format"] == 1  # RGBA32Float


def test_scripting_dictionary_has_the_24_reference_keys_and_defaults(host):  # PLTPT.cpp:177-204, PLTPT.h:74-108
    d, _ = _describe(host)
    got = d["dictionary"]
    assert set(got) == set(REFERENCE_DEFAULTS) and len(got) == 24
    for k, v in REFERENCE_DEFAULTS.items():
        assert got[k] == pytest.approx(v), k


def test_dictionary_round_trip(host):  # create(dict) -> getScriptingDictionary() reproduces the values (what Mogwai's script export relies on)
    d, _ = _describe(host, "maxBounces=8", "HWSS=2", "doMNEE=false", "emissiveSampler=0", "MNEESolverThreshold=0.001", "tileSize=256", "NEEUsePerTileSG=True")
    got = d["dictionary"]
    assert (got["maxBounces"], got["HWSS"], got["doMNEE"], got["emissiveSampler"], got["tileSize"], got["NEEUsePerTileSG"]) == (8, 2, False, 0, 256, True)
    assert got["MNEESolverThreshold"] == pytest.approx(1e-3)
    assert d["logged_errors"] == 0


def test_unknown_key_is_logged_not_thrown(host):  # PLTPT.cpp:173
    d, err = _describe(host, "maxBounce=3")
    assert d["logged_errors"] == 1
    assert d["last_message"] == "Unknown field 'maxBounce' in PLTPathTracer dictionary."
    assert "Unknown field 'maxBounce'" in err
    assert d["dictionary"]["maxBounces"] == 32  # the pass is still created, with defaults


def test_scene_file_round_trip(tmp_path, cornell):
    buffers, cam = cornell
    path = tmp_path / "cornell.pltscn"
    buffers.save(str(path), cam)
    raw = path.read_bytes()
    assert raw[:8] == b"PLTSCN1\0"
    head = np.frombuffer(raw, np.uint64, 14, 8)
    assert head[0] == buffers.scalars["vertexCount"] and head[2] == buffers.scalars["instanceCount"]


def test_scene_file_carries_textures_and_samplers(tmp_path):
    """The trailing texture section of the scene file: counts, per-texture header + texels, 32 B sampler records (what pltpt_host.cpp reads)."""
    from pltfalcor_b200 import scene
    buffers, cam = scene.textured_scene()
    path = tmp_path / "textured.pltscn"
    buffers.save(str(path), cam)
    raw = path.read_bytes()
    tex, smp = buffers.scalars["textures"], buffers.scalars["samplers"]
    tail = sum(16 + 8 + buffers.arrays[f"texture{k}"].nbytes for k in range(len(tex))) + 32 * len(smp)
    counts = np.frombuffer(raw, np.uint32, 2, len(raw) - tail - 8)
    assert tuple(counts) == (len(tex), len(smp)) and len(tex) == 6 and len(smp) == 4
    hdr = np.frombuffer(raw, np.uint32, 4, len(raw) - tail)
    assert tuple(hdr[:3]) == tex[0]
    last = np.frombuffer(raw, np.float32, 4, len(raw) - 16)
    assert np.allclose(last, smp[-1][3])


@pytest.mark.gpu
def test_plugin_renders_textured_scene(host, tmp_path, torch_cuda):
    """Textures and samplers reach the plugin through the scene file: the plugin's accumulation equals the direct C-ABI call's."""
    torch = torch_cuda
    from pltfalcor_b200 import binding, scene
    buffers, cam = scene.textured_scene()
    w, h, frames = 96, 54, 2
    path = tmp_path / "textured.pltscn"
    buffers.save(str(path), cam)
    plugin, exe = host
    r = subprocess.run([exe, "--plugin", plugin, "--scene", str(path), "--width", str(w), "--height", str(h), "--frames", str(frames), "--out", str(tmp_path / "tex"),
                        "--prop", "maxBounces=6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "tex.color.f64", np.float64).reshape(h, w, 4)
    p = binding.Pltpt(device=0, maxBounces=6)
    p.set_scene(buffers)
    acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
    p.render_accumulate(cam, 0, 1, frames, w, h, acc)
    torch.cuda.synchronize()
    assert np.array_equal(got[..., :3], acc.cpu().numpy()[..., :3])
    albedo = np.fromfile(tmp_path / "tex.albedo.f32", np.float32).reshape(h, w, 4)
    assert albedo[..., :3].std() > 0.05   # the checker texture shows in the guide
    p.close()


@pytest.mark.gpu
def test_render_through_plugin_matches_direct_c_abi(host, tmp_path, cornell, torch_cuda):
    """execute() x N through registerPlugin/createClass/setScene + host-side Halton jitter and double accumulation
    == pltpt_render_accumulate called directly (same frames, same jitter): bit-identical sums."""
    torch = torch_cuda
    from pltfalcor_b200 import binding
    buffers, cam = cornell
    w, h, frames = 128, 72, 4
    path = tmp_path / "cornell.pltscn"
    buffers.save(str(path), cam)
    plugin, exe = host
    r = subprocess.run([exe, "--plugin", plugin, "--scene", str(path), "--width", str(w), "--height", str(h), "--frames", str(frames), "--out", str(tmp_path / "img"),
                        "--prop", "maxBounces=8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["frames"] == frames and info["cleared_without_scene"] and info["logged_errors"] == 0
    got = np.fromfile(tmp_path / "img.color.f64", np.float64).reshape(h, w, 4)
    p = binding.Pltpt(device=0, maxBounces=8)
    p.set_scene(buffers)
    acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
    p.render_accumulate(cam, 0, 1, frames, w, h, acc)
    torch.cuda.synchronize()
    ref = acc.cpu().numpy()
    assert np.array_equal(got[..., :3], ref[..., :3])
    assert np.all(got[..., 3] == frames) and float(got[..., :3].mean()) > 0
    albedo = np.fromfile(tmp_path / "img.albedo.f32", np.float32).reshape(h, w, 4)
    assert albedo[..., 3].max() == pytest.approx(1.0) and np.isfinite(albedo).all()
    p.close()


@pytest.mark.gpu
def test_plugin_renders_env_lit_scene_and_debug_view(host, tmp_path, torch_cuda):
    """The scene file's environment-map section reaches the plugin (Scene::useEnvLight), `useEnvLights` and `debugView` are honoured."""
    torch = torch_cuda
    from pltfalcor_b200 import binding, scene
    buffers, cam = scene.env_scene()
    w, h, frames = 96, 54, 2
    path = tmp_path / "env.pltscn"
    buffers.save(str(path), cam)
    plugin, exe = host

    def run(tag, *props):
        args = [exe, "--plugin", plugin, "--scene", str(path), "--width", str(w), "--height", str(h), "--frames", str(frames), "--out", str(tmp_path / tag), "--prop", "maxBounces=6"]
        for pr in props:
            args += ["--prop", pr]
        r = subprocess.run(args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        return np.fromfile(tmp_path / (tag + ".color.f64"), np.float64).reshape(h, w, 4)

    lit, dark, normals = run("lit"), run("dark", "useEnvLights=false"), run("nrm", "debugView=5")
    for opts, got in ((dict(), lit), (dict(useEnvLights=False), dark), (dict(debugView=5), normals)):
        p = binding.Pltpt(device=0, maxBounces=6, **opts)
        p.set_scene(buffers)
        acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
        p.render_accumulate(cam, 0, 1, frames, w, h, acc)
        torch.cuda.synchronize()
        assert np.allclose(got[..., :3], acc.cpu().numpy()[..., :3], rtol=1e-12, atol=0), opts
        p.close()
    assert (lit[:4, :, :3] > 0).all() and (dark[:4, :, :3] == 0).all()  # the top rows look at the sky


@pytest.mark.gpu
def test_plugin_follows_moved_instances(host, tmp_path, cornell, torch_cuda):
    """RenderPass::onSceneUpdates(SceneGraphChanged) -> pltpt_update_transforms: after the animation keyframe the plugin renders what
    a pass created on the moved scene renders (same frame indices), bit for bit."""
    torch = torch_cuda
    from pltfalcor_b200 import binding, scene
    buffers, cam = cornell
    moved = np.array(buffers.arrays["worldMatrices"], copy=True)
    moved[7, :3, 3] += np.float32([-0.4, 0.2, 0.3])  # the glass box
    moved_it = np.stack([np.linalg.inv(m.astype(np.float64)).T for m in moved]).astype(np.float32)
    target = scene.SceneBuffers(dict(buffers.arrays, worldMatrices=moved, inverseTransposeWorldMatrices=moved_it), buffers.scalars)
    a, b = tmp_path / "a.pltscn", tmp_path / "b.pltscn"
    buffers.save(str(a), cam); target.save(str(b), cam)
    w, h, frames, after = 96, 54, 5, 2
    plugin, exe = host
    r = subprocess.run([exe, "--plugin", plugin, "--scene", str(a), "--moved-scene", str(b), "--move-after", str(after), "--width", str(w), "--height", str(h),
                        "--frames", str(frames), "--out", str(tmp_path / "anim"), "--prop", "maxBounces=6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "anim.color.f64", np.float64).reshape(h, w, 4)
    p = binding.Pltpt(device=0, maxBounces=6)
    p.set_scene(target)
    acc = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda")
    p.render_accumulate(cam, after, 1, frames - after, w, h, acc)
    torch.cuda.synchronize()
    assert np.allclose(got[..., :3], acc.cpu().numpy()[..., :3], rtol=1e-12, atol=0)
    still = binding.Pltpt(device=0, maxBounces=6)
    still.set_scene(buffers)
    acc2 = torch.zeros_like(acc)
    still.render_accumulate(cam, after, 1, frames - after, w, h, acc2)
    torch.cuda.synchronize()
    assert not np.allclose(got[..., :3], acc2.cpu().numpy()[..., :3], rtol=1e-6, atol=0)  # the move is visible
    p.close(); still.close()


@pytest.mark.gpu
def test_plugin_reports_library_errors_as_exceptions(host, tmp_path, cornell, torch_cuda):
    """A failing C-ABI call surfaces as Falcor::RuntimeError (exit code 1 of the harness), e.g. the LightBVH sampler the pass does
    not offer: 'Unknown emissive light sampler type' (PLTPT.cpp:409)."""
    buffers, cam = cornell
    path = tmp_path / "cornell.pltscn"
    buffers.save(str(path), cam)
    plugin, exe = host
    r = subprocess.run([exe, "--plugin", plugin, "--scene", str(path), "--width", "32", "--height", "32", "--frames", "1", "--out", str(tmp_path / "x"),
                        "--prop", "emissiveSampler=1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Unknown emissive light sampler type" in r.stderr
