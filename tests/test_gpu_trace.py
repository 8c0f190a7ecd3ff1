"""Ray-hit parity (north-star level 1): primitive IDs, barycentrics and t of the CUDA BVH traversal are bit-exact against
the oracle's brute-force loop over the same world-space triangles (oracle/plt_scene.h, 'Intersection contract')."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


def _scenes():
    from pltfalcor_b200 import scene
    return {"cornell": scene.cornell_box, "grating_sphere": scene.grating_sphere_scene, "layered": scene.layered_stress_scene,
            "city": lambda: scene.city_scene(40_000)}  # BASELINE config 5 generator at a size the brute-force oracle finishes in seconds (instanced spheres, all ten PLT materials)


@pytest.fixture(scope="module", params=["cornell", "grating_sphere", "layered", "city"])
def loaded(request, torch_cuda, oracle):
    from pltfalcor_b200 import binding
    buffers, cam = _scenes()[request.param]()
    p = binding.Pltpt(maxBounces=8)
    p.set_scene(buffers)
    osc = oracle.OracleScene(buffers)
    yield request.param, buffers, cam, p, osc
    p.close()


MODES = {"one_ray_per_thread": 1, "persistent_lanes": 2, "one_ray_per_thread_nearest_first": 3}  # include/pltpt.h PLTPT_TRACE_*: every traversal driver of the renderer


def _trace_gpu(torch, p, rays, mode=0):
    p.set_trace_mode(mode)
    r = torch.from_numpy(rays).cuda()
    hits = torch.zeros((len(rays), 4), dtype=torch.int32, device="cuda")
    t = torch.zeros(len(rays), dtype=torch.float32, device="cuda")
    occ = torch.zeros(len(rays), dtype=torch.int32, device="cuda")
    p.trace_closest(r, hits, t)
    p.trace_any(r, occ)
    torch.cuda.synchronize()
    return hits.cpu().numpy().view(np.uint32), t.cpu().numpy(), occ.cpu().numpy().view(np.uint32)


def test_bvh_build_stats(loaded):
    name, buffers, cam, p, osc = loaded
    s = p.stats()
    assert s["triangleCount"] == buffers.triangle_count == s["bvhLeafTriangles"]
    assert 1 <= s["bvhNodes"] <= max(1, buffers.triangle_count)
    assert s["bvhBuildMs"] > 0


@pytest.fixture(scope="module")
def brute_reference(loaded):
    """The fixed ray set of a scene and the oracle's brute-force answers (computed once, compared with both traversal drivers)."""
    name, buffers, cam, p, osc = loaded
    rng = np.random.default_rng(11)
    n_rand = 6000 if buffers.triangle_count > 1000 else 20000
    rays = np.concatenate([parity.camera_rays(cam, 96, 54), parity.random_rays(rng, buffers, n_rand), parity.random_rays(rng, buffers, 2000, tmax=0.5)])
    ref_hits, ref_t = osc.trace_closest(rays, brute=True)
    ref_occ = osc.trace_any(rays, brute=True)
    return rays, ref_hits, ref_t, ref_occ


def _assert_bit_exact(name, osc, rays, hits, t, ref_hits, ref_t):
    bad = np.nonzero((hits != ref_hits).any(axis=1) | (t.view(np.uint32) != ref_t.view(np.uint32)))[0]
    # exact ties are listed (north star): a mismatch is only tolerable where several triangles share the minimum t
    ties = [(int(i), osc.count_ties(rays[i])) for i in bad[:50]]
    assert len(bad) == 0, f"{name}: {len(bad)} of {len(rays)} rays differ; (ray, #triangles at min t) = {ties}"


@pytest.mark.parametrize("mode", sorted(MODES))
def test_closest_hit_bit_exact(loaded, brute_reference, torch_cuda, mode):
    """Both traversal drivers -- the one-ray-per-thread loop (tiny scenes, occluder chains) and the persistent-lane driver every
    rendering kernel of a non-trivial scene uses (travPersistent / processNode) -- against the brute-force loop, on every scene."""
    name, buffers, cam, p, osc = loaded
    rays, ref_hits, ref_t, ref_occ = brute_reference
    hits, t, occ = _trace_gpu(torch_cuda, p, rays, MODES[mode])
    _assert_bit_exact(f"{name}/{mode}", osc, rays, hits, t, ref_hits, ref_t)
    assert (hits[:, 0] != 0xFFFFFFFF).mean() > 0.2  # the ray set does exercise hits
    assert np.array_equal(occ, ref_occ)
    assert p.stats()["traversalStackOverflows"] == 0


def test_closest_hit_bit_exact_10m_triangles(torch_cuda, oracle):
    """BASELINE config 5 at its full size (9.99 M triangles, 1.6 M wide nodes): a fixed ray set against the oracle's brute-force loop
    over all triangles (independent of any tree), then a hundred times more rays against the oracle's own CPU BVH (a different
    builder over the same world-space triangles), through both traversal drivers."""
    from pltfalcor_b200 import binding, scene
    buffers, cam = scene.city_scene(10_000_000)
    p = binding.Pltpt(maxBounces=8)
    p.set_scene(buffers)
    osc = oracle.OracleScene(buffers)
    threads = oracle.hardware_threads()
    rng = np.random.default_rng(5)
    cam_rays = parity.camera_rays(cam, 192, 108)
    brute = np.concatenate([cam_rays[rng.choice(len(cam_rays), 1000, replace=False)], parity.random_rays(rng, buffers, 500), parity.random_rays(rng, buffers, 250, tmax=0.2)])
    ref_hits, ref_t = osc.trace_closest(brute, brute=True, threads=threads)
    ref_occ = osc.trace_any(brute, brute=True, threads=threads)
    many = np.concatenate([cam_rays, parity.random_rays(rng, buffers, 100_000), parity.random_rays(rng, buffers, 50_000, tmax=0.1)])
    tree_hits, tree_t = osc.trace_closest(many, brute=False, threads=threads)
    assert (ref_hits[:, 0] != 0xFFFFFFFF).mean() > 0.3
    for mode in sorted(MODES):
        hits, t, occ = _trace_gpu(torch_cuda, p, brute, MODES[mode])
        _assert_bit_exact(f"city10m/{mode}/brute", osc, brute, hits, t, ref_hits, ref_t)
        assert np.array_equal(occ, ref_occ)
        hits, t, occ = _trace_gpu(torch_cuda, p, many, MODES[mode])
        _assert_bit_exact(f"city10m/{mode}/tree", osc, many, hits, t, tree_hits, tree_t)
        assert np.array_equal(occ == 1, hits[:, 0] != 0xFFFFFFFF)
    s = p.stats()
    assert s["traversalStackOverflows"] == 0 and s["triangleCount"] == buffers.triangle_count
    p.close()


def test_any_hit_consistent_with_closest(loaded, torch_cuda):
    name, buffers, cam, p, osc = loaded
    rng = np.random.default_rng(12)
    rays = parity.random_rays(rng, buffers, 8192, tmax=0.7)
    hits, t, occ = _trace_gpu(torch_cuda, p, rays)
    assert np.array_equal(occ == 1, hits[:, 0] != 0xFFFFFFFF)


def test_tmin_tmax_are_exclusive(loaded, torch_cuda):
    """A hit needs tmin < t < tmax (DXR rule): re-tracing with tmax = t must miss that triangle, with tmin = t as well."""
    name, buffers, cam, p, osc = loaded
    rays = parity.camera_rays(cam, 64, 36)
    hits, t, _ = _trace_gpu(torch_cuda, p, rays)
    m = hits[:, 0] != 0xFFFFFFFF
    r2 = rays[m].copy()
    r2[:, 7] = t[m]
    h2, t2, _ = _trace_gpu(torch_cuda, p, r2)
    assert ((h2[:, 0] == 0xFFFFFFFF) | (t2 < t[m])).all()
    r3 = rays[m].copy()
    r3[:, 3] = t[m]
    h3, t3, _ = _trace_gpu(torch_cuda, p, r3)
    assert ((h3[:, 0] == 0xFFFFFFFF) | (t3 > t[m])).all()


def test_device_pointer_scene_matches_host_pointer_scene(torch_cuda, cornell):
    """The Falcor interop path hands device pointers (pointersAreDevice = 1); results must be identical."""
    from pltfalcor_b200 import binding
    buffers, cam = cornell
    rays = parity.camera_rays(cam, 64, 36)
    a = binding.Pltpt(maxBounces=4)
    a.set_scene(buffers)
    b = binding.Pltpt(maxBounces=4)
    b.set_scene(buffers, device_tensors=buffers.to_device(torch_cuda))
    ha, ta, _ = _trace_gpu(torch_cuda, a, rays)
    hb, tb, _ = _trace_gpu(torch_cuda, b, rays)
    assert np.array_equal(ha, hb) and np.array_equal(ta, tb)
    a.close(); b.close()


def test_single_triangle_and_tiny_scenes(torch_cuda, oracle):
    """Edge cases of the builder: 1, 2 and 3 triangles (no / trivial internal nodes)."""
    from pltfalcor_b200 import binding, scene
    for n_tris in (1, 2, 3):
        b = scene.SceneBuilder()
        m = b.add_diffuse((0.5, 0.5, 0.5))
        P = np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1], [0, 1, 0]], np.float32)
        idx = np.array([0, 1, 2, 0, 2, 3, 0, 1, 4][: 3 * n_tris], np.uint32)
        N = np.tile([0, 1, 0], (5, 1)).astype(np.float32)
        UV = np.zeros((5, 2), np.float32)
        b.add_mesh((P, N, UV, idx), m)
        buffers = b.finalize()
        p = binding.Pltpt()
        p.set_scene(buffers)
        osc = oracle.OracleScene(buffers)
        rng = np.random.default_rng(n_tris)
        rays = parity.random_rays(rng, buffers, 4096)
        rays[:, 0:3] += np.array([0, 0.5, 0], np.float32)
        hits, t, occ = _trace_gpu(torch_cuda, p, rays)
        ref_hits, ref_t = osc.trace_closest(rays, brute=True)
        assert np.array_equal(hits, ref_hits) and np.array_equal(t, ref_t), n_tris
        assert (hits[:, 0] != 0xFFFFFFFF).any()
        p.close()


def binding_scene(binding, buffers):
    h = binding.Pltpt(maxBounces=4)
    h.set_scene(buffers)
    return h


@pytest.mark.parametrize("mode", ["rebuild", "refit"])
def test_updated_transforms_equal_a_scene_built_with_them(torch_cuda, oracle, mode):
    """pltpt_update_transforms (animated instances; the reference rebuilds or refits its TLAS, Scene.cpp:3624-3700, by the scene's
    UpdateMode, Scene.h:169-173): after moving two meshes the BVH -- rebuilt, or refitted bottom-up with its topology kept -- answers
    exactly like the BVH of a scene created with the moved meshes, and the library-built light collection follows."""
    torch = torch_cuda
    from pltfalcor_b200 import binding, scene
    buffers, cam = scene.cornell_box()
    moved = np.array(buffers.arrays["worldMatrices"], copy=True)
    moved[6, :3, 3] += np.float32([0.25, 0.1, -0.2])      # the gold box
    moved[5, :3, 3] += np.float32([0.3, -0.4, 0.1])       # the emitter quad
    moved_it = np.stack([np.linalg.inv(m.astype(np.float64)).T for m in moved]).astype(np.float32)
    target = scene.SceneBuffers(dict(buffers.arrays, worldMatrices=moved, inverseTransposeWorldMatrices=moved_it), buffers.scalars)
    rays = np.concatenate([parity.camera_rays(cam, 128, 72), parity.random_rays(np.random.default_rng(5), target, 8000)])
    p, q = binding.Pltpt(maxBounces=4), binding.Pltpt(maxBounces=4)
    p.set_scene(buffers)
    p.build_light_collection()
    p.set_bvh_update_mode(1 if mode == "refit" else 0)
    p.update_transforms(moved, moved_it)
    assert (p.stats()["bvhRefitMs"] > 0) == (mode == "refit")
    q.set_scene(target)
    q.build_light_collection()
    hp, tp, op = _trace_gpu(torch, p, rays)
    hq, tq, oq = _trace_gpu(torch, q, rays)
    assert np.array_equal(hp, hq) and np.array_equal(tp.view(np.uint32), tq.view(np.uint32)) and np.array_equal(op, oq)
    ref_hits, ref_t = oracle.OracleScene(target).trace_closest(rays, brute=True)
    assert np.array_equal(hp, ref_hits) and np.array_equal(tp.view(np.uint32), ref_t.view(np.uint32))
    unmoved, _, _ = _trace_gpu(torch, binding_scene(binding, buffers), rays)
    assert (unmoved != hp).any()  # the move is visible to the ray set
    n = buffers.scalars["instanceCount"]
    a, b = p.get_light_collection(n), q.get_light_collection(n)
    assert a["emissiveTriangles"].tobytes() == b["emissiveTriangles"].tobytes() and a["triangleAliasTable"].tobytes() == b["triangleAliasTable"].tobytes()
    img = [torch.zeros((54, 96, 4), dtype=torch.float64, device="cuda") for _ in range(2)]
    p.render_accumulate(cam, 0, 1, 2, 96, 54, img[0]); q.render_accumulate(cam, 0, 1, 2, 96, 54, img[1])
    torch.cuda.synchronize()
    assert torch.equal(img[0], img[1])
    p.close(); q.close()


def test_refit_of_a_large_scene_keeps_hits_exact(torch_cuda, oracle):
    """Refit (PLTPT_BVH_UPDATE_REFIT) on a scene whose tree is deep: every instance of the city moved by its own offset, twice in a row
    (the second refit starts from refitted boxes); closest hits, hit distances and occlusion answers equal the brute-force oracle's on
    the moved scene bit for bit through every traversal driver, and the traversal stack never overflows."""
    torch = torch_cuda
    from pltfalcor_b200 import binding, scene
    buffers, cam = scene.city_scene(40_000)
    rng = np.random.default_rng(11)
    p = binding.Pltpt(maxBounces=4)
    p.set_scene(buffers)
    build_ms = p.stats()["bvhBuildMs"]
    p.set_bvh_update_mode(1)
    moved = np.array(buffers.arrays["worldMatrices"], copy=True)
    for step in range(2):
        moved[:, :3, 3] += rng.uniform(-0.4, 0.4, (len(moved), 3)).astype(np.float32)
        moved_it = np.stack([np.linalg.inv(m.astype(np.float64)).T for m in moved]).astype(np.float32)
        p.update_transforms(moved, moved_it)
    target = scene.SceneBuffers(dict(buffers.arrays, worldMatrices=moved, inverseTransposeWorldMatrices=moved_it), buffers.scalars)
    rays = np.concatenate([parity.camera_rays(cam, 96, 54), parity.random_rays(rng, target, 3000)])
    osc = oracle.OracleScene(target)
    ref_hits, ref_t = osc.trace_closest(rays, brute=True)
    ref_occ = osc.trace_any(rays, brute=True)
    assert (ref_hits[:, 0] != 0xFFFFFFFF).mean() > 0.2
    for mode in sorted(MODES):
        h, t, o = _trace_gpu(torch, p, rays, MODES[mode])
        _assert_bit_exact(f"city-refit/{mode}", osc, rays, h, t, ref_hits, ref_t)
        assert np.array_equal(o, ref_occ)
    s = p.stats()
    assert s["traversalStackOverflows"] == 0 and 0 < s["bvhRefitMs"]
    print(f"city 40k: build {build_ms:.2f} ms, refit {s['bvhRefitMs']:.3f} ms")
    p.close()
