"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical sample sets.

Tolerances (BASELINE.json north_star): images within a per-channel mean absolute error of 1e-3;
geodesic endpoints within 1e-4 relative error.  Everything discrete (faces, object refs, u8
pixels up to the documented libm ulp) is compared exactly.
"""
import numpy as np
import pytest

import oracle_ffi as O
from common import (LENS_SCENE, LENS_VOLUME, SCENES, engine_render, load_pair, mae_per_channel, oracle_render)

pytestmark = pytest.mark.gpu

IMAGE_MAE = 1e-3        # north_star: per-channel MAE on identical sample sets
ENDPOINT_REL = 1e-4     # north_star: geodesic endpoints, relative


def _res(name):
    return (96, 96) if "cornell" in name else (128, 72)


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("subsample", [0, 2])
def test_camera_rays(oracle, name, subsample):
    import bendy_tracer_b200 as bt
    w, h = _res(name)
    osc, esc, cam = load_pair(name, w, h)
    rng = np.random.default_rng(1)
    n = 4096
    xs, ys = rng.integers(0, w, n), rng.integers(0, h, n)
    pi = rng.integers(0, 64, n)
    ref = osc.camera_rays(cam, O.make_config(samples=16, subsample=subsample), w, h, xs, ys, pi, seed=7, sample_base=3)
    tr = bt.Tracer(bt.Config(), seed=7)
    got = tr.camera_rays(esc, cam, bt.RenderConfig.with_samples_subsample(16, bt.Subsample(subsample)), w, h, xs, ys, pi,
                         sample_base=3)
    assert np.abs(got - ref).max() <= 2e-6   # sincosf (CUDA) vs sinf/cosf (glibc): <= 2 ulp on unit vectors
    assert np.abs(got[:, :3] - ref[:, :3]).max() <= 1e-6     # origins: camera translation (+ DoF offset)


@pytest.mark.parametrize("name", SCENES)
def test_trace_segments_flat(oracle, name):
    """ChunkState::try_hit on the camera rays of every shipped scene"""
    import bendy_tracer_b200 as bt
    w, h = _res(name)
    osc, esc, cam = load_pair(name, w, h)
    ys, xs = np.mgrid[0:h, 0:w]
    xs, ys = xs.ravel(), ys.ravel()
    pi = np.zeros(len(xs), np.uint64)
    cfg = O.make_config(samples=1)
    rays = osc.camera_rays(cam, cfg, w, h, xs, ys, pi)
    ref = osc.probe(cfg, rays[:, :3], rays[:, 3:])
    got = bt.Tracer(bt.Config()).trace_segments(esc, rays[:, :3], rays[:, 3:])
    same = (got["face"] == ref["face"]) & (got["object_ref"] == ref["object_ref"])
    assert same.mean() >= 0.9999, f"{(~same).sum()} rays disagree on the hit object/face"
    hit = same & (ref["face"] >= 0)
    # FMA-contracted dot products + MUFU.RCP in the device scan: a few ulp on t, amplified by the
    # conditioning of the reference's own formulas (|oc|^2 - r^2 on the r = 100 ground sphere)
    assert np.allclose(got["t"][hit], ref["t"][hit], rtol=5e-5, atol=1e-5)
    assert np.allclose(got["position"][hit], ref["position"][hit], rtol=5e-5, atol=5e-5)
    assert np.allclose(got["normal"][hit], ref["normal"][hit], rtol=5e-5, atol=5e-6)


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("output", [0, 1, 2, 3])
def test_render_parity(oracle, name, output):
    """Tracer::render, every Output, 16 spp = 4 passes x Subpixel(2) (BASELINE config C1's sampling)"""
    w, h = _res(name)
    osc, esc, cam = load_pair(name, w, h)
    ref, n_ref, st_ref = oracle_render(osc, cam, w, h, 4, 2, output, seed=11)
    got, n_got, st_got = engine_render(esc, cam, w, h, 4, 2, output, seed=11)
    assert n_got == n_ref == 16 and int(st_got) == st_ref == 1
    assert np.array_equal(got[..., 3], ref[..., 3])          # alpha untouched (buffer.rs:159-164)
    mae = mae_per_channel(got, ref, n_ref)
    assert (mae <= IMAGE_MAE).all(), mae
    if output != 0:   # AOVs depend on the first events only: f32-rounding level
        assert (mae <= 1e-4).all(), mae


@pytest.mark.parametrize("name", SCENES)
def test_arithmetic_flavours(oracle, name):
    """The kernels exist in two arithmetic flavours (bt_scene_set_precision).  EXACT performs the
    reference's IEEE operation sequence: the image equals the oracle's to rounding of libm's sin / cos
    (MAE <= 1e-6, a handful of pixels).  FAST (MUFU reciprocal / square roots, FMA rect tests, ~1 ulp
    each) is held to the north-star bar; on the volumetric scenes ulp-level differences flip scatter
    decisions on ~5e-4 of the paths, which is why AUTO renders those with the exact flavour."""
    w, h = 128, 72
    osc, esc, cam = load_pair(name, w, h)
    ref, n, _ = oracle_render(osc, cam, w, h, 2, 2, 0, seed=2)
    esc.set_precision("exact")
    exact = engine_render(esc, cam, w, h, 2, 2, 0, seed=2)[0].copy()
    esc.set_precision("fast")
    fast = engine_render(esc, cam, w, h, 2, 2, 0, seed=2)[0].copy()
    esc.set_precision("auto")
    auto = engine_render(esc, cam, w, h, 2, 2, 0, seed=2)[0].copy()
    volumetric = name in ("cloud", "volume")
    assert (mae_per_channel(exact, ref, n) <= 1e-6).all(), mae_per_channel(exact, ref, n)
    d = np.abs(exact[..., :3] - ref[..., :3]).sum(-1) / n
    assert (d > 1e-5).mean() < 2e-3
    assert (mae_per_channel(fast, ref, n) <= IMAGE_MAE).all(), mae_per_channel(fast, ref, n)
    if not volumetric:
        assert (mae_per_channel(fast, ref, n) <= 1e-4).all(), mae_per_channel(fast, ref, n)
    assert np.array_equal(auto, exact if volumetric else fast)


def test_render_c1_cornell_512(oracle):
    """BASELINE config C1: cornell.json.gz 512x512 at 16 spp"""
    w = h = 512
    osc, esc, cam = load_pair("cornell", w, h)
    ref, n, _ = oracle_render(osc, cam, w, h, 4, 2, 0, seed=0)
    got, n2, _ = engine_render(esc, cam, w, h, 4, 2, 0, seed=0, device="cuda:0")
    assert n == n2 == 16
    mae = mae_per_channel(got, ref, n)
    assert (mae <= IMAGE_MAE).all(), mae
    # stronger than the stated bar: all but a few paths agree to f32 rounding
    rel = np.abs(got[..., :3] - ref[..., :3]) / (np.abs(ref[..., :3]) + 1.0)
    assert (rel > 1e-4).mean() < 1e-3


def test_render_contract(oracle):
    """samples == 0 -> Done and untouched; accumulate-in-place; host buffer == device buffer"""
    import bendy_tracer_b200 as bt
    w, h = 64, 48
    osc, esc, cam = load_pair("scene", w, h)
    tracer = bt.Tracer(bt.Config(), seed=5)
    buf = bt.Buffer(w, h)
    before = buf.data.copy()
    assert tracer.render(esc, cam, bt.RenderConfig.with_samples(0), buf) == bt.Status.Done
    assert np.array_equal(buf.data, before) and buf.samples() == 0
    assert tracer.render(esc, cam, bt.RenderConfig.with_samples(3), buf) == bt.Status.InProgress
    assert tracer.render(esc, cam, bt.RenderConfig.with_samples(5), buf) == bt.Status.InProgress
    assert buf.samples() == 8
    one = bt.Buffer(w, h)
    tracer.render(esc, cam, bt.RenderConfig.with_samples(8), one)
    assert np.allclose(buf.data, one.data, rtol=1e-5, atol=1e-5)     # same sample set, different f32 sum grouping
    dev = bt.Buffer(w, h, device="cuda:0")
    tracer.render(esc, cam, bt.RenderConfig.with_samples(8), dev)
    assert np.array_equal(dev.data.cpu().numpy(), one.data)           # bit-identical, deterministic
    ref, n, _ = oracle_render(osc, cam, w, h, 8, 0, 0, seed=5)
    assert (mae_per_channel(one.data, ref, 8) <= IMAGE_MAE).all()


@pytest.mark.parametrize("name,lens", [("cornell2", None), ("scene", LENS_SCENE)])
def test_host_buffer_bands(oracle, name, lens):
    """bt_render with a host buffer pipelines the frame in row bands over two streams (upload and
    download under the neighbouring band's kernel).  The image must not depend on the banding: one
    band, three ragged bands, many bands and a device-resident buffer agree bit for bit, and the
    running sums accumulate into a non-zero host buffer."""
    import bendy_tracer_b200 as bt
    w, h = 100, 70                                                  # ragged against 16-row CTAs and 8x4 tiles
    _, esc, cam = load_pair(name, w, h, lenses=lens)
    start = np.random.default_rng(3).random((h, w, 4), dtype=np.float32)
    images = {}
    for bands in ("1", "3", "64", None):
        bt.Engine.default(0).set_tuning(host_bands=None if bands is None else int(bands))   # (knobs are read per engine, not per call)
        try:
            buf = bt.Buffer(w, h)
            buf.data[...] = start
            images[bands] = engine_render(esc, cam, w, h, 3, 2, 0, seed=9, buffer=buf)[0].copy()
        finally:
            bt.Engine.default(0).set_tuning(host_bands=None)
        assert buf.samples() == 12
    for bands in ("3", "64", None):
        assert np.array_equal(images[bands], images["1"]), bands
    dev = bt.Buffer(w, h, device="cuda:0")
    dev.data.copy_(__import__("torch").from_numpy(start))
    got = engine_render(esc, cam, w, h, 3, 2, 0, seed=9, buffer=dev)[0]
    assert np.array_equal(got, images["1"])
    assert np.array_equal(images["1"][..., 3], start[..., 3])       # alpha untouched
    assert (images["1"][..., :3] != start[..., :3]).mean() > 0.5      # (a Diffuse pdf may be negative: sums can shrink)


def test_render_config_overrides(oracle):
    """RenderConfig overrides merged onto the tracer's Config (ChunkConfig::with_configs, mod.rs:218-229):
    max_bounces (which also overrides max_volume_bounces -- the :224 quirk), volume_step, output."""
    import bendy_tracer_b200 as bt
    w, h = 96, 64
    for name, kw, okw in (("cornell2", dict(max_bounces=2), dict(r_max_bounces=2)),
                          ("cloud", dict(max_bounces=3), dict(r_max_bounces=3)),          # march cut to 3 steps
                          ("cloud", dict(volume_step=0.05, max_volume_bounces=7), dict(r_volume_step=0.05, r_max_volume_bounces=7)),
                          ("scene", dict(output=bt.Output.Normal, max_bounces=0), dict(r_output=2, r_max_bounces=0))):
        osc, esc, cam = load_pair(name, w, h)
        cfg = O.make_config(samples=2, subsample=2, **okw)
        ref, n, _ = osc.render(cam, cfg, w, h, seed=31, sample_base=0)
        default, _, _ = oracle_render(osc, cam, w, h, 2, 2, 0, seed=31)
        assert not np.array_equal(ref, default)                   # the override changes the image
        esc.set_precision("exact")
        buf = bt.Buffer(w, h)
        rc = bt.RenderConfig(samples=2, subsample=bt.Subsample(2), **kw)
        bt.Tracer(bt.Config(chunks_x=8, chunks_y=4), seed=31).render(esc, cam, rc, buf)
        assert buf.samples() == n == 8
        assert (mae_per_channel(buf.data, ref, n) <= 1e-6).all(), (name, kw, mae_per_channel(buf.data, ref, n))


def test_many_lights(oracle):
    """Three LIGHT objects of different kinds (rect, sphere, and a flagged camera -> Object::random_point's
    `_ => translation` arm with pdf None): Uniform::new(0, 3) goes through the rejection zone
    (RenderParams::light_zone), every light kind is sampled and its pdf evaluated."""
    import copy
    import json
    import bendy_tracer_b200 as bt
    doc = copy.deepcopy(O.read_scene_json(O.scene_path("cornell")))
    objs = doc["objects"]["collection"]
    key = doc["objects"]["next_key"]
    tf = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, -1.0, 3.0, -1.5]
    objs[str(key)] = {"object_ref": key, "tag": None, "flags": {"bits": 1},
                      "transform": {"transform_world": tf, "transform_local": tf, "transform_parent": None},
                      "inner": {"Sphere": {"material": 1, "volume": None, "radius": 0.3}}, "children": None}
    doc["objects"]["next_key"] = key + 1
    objs["0"]["flags"]["bits"] = 1                                # the camera as a (point) light: never hit, pdf 0
    w, h = 96, 96
    osc = O.OracleScene(doc)
    esc = bt.Scene.from_json(json.dumps(doc))
    cam = esc.find_by_tag("camera")
    for s_ in (osc, esc):
        s_.set_camera_aspect(cam, 1.0)
    assert esc.info()["n_lights"] == 3
    ref, n, _ = oracle_render(osc, cam, w, h, 4, 2, 0, seed=41)
    esc.set_precision("exact")
    exact = engine_render(esc, cam, w, h, 4, 2, 0, seed=41)[0].copy()
    # a small, intense sphere light sampled without a cosine: single paths carry sums of 10..100, so one
    # path that libm's and CUDA's sincos round differently (~1e-4 of arguments) moves the mean by 1e-5 --
    # all but a handful of pixels agree to f32 rounding
    assert (mae_per_channel(exact, ref, n) <= 1e-4).all(), mae_per_channel(exact, ref, n)
    rel = np.abs(exact[..., :3] - ref[..., :3]) / (np.abs(ref[..., :3]) + 1.0)
    assert (rel > 1e-4).mean() < 2e-3, [(t, float((rel > t).mean())) for t in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2)]
    esc.set_precision("fast")
    fast = engine_render(esc, cam, w, h, 4, 2, 0, seed=41)[0]
    assert (mae_per_channel(fast, ref, n) <= IMAGE_MAE).all()


@pytest.mark.parametrize("w,h", [(1, 1), (7, 3), (17, 33)])
def test_tiny_and_ragged_frames(oracle, w, h):
    """frames smaller than one 8x4 warp tile / one 16x16 CTA, and ragged against both"""
    for name in ("cornell", "scene"):
        osc, esc, cam = load_pair(name, w, h)
        ref, n, _ = oracle_render(osc, cam, w, h, 3, 2, 0, seed=8)
        esc.set_precision("exact")
        got, n2, _ = engine_render(esc, cam, w, h, 3, 2, 0, seed=8)
        assert n == n2 == 12 and got.shape == (h, w, 4)
        assert (mae_per_channel(got, ref, n) <= 1e-6).all()
        assert np.array_equal(got[..., 3], ref[..., 3])


def test_render_errors(oracle):
    import json
    import bendy_tracer_b200 as bt
    scene = O.read_scene_json(O.scene_path("cornell"))
    for o in scene["objects"]["collection"].values():
        o["flags"]["bits"] = 0                                    # no LIGHT left: Uniform::new(0, 0) panics
    esc = bt.Scene.from_json(json.dumps(scene))
    with pytest.raises(bt.ScenePanic):
        bt.Tracer().render(esc, 0, bt.RenderConfig.with_samples(1), bt.Buffer(8, 8))
    esc = bt.Scene.load(O.scene_path("cornell"))
    with pytest.raises(bt.ScenePanic):                            # "expected a camera object"
        bt.Tracer().render(esc, 1, bt.RenderConfig.with_samples(1), bt.Buffer(8, 8))
    with pytest.raises(bt.ScenePanic):                            # "invalid object ref"
        bt.Tracer().render(esc, 99, bt.RenderConfig.with_samples(1), bt.Buffer(8, 8))


@pytest.mark.parametrize("cs", [0, 1, 2, 3])
def test_resolve_u8(oracle, cs):
    import bendy_tracer_b200 as bt
    w, h = 128, 72
    osc, esc, cam = load_pair("scene", w, h)
    buf = bt.Buffer(w, h, color_space=bt.ColorSpace(cs))
    bt.Tracer(bt.Config(output=bt.Output.Normal if cs == 1 else bt.Output.Full)).render(
        esc, cam, bt.RenderConfig.with_samples(4), buf)
    got = buf.preview()
    ref = O.resolve_u8(buf.data, buf.samples(), cs)
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 2e-3            # powf ulp at a truncation boundary
    dev = bt.Buffer(w, h, color_space=bt.ColorSpace(cs), device="cuda:0")
    dev.data.copy_(__import__("torch").from_numpy(buf.data))
    dev._samples = buf.samples()
    assert np.array_equal(dev.preview(), got)


@pytest.mark.parametrize("name", ["cornell", "cornell2"])
def test_box_slab_test_equals_six_rect_tests(oracle, name):
    """A box-shaped cuboid is intersected with one slab test; BT_ACCEL_LINEAR_FACES runs the reference's
    six Rect::hit tests per cuboid (cuboid.rs:83-105).  Same closest face for (almost) every ray --
    rays through a box edge within rounding may pick the neighbouring face -- and the same image."""
    import bendy_tracer_b200 as bt
    w, h = 256, 256
    _, esc, cam = load_pair(name, w, h)
    rng = np.random.default_rng(11)
    n = 200000
    # rays from inside the room in every direction, and from inside the boxes outwards
    origins = np.concatenate([rng.uniform([-2.4, 0.1, -4.9], [2.4, 4.9, 4.0], (n, 3)),
                              rng.normal(0, 0.1, (n // 4, 3)) + [-1.2, 1.0, -3.2],
                              rng.normal(0, 0.1, (n // 4, 3)) + [1.2, 0.5, -1.8]]).astype(np.float32)
    dirs = rng.normal(size=origins.shape)
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    tracer = bt.Tracer(bt.Config(), seed=1)
    assert esc.info()["n_boxes"] == 2
    box = tracer.trace_segments(esc, origins, dirs)
    img_box = engine_render(esc, cam, w, h, 2, 2, 0, seed=4)[0].copy()
    esc.set_accel("linear_faces")
    assert esc.info()["n_boxes"] == 0
    faces = tracer.trace_segments(esc, origins, dirs)
    img_faces = engine_render(esc, cam, w, h, 2, 2, 0, seed=4)[0].copy()
    same = (box["face"] == faces["face"]) & (box["object_ref"] == faces["object_ref"])
    assert same.mean() >= 0.9999, (~same).sum()
    on_box = same & (box["object_ref"] >= 7)
    assert on_box.mean() > 0.1                                               # the boxes really were hit
    terr = np.abs(box["t"][same] - faces["t"][same]) / (1.0 + faces["t"][same])
    assert terr.max() <= 2e-5 and np.quantile(terr, 0.999) <= 2e-6, (terr.max(), np.quantile(terr, 0.999))
    assert np.abs(box["normal"][same] - faces["normal"][same]).max() == 0.0  # the same face record
    assert (mae_per_channel(img_box, img_faces, 8) <= 1e-4).all()
    ref, nref, _ = oracle_render(load_pair(name, w, h)[0], cam, w, h, 2, 2, 0, seed=4)
    assert (mae_per_channel(img_box, ref, nref) <= IMAGE_MAE).all()


@pytest.mark.parametrize("name", ["cornell", "cornell2"])
def test_axis_aligned_rect_test_is_bit_identical(name, monkeypatch):
    """Rects whose normal and plane axes are exact coordinate axes (the Cornell walls) run
    rect_test_aa, which picks components instead of forming dot products.  It must round exactly like
    the general Rect::hit restatement: same images (both flavours) and same probe segments with the
    specialisation switched off at flatten time (BT_NO_AA_RECTS)."""
    import bendy_tracer_b200 as bt
    w, h = 160, 120
    rng = np.random.default_rng(5)
    n = 100000
    origins = rng.uniform([-2.4, 0.1, -4.9], [2.4, 4.9, -0.1], (n, 3)).astype(np.float32)
    dirs = rng.normal(size=(n, 3))
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    out = {}
    for aa in (True, False):
        if aa:
            monkeypatch.delenv("BT_NO_AA_RECTS", raising=False)
        else:
            monkeypatch.setenv("BT_NO_AA_RECTS", "1")
        _, esc, cam = load_pair(name, w, h)                      # (flattened on first use, under this setting)
        for prec in ("fast", "exact"):
            esc.set_precision(prec)
            out[aa, prec] = engine_render(esc, cam, w, h, 2, 2, 0, seed=13)[0].copy()
            out[aa, prec, "seg"] = bt.Tracer(bt.Config(), seed=1).trace_segments(esc, origins, dirs)
    for prec in ("fast", "exact"):
        assert np.array_equal(out[True, prec], out[False, prec]), prec
        a, b = out[True, prec, "seg"], out[False, prec, "seg"]
        for field in ("t", "face", "object_ref", "position", "normal"):
            assert np.array_equal(a[field], b[field]), (prec, field)
        assert ((a["object_ref"] >= 1) & (a["object_ref"] <= 6)).mean() > 0.5                  # mostly walls


# ---- lens field -----------------------------------------------------------------------------
def test_flat_limit_is_exact(oracle):
    """r_s = 0 masses must reproduce the unlensed image bit for bit"""
    w, h = 128, 72
    _, esc, cam = load_pair("scene", w, h)
    base, _, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=3)
    esc.set_lenses(np.array([[1.0, 1.0, 5.0, 0.0]], np.float32))
    flat, _, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=3)
    assert np.array_equal(base, flat)


def test_geodesic_segments_vs_f64(oracle):
    """camera rays of C3 through the lens: hit points / escape directions vs the f64 oracle.

    Rays that graze the photon sphere are exponentially sensitive (d alpha / d b ~ 1 / (b - b_c)), so
    no two floating-point implementations agree on them; the 1e-4 bar is asserted on the rays whose
    impact parameter clears the critical one by 25 % and, as a quantile, on all rays."""
    import bendy_tracer_b200 as bt
    w, h = 160, 90
    osc, esc, cam = load_pair("scene", w, h, lenses=LENS_SCENE)
    ys, xs = np.mgrid[0:h, 0:w]
    xs, ys = xs.ravel(), ys.ravel()
    cfg = O.make_config(samples=1)
    rays = osc.camera_rays(cam, cfg, w, h, xs, ys, np.zeros(len(xs), np.uint64))
    ref = osc.probe(cfg, rays[:, :3], rays[:, 3:], use_f64=True)
    got = bt.Tracer(bt.Config()).trace_segments(esc, rays[:, :3], rays[:, 3:])
    assert (got["steps"] > 0).mean() > 0.99                       # the stepper really ran
    assert (got["face"] == -2).mean() > 0.05                      # the shadow of the mass is in view
    same = (got["face"] == ref["face"]) & (got["object_ref"] == ref["object_ref"])
    assert same.mean() >= 0.995, same.mean()                       # silhouettes / capture rim may flip
    b = np.linalg.norm(np.cross(LENS_SCENE[0, :3] - rays[:, :3], rays[:, 3:]), axis=1)
    clear = b > 1.25 * 2.598 * LENS_SCENE[0, 3]
    hit = same & (ref["face"] >= 0)
    scale = np.linalg.norm(ref["position"], axis=1) + 1.0
    err = np.linalg.norm(got["position"] - ref["position"], axis=1) / scale
    # hit points also carry the f32 conditioning of the reference's sphere formula (|oc|^2 - r^2 on
    # the r = 100 ground sphere, near-tangent roots), so the bar is a quantile for them; the pure
    # stepper output (escape directions, below) is held to the bar as a maximum
    assert np.quantile(err[hit & clear], 0.999) <= ENDPOINT_REL and np.quantile(err[hit], 0.99) <= ENDPOINT_REL
    assert err[hit & clear].max() <= 10 * ENDPOINT_REL
    esc_ = same & (ref["face"] == -1)
    derr = np.linalg.norm(got["direction"] - ref["direction"], axis=1)
    assert derr[esc_ & clear].max() <= ENDPOINT_REL, derr[esc_ & clear].max()   # pure stepper: escape directions
    assert np.quantile(derr[esc_], 0.99) <= ENDPOINT_REL
    # the IEEE stepper reproduces the f32 oracle's segments exactly
    esc.set_lenses(LENS_SCENE, bt.LensConfig(exact_rsqrt=True))
    ref32 = osc.probe(cfg, rays[:, :3], rays[:, 3:], use_f64=False)
    got32 = bt.Tracer(bt.Config()).trace_segments(esc, rays[:, :3], rays[:, 3:])
    assert np.array_equal(got32["face"], ref32["face"]) and np.array_equal(got32["steps"], ref32["steps"])
    assert np.abs(got32["position"] - ref32["position"]).max() <= 1e-6 * (np.abs(ref32["position"]).max() + 1)   # sphere tests are exact


def _stepper_case(n_lens, n=8192):
    rng = np.random.default_rng(1234)
    lenses = np.zeros((n_lens, 4), np.float32)
    lenses[:, :3] = rng.uniform(-0.5, 0.5, (n_lens, 3))        # a compact cluster of masses
    lenses[0, :3] = 0
    lenses[:, 3] = 1.0 / n_lens
    b = rng.uniform(4.0, 40.0, n)                                # clear of every photon sphere
    phi = rng.uniform(0, 2 * np.pi, n)
    xv = np.zeros((n, 6), np.float32)
    xv[:, 0], xv[:, 1], xv[:, 2] = b * np.cos(phi), b * np.sin(phi), 20.0
    xv[:, 5] = -1.0
    return lenses, xv


@pytest.mark.parametrize("n_lens", [1, 4, 16])
def test_geodesic_integrate_vs_f64(oracle, n_lens):
    """the stepper kernel in isolation: 256 RK4 steps, endpoints within 1e-4 relative of f64"""
    import bendy_tracer_b200 as bt
    lenses, xv = _stepper_case(n_lens)
    ref = O.integrate(lenses, xv, 256, use_f64=True)
    for exact in (False, True):
        got = bt.Engine.default().geodesic_integrate(lenses, xv, 256, bt.LensConfig(exact_rsqrt=exact))
        scale = np.linalg.norm(ref[:, :3], axis=1) + 1.0
        err = np.linalg.norm(got[:, :3] - ref[:, :3], axis=1) / scale
        verr = np.linalg.norm(got[:, 3:] - ref[:, 3:], axis=1)
        assert err.max() <= ENDPOINT_REL and verr.max() <= ENDPOINT_REL, (exact, err.max(), verr.max())


@pytest.mark.parametrize("n_lens", [1, 4])
def test_geodesic_integrate_exact_is_bit_identical(oracle, n_lens):
    """BT_LENS_EXACT_RSQRT: every operation of the stepper is IEEE, so f32 results equal the oracle's"""
    import bendy_tracer_b200 as bt
    lenses, xv = _stepper_case(n_lens, n=4096)
    ref = O.integrate(lenses, xv, 128, use_f64=False)
    got = bt.Engine.default().geodesic_integrate(lenses, xv, 128, bt.LensConfig(exact_rsqrt=True))
    assert (got.view(np.uint32) == ref.view(np.uint32)).all(axis=1).mean() >= 0.999   # 2^-29 double-rounding cases


@pytest.mark.parametrize("name,lens", [("scene", LENS_SCENE), ("cloud", LENS_VOLUME)])
def test_render_parity_lensed_exact(oracle, name, lens):
    """lensed image parity on identical sample sets with the IEEE stepper (MAE bar 1e-3)"""
    import bendy_tracer_b200 as bt
    w, h = 128, 72
    osc, esc, cam = load_pair(name, w, h)
    osc.set_lenses(lens)
    esc.set_lenses(lens, bt.LensConfig(exact_rsqrt=True))
    ref, n, _ = oracle_render(osc, cam, w, h, 2, 2, 0, seed=2)
    got, n2, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=2)
    assert n == n2 == 8
    mae = mae_per_channel(got, ref, n)
    assert (mae <= IMAGE_MAE).all(), mae
    unl, _, _ = engine_render(load_pair(name, w, h)[1], cam, w, h, 2, 2, 0, seed=2)
    assert np.abs(unl - got).mean() > 1e-3                          # the lens really changes the image


@pytest.mark.parametrize("name,lens", [("scene", LENS_SCENE), ("cloud", LENS_VOLUME)])
def test_render_parity_lensed_fast(oracle, name, lens):
    """default stepper (MUFU.RSQ, <= 2 ulp): positions differ from the oracle by ~1e-5, which flips a
    Bernoulli decision (volume scatter, Fresnel) on ~1e-4..1e-3 of the paths; those paths are
    individually different but identically distributed.  Asserted: almost all pixels agree to f32
    rounding, and the image means agree."""
    w, h = 128, 72
    osc, esc, cam = load_pair(name, w, h, lenses=lens)
    ref, n, _ = oracle_render(osc, cam, w, h, 2, 2, 0, seed=2)
    got, _, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=2)
    d = np.abs(got[..., :3] - ref[..., :3]).sum(-1) / n
    assert (d > 1e-3).mean() < 1e-2, (d > 1e-3).mean()
    assert np.median(d) <= 1e-5
    m_got, m_ref = got[..., :3].mean() / n, ref[..., :3].mean() / n
    assert abs(m_got - m_ref) <= 0.02 * abs(m_ref) + 5e-3


@pytest.mark.parametrize("name,lens", [("scene", LENS_SCENE), ("cloud", LENS_VOLUME), ("cornell2", np.array([[0.3, 2.2, 2.0, 0.1]], np.float32))])
def test_chord_skip_changes_nothing(name, lens):
    """The stepper intersects a chord only when it is at least as long as the free distance around
    its start (a conservative bound refreshed by every intersection pass).  BT_LENS_NO_SKIP tests
    every chord, as the spec's loop is written: images and probe segments must be bit-identical."""
    import bendy_tracer_b200 as bt
    w, h = 192, 108
    _, esc, cam = load_pair(name, w, h)
    imgs, segs = [], []
    ys, xs = np.mgrid[0:h, 0:w]
    for no_skip in (False, True):
        esc.set_lenses(lens, bt.LensConfig(no_skip=no_skip))
        imgs.append(engine_render(esc, cam, w, h, 2, 2, 0, seed=5)[0].copy())
        tracer = bt.Tracer(bt.Config(), seed=5)
        rays = tracer.camera_rays(esc, cam, bt.RenderConfig.with_samples(1), w, h, xs.ravel(), ys.ravel(),
                                  np.zeros(w * h, np.uint64))
        segs.append(tracer.trace_segments(esc, rays[:, :3], rays[:, 3:]))
    assert np.array_equal(imgs[0], imgs[1])
    for k in ("face", "steps", "object_ref", "t", "position", "normal", "direction"):
        assert np.array_equal(segs[0][k], segs[1][k]), k
    assert (segs[0]["steps"] > 10).mean() > 0.5                      # long flights: the skip had work to do


# ---- full-size properties (BASELINE configs C2 / C4) ------------------------------------------
def test_full_size_properties():
    """1920x1080: determinism, pass-range additivity (the multi-GPU sharding law), alpha, finiteness"""
    import bendy_tracer_b200 as bt
    w, h = 1920, 1080
    esc = bt.Scene.load(O.scene_path("cornell2"))
    cam = esc.find_by_tag("camera")
    esc.set_camera_aspect(cam, w / h)
    tracer = bt.Tracer(bt.Config(), seed=1)
    rc = bt.RenderConfig.with_samples_subsample(8, bt.Subsample(2))
    a = bt.Buffer(w, h, device="cuda:0")
    b = bt.Buffer(w, h, device="cuda:0")
    tracer.render(esc, cam, rc, a)
    tracer.render(esc, cam, rc, b)
    assert bool((a.data == b.data).all())                            # deterministic
    assert bool((a.data[..., 3] == 1.0).all()) and bool(a.data.isfinite().all())
    c = bt.Buffer(w, h, device="cuda:0")
    half = bt.RenderConfig.with_samples_subsample(4, bt.Subsample(2))
    tracer.render(esc, cam, half, c, sample_base=0)
    tracer.render(esc, cam, half, c, sample_base=4)
    assert c.samples() == a.samples() == 32
    assert float((a.data - c.data).abs().max()) <= 1e-3 * float(a.data[..., :3].abs().max())
    m = (a.data[..., :3].mean(dim=(0, 1)) / a.samples()).cpu().numpy()
    assert (m > 0.05).all() and (m < 1.0).all()                      # a lit Cornell box


def test_full_size_properties_c3_c4():
    """BASELINE configs C3 (scene + lens, 3840x2160) and C4 (cloud, 1920x1080) at full frame size and
    reduced spp: determinism, pass-range additivity, the r_s = 0 control (bit-identical to the unlensed
    frame), a host frame through the band pipeline == the device-resident frame, finite sums."""
    import torch
    import bendy_tracer_b200 as bt
    for name, w, h, lens in (("scene", 3840, 2160, LENS_SCENE), ("cloud", 1920, 1080, None)):
        esc = bt.Scene.load(O.scene_path(name))
        cam = esc.find_by_tag("camera")
        esc.set_camera_aspect(cam, w / h)
        tracer = bt.Tracer(bt.Config(), seed=2)
        rc = bt.RenderConfig.with_samples_subsample(2, bt.Subsample(2))
        one = bt.RenderConfig.with_samples_subsample(1, bt.Subsample(2))
        flat = bt.Buffer(w, h, device="cuda:0")
        tracer.render(esc, cam, rc, flat)
        if lens is not None:
            esc.set_lenses(np.array([[1.0, 1.0, 5.0, 0.0]], np.float32))   # a mass without mass
            ctrl = bt.Buffer(w, h, device="cuda:0")
            tracer.render(esc, cam, rc, ctrl)
            assert bool((ctrl.data == flat.data).all())
            esc.set_lenses(lens)
        a = bt.Buffer(w, h, device="cuda:0")
        b = bt.Buffer(w, h, device="cuda:0")
        tracer.render(esc, cam, rc, a)
        tracer.render(esc, cam, rc, b)
        assert bool((a.data == b.data).all()) and bool(a.data.isfinite().all()) and bool((a.data[..., 3] == 1.0).all())
        if lens is not None:
            assert float((a.data - flat.data).abs().mean()) > 1e-3       # the lens really bends the image
        c = bt.Buffer(w, h, device="cuda:0")
        tracer.render(esc, cam, one, c, sample_base=0)
        tracer.render(esc, cam, one, c, sample_base=1)
        assert c.samples() == a.samples() == 8
        assert float((a.data - c.data).abs().max()) <= 1e-3 * float(a.data[..., :3].abs().max())
        host = bt.Buffer(w, h)                                           # 133 MB at 4K: 8 bands over two streams
        tracer.render(esc, cam, rc, host)
        assert np.array_equal(host.data, a.data.cpu().numpy())
        del a, b, c, flat
        torch.cuda.empty_cache()


# ---- BVH (scenes above the linear-scan budget) -------------------------------------------------
@pytest.mark.parametrize("name,lens", [("cornell", None), ("scene", None), ("scene", LENS_SCENE)])
def test_bvh_equals_linear_scan_on_shipped_scenes(name, lens):
    """Forcing the BVH on a shipped scene must not change a single bit of the image in the exact
    flavour (against the scan with the same per-face rect tests; the box slab test of the default scan
    has its own test).  The fast flavour is built with FMA contraction, and the BVH kernel and the
    content-specialised scan kernel are separate compilations of the shading code: there the two
    images agree to rounding (1e-4 of the bar) rather than bit for bit."""
    w, h = _res(name)
    _, esc, cam = load_pair(name, w, h, lenses=lens)
    for precision in ("exact", "fast"):
        esc.set_precision(precision)
        esc.set_accel("linear_faces")
        a = engine_render(esc, cam, w, h, 2, 2, 0, seed=4)[0].copy()
        esc.set_accel("bvh")
        assert esc.info()["n_bvh_nodes"] > 0
        b = engine_render(esc, cam, w, h, 2, 2, 0, seed=4)[0].copy()
        if precision == "exact":
            assert np.array_equal(a, b)
        else:
            d = np.abs(a[..., :3] - b[..., :3]).sum(-1) / 8
            assert (d > 1e-3).mean() < 5e-3 and np.median(d) <= 1e-6, ((d > 1e-3).mean(), np.median(d))


def test_bvh_deep_narrow_tree_equals_linear_scan():
    """Coincident spheres (leaves of up to 31 records, exact-distance ties on every ray that meets them) and a geometric chain of disjoint spheres (every SAH split
    lopsided: a deep, narrow tree): BVH == linear scan bit for bit in the exact flavour -- through the lane kernel and
    the pooled traversal, with the stack in shared memory and with all but one level of it in its slow tail."""
    import json
    import bendy_tracer_b200 as bt
    from common import skewed_scene
    w, h = 128, 72
    for chain_radius in (0.0, 0.3):
        esc = bt.Scene.from_json(json.dumps(skewed_scene(120, 60, chain_radius)))
        cam = esc.find_by_tag("camera")
        esc.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
        esc.set_precision("exact")
        esc.set_accel("linear_faces")
        ref = engine_render(esc, cam, w, h, 2, 2, 0, seed=6)[0].copy()
        assert np.isfinite(ref).all() and ref[..., :3].sum() > 0
        esc.set_accel("bvh")
        assert esc.info()["n_bvh_nodes"] > 3
        eng = bt.Engine.default(0)
        for knobs in ({}, {"pool_w": 0}, {"bvh_stack_k": 1}, {"pool_w": 0, "bvh_stack_k": 1}):
            eng.set_tuning(**knobs)
            try:
                got = engine_render(esc, cam, w, h, 2, 2, 0, seed=6)[0].copy()
            finally:
                eng.set_tuning(**{k: None for k in knobs})
            assert np.array_equal(got, ref), (chain_radius, knobs, int((got != ref).any(-1).sum()))


def test_bvh_synthetic_scene_vs_oracle(oracle):
    """a 500-primitive scene: BVH render == linear-scan render bit for bit, and both match the oracle"""
    import json
    import bendy_tracer_b200 as bt
    from common import synthetic_scene
    scene = synthetic_scene(200, 100, 20, seed=5)          # 2 + 200 + 100 + 120 = 422 flattened primitives
    w, h = 128, 72
    osc = O.OracleScene(scene)
    esc = bt.Scene.from_json(json.dumps(scene))
    cam = esc.find_by_tag("camera")
    for s in (osc, esc):
        s.set_camera_aspect(cam, w / h)
    info = esc.info()
    assert info["n_primitives"] == 422 and info["n_bvh_nodes"] > 30      # automatic above 64 primitives
    ref, n, _ = oracle_render(osc, cam, w, h, 2, 2, 0, seed=6)
    got, _, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=6)
    mae = mae_per_channel(got, ref, n)
    assert (mae <= IMAGE_MAE).all(), mae
    esc.set_precision("exact")                             # bit equality across structures: the exact flavour
    bvh_exact = engine_render(esc, cam, w, h, 2, 2, 0, seed=6)[0].copy()
    esc.set_accel("linear_faces")
    lin, _, _ = engine_render(esc, cam, w, h, 2, 2, 0, seed=6)
    assert np.array_equal(lin, bvh_exact)
    assert (mae_per_channel(bvh_exact, ref, n) <= 1e-5).all()
    esc.set_precision("auto")
    # first hits agree with the oracle's scan
    ys, xs = np.mgrid[0:h, 0:w]
    cfg = O.make_config(samples=1)
    rays = osc.camera_rays(cam, cfg, w, h, xs.ravel(), ys.ravel(), np.zeros(w * h, np.uint64))
    r = osc.probe(cfg, rays[:, :3], rays[:, 3:])
    esc.set_accel("bvh")
    g = bt.Tracer(bt.Config()).trace_segments(esc, rays[:, :3], rays[:, 3:])
    assert ((g["face"] == r["face"]) & (g["object_ref"] == r["object_ref"])).mean() >= 0.9995


@pytest.mark.parametrize("two_lights", [True, False])
def test_cuboid_light(oracle, two_lights):
    """A Cuboid with ObjectFlags::LIGHT: WeightedIndex face pick by area + Rect::random_point
    (cuboid.rs:48-54) and Cuboid::pdf (cuboid.rs:56-81), alone and beside the ceiling rect light
    (Uniform::new(0, 2) light index).  The exact flavour equals the oracle to libm rounding under the
    box slab test, the six literal rect tests and the BVH; the fast flavour meets the image bar."""
    import json
    import bendy_tracer_b200 as bt
    from common import cornell_with_cuboid_light
    doc = cornell_with_cuboid_light(two_lights)
    w, h = 96, 96
    osc = O.OracleScene(doc)
    esc = bt.Scene.from_json(json.dumps(doc))
    cam = esc.find_by_tag("camera")
    for s in (osc, esc):
        s.set_camera_aspect(cam, 1.0)
    assert esc.info()["n_lights"] == (2 if two_lights else 1)
    ref, n, _ = oracle_render(osc, cam, w, h, 4, 2, 0, seed=21)
    plain = oracle_render(load_pair("cornell", w, h)[0], cam, w, h, 4, 2, 0, seed=21)[0]
    assert mae_per_channel(ref, plain, n).max() > 1e-2            # the box light really changes the image
    fast = engine_render(esc, cam, w, h, 4, 2, 0, seed=21)[0].copy()
    assert (mae_per_channel(fast, ref, n) <= IMAGE_MAE).all(), mae_per_channel(fast, ref, n)
    esc.set_precision("exact")
    images = {}
    for accel in ("auto", "linear_faces", "bvh"):
        esc.set_accel(accel)
        images[accel] = engine_render(esc, cam, w, h, 4, 2, 0, seed=21)[0].copy()
        mae = mae_per_channel(images[accel], ref, n)
        assert (mae <= 1e-6).all(), (accel, mae)
    assert np.array_equal(images["linear_faces"], images["bvh"])   # same per-face tests, same tie rule
    for output in (1, 2, 3):                                       # AOVs through the generic kernel
        esc.set_accel("auto")
        r = oracle_render(osc, cam, w, h, 2, 2, output, seed=22)[0]
        g = engine_render(esc, cam, w, h, 2, 2, output, seed=22)[0]
        assert (mae_per_channel(g, r, 8) <= 1e-6).all()
    # WeightedIndex::new(all-zero areas).unwrap() panics (cuboid.rs:49)
    for _, rect in doc["objects"]["collection"]["8"]["inner"]["Cuboid"]["faces"]:
        rect["half_width"] = 0.0
    bad = bt.Scene.from_json(json.dumps(doc))
    with pytest.raises(bt.ScenePanic):
        bt.Tracer().render(bad, cam, bt.RenderConfig.with_samples(1), bt.Buffer(8, 8))


def test_cli_progressive_render_and_png(tmp_path):
    """csrc/bendy_b200_cli: 1 pass per iteration until --samples, PNG screenshot == Buffer::preview of the
    same progressive sequence through the Python mirror"""
    import os
    import subprocess
    from PIL import Image
    import bendy_tracer_b200 as bt
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "bendy_tracer_b200", "csrc", "bendy_b200_cli")
    png = tmp_path / "shots" / "render.png"
    r = subprocess.run([cli, "--output", "full", "--width", "96", "--height", "64", "--samples", "16", "--subsample", "2",
                        "--scene", O.scene_path("scene"), "--screenshot", str(png), "--seed", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "16/16 samples" in r.stderr and "saved screenshot" in r.stderr
    img = np.asarray(Image.open(png))
    assert img.shape == (64, 96, 4) and (img[..., 3] == 255).all() and img[..., :3].mean() > 10
    scene = bt.Scene.load(O.scene_path("scene"))
    cam = scene.find_by_tag("camera")
    scene.set_camera_aspect(cam, float(np.float32(96) / np.float32(64)))
    buf = bt.Buffer(96, 64, bt.ColorSpace.SRgb)
    tracer = bt.Tracer(bt.Config(chunks_x=8, chunks_y=4), seed=3)
    while buf.samples() < 16:                                      # main.rs:245-254
        tracer.render(scene, cam, bt.RenderConfig.with_samples_subsample(1, bt.Subsample(2)), buf)
    assert np.array_equal(buf.preview(), img)


@pytest.mark.parametrize("name", SCENES)
def test_render_upstream_scene_bytes(oracle, name):
    """Tracer::render on the reference's shipped scene files loaded byte for byte (tests/golden/scenes_upstream): the image
    is the one the re-serialised copy gives (same bits) and matches the oracle, which reads the same bytes"""
    import os

    import bendy_tracer_b200 as bt
    w, h = _res(name)
    raw = open(os.path.join(os.path.dirname(__file__), "golden", "scenes_upstream", name + ".json.gz"), "rb").read()
    up = bt.Scene(raw)
    cam = up.find_by_tag("camera")
    aspect = float(np.float32(w) / np.float32(h))
    up.set_camera_aspect(cam, aspect)
    got = engine_render(up, cam, w, h, 2, 2, 0, seed=5)[0].copy()
    osc, esc, _ = load_pair(name, w, h)
    same = engine_render(esc, cam, w, h, 2, 2, 0, seed=5)[0]
    assert np.array_equal(got, same)
    ousc = O.OracleScene.load(os.path.join(os.path.dirname(__file__), "golden", "scenes_upstream", name + ".json.gz"))
    ousc.set_camera_aspect(cam, aspect)
    ref, n, _ = oracle_render(ousc, cam, w, h, 2, 2, 0, seed=5)
    assert (mae_per_channel(got, ref, n) <= IMAGE_MAE).all()


@pytest.mark.parametrize("name,lens", [("scene", LENS_SCENE), ("cloud", LENS_VOLUME), ("cornell2", np.array([[0.3, 2.2, 2.0, 0.1]], np.float32))])
def test_render_parity_lensed_default_path_mae(oracle, name, lens):
    """The north-star bar on the path that C3 / C5 actually run and bench.py times: the DEFAULT lensed configuration
    (MUFU.RSQ stepper, arithmetic flavour AUTO, pooled kernel, free-distance grid) against the oracle on identical sample
    sets at 64 spp -- per-channel mean absolute error of the resolved images <= 1e-3.  (A flipped Bernoulli decision
    makes one of a pixel's 64 paths differ; the per-path statements are in test_render_parity_lensed_fast / _exact.)"""
    w, h = 96, 54
    osc, esc, cam = load_pair(name, w, h, lenses=lens)
    ref, n, _ = oracle_render(osc, cam, w, h, 16, 2, 0, seed=6)
    got, n_got, _ = engine_render(esc, cam, w, h, 16, 2, 0, seed=6, device="cuda:0")
    assert n == n_got == 64
    mae = mae_per_channel(got, ref, n)
    assert (mae <= IMAGE_MAE).all(), mae
    rel = np.abs(got[..., :3].mean((0, 1)) - ref[..., :3].mean((0, 1))) / ref[..., :3].mean((0, 1))
    assert (rel <= 5e-3).all(), rel                 # the frame means agree to half a per cent


@pytest.mark.parametrize("config,name,w,h,lens", [("C2", "cornell2", 1920, 1080, None), ("C4", "cloud", 1920, 1080, None),
                                                   ("C4", "volume", 1920, 1080, None), ("C3", "scene", 3840, 2160, LENS_SCENE)])
def test_full_size_frames_vs_oracle(oracle, config, name, w, h, lens):
    """BASELINE configs C2 / C3 / C4 at their FULL frame sizes, 4 spp (one pass x Subpixel(2)), against the oracle on the
    same sample set: per-channel MAE <= 1e-3 (the C3 frame is 8.3 Mpixel: ~30 s of oracle time on 16 host threads)."""
    import os
    osc, esc, cam = load_pair(name, w, h, lenses=lens)
    cfg = O.make_config(samples=1, subsample=2)
    ref, n, _ = osc.render(cam, cfg, w, h, seed=0, n_threads=os.cpu_count() or 1)
    got, n_got, _ = engine_render(esc, cam, w, h, 1, 2, 0, seed=0, device="cuda:0")
    assert n == n_got == 4 and got.shape == ref.shape == (h, w, 4)
    mae = mae_per_channel(got, ref, n)
    assert (mae <= IMAGE_MAE).all(), (config, mae)
    assert np.array_equal(got[..., 3], ref[..., 3])
    d = np.abs(got[..., :3] - ref[..., :3]).sum(-1) / n
    assert (d > 1e-3).mean() < (2e-2 if lens is not None or name in ("cloud", "volume") else 1e-3), (d > 1e-3).mean()
