"""Scene-graph updates between renders (reference src/scene/mod.rs:154-239, object/mod.rs:200-236):
bt_scene_apply_transform + bt_scene_commit rewrite the moved objects' records in place, refit the BVH and upload only the
ranges that changed.  The image must equal the one of a scene flattened from scratch in the same state."""
import json
import time

import numpy as np
import pytest

from common import LENS_SCENE, synthetic_scene

pytestmark = pytest.mark.gpu


def _render(sc, cam, w, h, passes=2, seed=8):
    import bendy_tracer_b200 as bt
    buf = bt.Buffer(w, h, device="cuda:0")
    bt.Tracer(bt.Config(), seed=seed).render(sc, cam, bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(2)), buf)
    return buf.data.cpu().numpy()


def _fresh_copy(sc, lens=None, accel=None):
    """the same scene state through the wire format: flattened from scratch"""
    import bendy_tracer_b200 as bt
    fresh = bt.Scene.from_json(sc.to_json())
    if lens is not None:
        fresh.set_lenses(lens)
    if accel:
        fresh.set_accel(accel)
    return fresh


def _translate(dx, dy, dz):
    return [1, 0, 0, 0, 1, 0, 0, 0, 1, dx, dy, dz]


def _rot_y(deg, t=(0, 0, 0)):
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    return [c, 0, -s, 0, 1, 0, s, 0, c, *t]


@pytest.mark.parametrize("name,lens,edits", [
    ("cornell", None, [(8, _translate(0.4, 0.0, 0.3)), (7, _rot_y(15.0))]),          # a box-shaped cuboid moves, another turns
    ("cornell2", None, [(6, _translate(-0.5, 0.0, 0.5))]),                           # the ceiling light (LIGHT rect) moves
    ("scene", None, [(3, _translate(-2.0, 1.0, 1.0)), (4, _translate(0.5, 0.2, 0.0))]),  # the sphere light and the glass ball
    ("scene", LENS_SCENE, [(5, _translate(0.0, 0.5, 1.0))]),                         # under a lens field: the grid is rebuilt
    ("cloud", None, [(2, _translate(0.3, 0.1, 0.0))]),                               # the volumetric sphere
])
def test_transform_update_equals_fresh_flatten(name, lens, edits):
    import bendy_tracer_b200 as bt
    import oracle_ffi as O
    w, h = 128, 72
    sc = bt.Scene.load(O.scene_path(name))
    cam = sc.find_by_tag("camera")
    sc.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    if lens is not None:
        sc.set_lenses(lens)
    before = _render(sc, cam, w, h)                       # flatten + upload the original state
    for ref, affine in edits:
        sc.apply_transform(ref, affine)
    sc.commit()
    after = _render(sc, cam, w, h)                        # in place: rewritten records, partial upload
    fresh = _fresh_copy(sc, lens)
    fresh.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    assert np.array_equal(after, _render(fresh, cam, w, h))
    assert not np.array_equal(after, before)
    for ref, affine in edits:                              # edits applied lazily (no commit call) behave the same
        sc.apply_transform(ref, affine)
    again = _render(sc, cam, w, h)
    fresh2 = _fresh_copy(sc, lens)
    fresh2.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    assert np.array_equal(again, _render(fresh2, cam, w, h))


def test_bvh_refit_32k_primitives_is_fast_and_exact():
    """32k flattened primitives under the BVH: one object moves -> records rewritten, the BVH REFIT (not rebuilt), 1 MB of nodes
    + 80 B of records uploaded.  Host time of the update is a small fraction of a rebuild; the image equals a rebuilt scene's."""
    import bendy_tracer_b200 as bt
    doc = synthetic_scene(20000, 6000, 1000, seed=1, extent=14.0)
    sc = bt.Scene.from_json(json.dumps(doc))
    cam = sc.find_by_tag("camera")
    w, h = 192, 108
    sc.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    assert sc.info()["n_bvh_nodes"] > 1000 and sc.info()["n_primitives"] == 32002
    _render(sc, cam, w, h, passes=1)
    t0 = time.perf_counter()
    rebuilt = bt.Scene.from_json(json.dumps(doc))
    rebuilt.info()
    t_build = time.perf_counter() - t0
    moved = [100, 20500, 26100]                           # a sphere, a rect, a cuboid
    t_update = 0.0
    for ref in moved:
        sc.apply_transform(ref, _translate(0.3, 0.2, -0.1))
        t0 = time.perf_counter()
        sc.commit()
        t_update += time.perf_counter() - t0
    t0 = time.perf_counter()
    got = _render(sc, cam, w, h, passes=1)
    t_first = time.perf_counter() - t0
    fresh = _fresh_copy(sc)
    fresh.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    ref_img = _render(fresh, cam, w, h, passes=1)
    assert np.array_equal(got, ref_img)
    print(f"rebuild {t_build * 1e3:.1f} ms; refit update {t_update / len(moved) * 1e3:.2f} ms per edit; first render after it {t_first * 1e3:.1f} ms")
    assert t_update / len(moved) < 0.25 * t_build
    assert t_update / len(moved) < 5e-3                   # milliseconds, not a rebuild


def test_update_that_changes_the_layout_falls_back():
    """a non-rigid transform turns a box-shaped cuboid into six free rects (no BOX record any more): flattened from scratch"""
    import bendy_tracer_b200 as bt
    import oracle_ffi as O
    w, h = 96, 96
    sc = bt.Scene.load(O.scene_path("cornell"))
    cam = sc.find_by_tag("camera")
    assert sc.info()["n_boxes"] == 2
    _render(sc, cam, w, h)
    sc.apply_transform(8, [1, 0.3, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0])   # a shear
    got = _render(sc, cam, w, h)
    assert sc.info()["n_boxes"] == 1
    fresh = _fresh_copy(sc)
    assert np.array_equal(got, _render(fresh, cam, w, h))


@pytest.mark.parametrize("name,lens", [("scene", LENS_SCENE), ("cornell2", np.array([[0.3, 2.2, 2.0, 0.1]], np.float32))])
def test_free_distance_grid_device_build_equals_host_build(name, lens, monkeypatch):
    """the free-distance grid is filled by a kernel from the uploaded scene (one thread per cell); BT_DIST_GRID_HOST=1 fills it
    on the host with the same arithmetic.  Same grid => the same chords are skipped: identical work counters (intersection
    passes, RK4 steps) and identical images -- before and after a transform edit (which re-plans and refills the grid)."""
    import bendy_tracer_b200 as bt
    import oracle_ffi as O
    w, h = 128, 72
    out = {}
    for mode in ("device", "host"):
        if mode == "host":
            monkeypatch.setenv("BT_DIST_GRID_HOST", "1")
        else:
            monkeypatch.delenv("BT_DIST_GRID_HOST", raising=False)
        sc = bt.Scene.load(O.scene_path(name))
        cam = sc.find_by_tag("camera")
        sc.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
        sc.set_lenses(lens)
        tr = bt.Tracer(bt.Config(), seed=8)
        rc = bt.RenderConfig.with_samples_subsample(2, bt.Subsample(2))
        img0, st0 = _render(sc, cam, w, h), tr.render_stats(sc, cam, rc, w, h)
        sc.apply_transform(2, _translate(0.2, 0.1, -0.3))
        t0 = time.perf_counter()
        sc.commit()
        t_commit = time.perf_counter() - t0
        img1, st1 = _render(sc, cam, w, h), tr.render_stats(sc, cam, rc, w, h)
        out[mode] = (img0, st0, img1, st1, t_commit)
    for k in (0, 2):
        assert np.array_equal(out["device"][k], out["host"][k])
    assert out["device"][1] == out["host"][1] and out["device"][3] == out["host"][3]
    assert out["device"][1]["scans"] < 0.2 * out["device"][1]["rk4_steps"]           # (and the grid does skip chords: most RK4 steps need no intersection pass)
    print(f"{name}: commit of one moved object under a lens field: grid on the device {out['device'][4] * 1e3:.2f} ms, on the host {out['host'][4] * 1e3:.1f} ms")
    assert out["device"][4] < 0.2 * out["host"][4]
