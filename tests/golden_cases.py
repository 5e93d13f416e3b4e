"""The cases of tests/golden/oracle_vectors.npz, shared by tools/make_golden.py and tests/test_golden.py."""
import numpy as np

import oracle_ffi as O
from common import LENS_SCENE, LENS_VOLUME

W, H = 64, 36


def _case(scene, samples=1, subsample=2, output=0, seed=7, lens=None):
    return dict(scene=scene, samples=samples, subsample=subsample, output=output, seed=seed, lens=lens)


CASES = {
    "cornell": _case("cornell"),
    "cornell2": _case("cornell2"),
    "scene": _case("scene"),
    "volume": _case("volume"),
    "cloud": _case("cloud"),
    "cornell_none_subsample": _case("cornell", samples=3, subsample=0, seed=11),
    "cornell_albedo": _case("cornell", output=1),
    "cornell_normal": _case("cornell", output=2),
    "cornell_depth": _case("cornell", output=3),
    "scene_lens": _case("scene", lens=LENS_SCENE),
    "cloud_lens": _case("cloud", lens=LENS_VOLUME),
}


def load_oracle_case(case):
    osc = O.OracleScene.load(O.scene_path(case["scene"]))
    cam = osc.find_by_tag("camera")
    osc.set_camera_aspect(cam, float(np.float32(W) / np.float32(H)))
    if case["lens"] is not None:
        osc.set_lenses(case["lens"])
    return osc, cam


def probe_rays():
    """one camera ray per pixel (path index 0)"""
    ys, xs = np.mgrid[0:H, 0:W]
    return xs.ravel().astype(np.uint32), ys.ravel().astype(np.uint32), np.zeros(W * H, np.uint64)


def stepper_case(n_lens, n=512):
    rng = np.random.default_rng(99)
    lenses = np.zeros((n_lens, 4), np.float32)
    lenses[:, :3] = rng.uniform(-0.5, 0.5, (n_lens, 3))
    lenses[0, :3] = 0
    lenses[:, 3] = 1.0 / n_lens
    b = rng.uniform(4.0, 40.0, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    xv = np.zeros((n, 6), np.float32)
    xv[:, 0], xv[:, 1], xv[:, 2] = b * np.cos(phi), b * np.sin(phi), 20.0
    xv[:, 5] = -1.0
    return lenses, xv
