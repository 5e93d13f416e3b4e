"""Shared helpers of the parity tests: load the same scene into the oracle and the engine."""
import numpy as np

import oracle_ffi as O

SCENES = ["cornell", "cornell2", "scene", "volume", "cloud"]

# C3/C4 synthetic lens: a mass 6 units in front of the camera of scene.json.gz
LENS_SCENE = np.array([[1.362, 1.577, 6.114, 0.2]], np.float32)
LENS_VOLUME = np.array([[2.4 - 6 * 0.1705, 2.7 - 6 * 0.1908, 12.0 - 6 * 0.9667, 0.2]], np.float32)


def load_pair(name, width, height, lenses=None):
    """(oracle scene, engine scene, camera ref) with camera.aspect_ratio = width / height (main.rs:218-223)"""
    import bendy_tracer_b200 as bt
    path = O.scene_path(name)
    osc = O.OracleScene.load(path)
    esc = bt.Scene.load(path)
    cam = esc.find_by_tag("camera")
    assert cam == osc.find_by_tag("camera")
    aspect = float(np.float32(width) / np.float32(height))
    osc.set_camera_aspect(cam, aspect)
    esc.set_camera_aspect(cam, aspect)
    if lenses is not None:
        osc.set_lenses(lenses)
        esc.set_lenses(lenses)
    return osc, esc, cam


def oracle_render(osc, cam, width, height, samples, subsample=0, output=0, seed=0, sample_base=0, **cfg):
    c = O.make_config(samples=samples, subsample=subsample, output=output, **cfg)
    buf, n, status = osc.render(cam, c, width, height, seed=seed, sample_base=sample_base)
    return buf, n, status


def engine_render(esc, cam, width, height, samples, subsample=0, output=0, seed=0, sample_base=0, device=None,
                  buffer=None, **cfg):
    import bendy_tracer_b200 as bt
    tracer = bt.Tracer(bt.Config(chunks_x=8, chunks_y=4, output=bt.Output(output), **cfg), seed=seed)
    if buffer is None:
        buffer = bt.Buffer(width, height, device=device)
    rc = bt.RenderConfig.with_samples_subsample(samples, bt.Subsample(subsample))
    status = tracer.render(esc, cam, rc, buffer, sample_base=sample_base)
    data = buffer.data if isinstance(buffer.data, np.ndarray) else buffer.data.cpu().numpy()
    return data, buffer.samples(), status


def mae_per_channel(a, b, n):
    """per-channel mean absolute error of the resolved (sum / samples) images"""
    return np.abs(a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)).mean(axis=(0, 1)) / n
