"""The multi-device engine behind the C ABI (bt_engine_create_multi): Tracer::render's internal fan-out
(reference src/tracer/mod.rs:190-197) as a pass split across GPUs + ONE framebuffer reduce over peer memory.

A device may be listed more than once, so the sharding, the slice frames, the peer-sum kernel and the event
ordering are exercised on a one-GPU box too; the tests that need real peers skip below two devices.
The union of the slices is the one-device sample set (RNG keyed by the global pass index): the images agree up
to f32 summation order.
"""
import numpy as np
import pytest

from common import LENS_SCENE

pytestmark = pytest.mark.gpu


def _setup(name, w, h, lens=None):
    import bendy_tracer_b200 as bt
    import oracle_ffi as O
    sc = bt.Scene.load(O.scene_path(name))
    cam = sc.find_by_tag("camera")
    sc.set_camera_aspect(cam, float(np.float32(w) / np.float32(h)))
    if lens is not None:
        sc.set_lenses(lens)
    return sc, cam


def _render(engine, sc, cam, w, h, passes, sub, device, sample_base=0, start=None, seed=4):
    import bendy_tracer_b200 as bt
    buf = bt.Buffer(w, h, device=device)
    if start is not None:
        if device is None:
            buf.data[...] = start
        else:
            import torch
            buf.data.copy_(torch.from_numpy(start))
    tr = bt.Tracer(bt.Config(), engine=engine, seed=seed)
    st = tr.render(sc, cam, bt.RenderConfig.with_samples_subsample(passes, bt.Subsample(sub)), buf, sample_base=sample_base)
    data = buf.data if device is None else buf.data.cpu().numpy()
    return np.array(data), buf.samples(), st


@pytest.mark.parametrize("name,lens", [("cornell2", None), ("scene", LENS_SCENE), ("cloud", None)])
@pytest.mark.parametrize("n", [2, 3])
@pytest.mark.parametrize("device", ["cuda:0", None])
def test_multi_engine_equals_single(name, lens, n, device):
    import bendy_tracer_b200 as bt
    w, h, passes = 100, 60, 7                       # 7 passes over 2 / 3 devices: ragged slices
    start = np.random.default_rng(1).random((h, w, 4), dtype=np.float32)
    ref, n_ref, st_ref = _render(bt.Engine(0), *_setup(name, w, h, lens), w, h, passes, 2, device, sample_base=3, start=start)
    multi = bt.Engine(devices=[0] * n)
    got, n_got, st_got = _render(multi, *_setup(name, w, h, lens), w, h, passes, 2, device, sample_base=3, start=start)
    assert n_got == n_ref == passes * 4 and st_got == st_ref == bt.Status.InProgress
    assert np.array_equal(got[..., 3], start[..., 3])                       # alpha untouched
    assert np.allclose(got[..., :3], ref[..., :3], rtol=2e-6, atol=2e-5)    # f32 summation order only
    assert multi.launch_count >= n + 1                                      # n slices + the reduce


def test_multi_engine_more_devices_than_passes_and_zero_samples():
    import bendy_tracer_b200 as bt
    w, h = 64, 32
    multi, single = bt.Engine(devices=[0, 0, 0, 0]), bt.Engine(0)
    sc, cam = _setup("cornell", w, h)                                       # (a scene binds to the first engine that renders it)
    got, n, st = _render(multi, sc, cam, w, h, 2, 0, "cuda:0")              # 2 passes over 4 devices: two render nothing
    ref, _, _ = _render(single, *_setup("cornell", w, h), w, h, 2, 0, "cuda:0")
    assert n == 2 and np.allclose(got, ref, rtol=2e-6, atol=2e-5)
    got, n, st = _render(multi, sc, cam, w, h, 0, 0, "cuda:0")              # samples == 0: Done, nothing touched (mod.rs:186-188)
    assert st == bt.Status.Done and n == 0 and (got[..., :3] == 0).all()


def test_multi_engine_progressive_calls():
    """two calls into one buffer == one call of twice the passes (the ev_reduced / ev_slice ordering across calls)"""
    import bendy_tracer_b200 as bt
    w, h = 96, 48
    sc, cam = _setup("scene", w, h, LENS_SCENE)
    multi = bt.Engine(devices=[0, 0])
    buf = bt.Buffer(w, h, device="cuda:0")
    tr = bt.Tracer(bt.Config(), engine=multi, seed=2)
    rc = bt.RenderConfig.with_samples_subsample(3, bt.Subsample(2))
    tr.render(sc, cam, rc, buf, sync=False)
    tr.render(sc, cam, rc, buf)
    once, _, _ = _render(bt.Engine(0), *_setup("scene", w, h, LENS_SCENE), w, h, 6, 2, "cuda:0", seed=2)
    assert buf.samples() == 24
    assert np.allclose(buf.data.cpu().numpy(), once, rtol=2e-6, atol=2e-5)


def test_multi_engine_on_real_peers():
    """every GPU of the box, the slices read over NVLink peer mappings"""
    import torch

    import bendy_tracer_b200 as bt
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    w, h, passes = 320, 200, 2 * n + 1
    sc, cam = _setup("scene", w, h, LENS_SCENE)
    multi = bt.Engine(devices=list(range(n)))
    assert lib_count(multi) == n
    for device in ("cuda:0", None):
        got, cnt, _ = _render(multi, sc, cam, w, h, passes, 2, device)
        ref, _, _ = _render(bt.Engine.default(0), *_setup("scene", w, h, LENS_SCENE), w, h, passes, 2, device)
        assert cnt == passes * 4
        assert np.allclose(got, ref, rtol=2e-6, atol=2e-5)


def lib_count(engine):
    from bendy_tracer_b200._ffi import lib
    return lib.bt_engine_device_count(engine.handle)


def test_render_sharded_nccl_ranks():
    """one process per GPU (torch.distributed, NCCL): every rank renders its pass slice, one reduce -- the sum equals the
    one-GPU frame (tools/check_multi_gpu.py under torchrun on every GPU of the box)"""
    import os
    import subprocess
    import sys

    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "check_multi_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "ALL OK" in out.stdout
