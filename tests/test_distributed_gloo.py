"""world_size-2/3 gloo tests of the multi-GPU host path (render_sharded): pass-range sharding, the
single framebuffer reduce and the accumulate-into-caller's-buffer semantics.  The kernel launch is
replaced by a deterministic stand-in (there is no GPU here); the sharding / reduce code is the
product's own."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bendy_tracer_b200 as bt


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakeTracer:
    """adds, for every global pass index g in [sample_base, sample_base + samples), the image
    (g + 1) * pattern to the buffer -- so the sum over any partition of the pass range is known."""

    def __init__(self):
        self.calls = []

    def render(self, scene, camera, rc, buffer, *, sample_base=None, stream=None, sync=True):
        self.calls.append((sample_base, rc.samples))
        if rc.samples == 0:
            return bt.Status.Done
        h, w = buffer.height(), buffer.width()
        pattern = np.arange(h * w * 3, dtype=np.float32).reshape(h, w, 3) / 7.0
        k = rc.subsample.subpixel_count()
        for g in range(sample_base, sample_base + rc.samples):
            buffer.data[..., :3] += (g + 1) * k * pattern
        buffer._samples += rc.samples * k
        return bt.Status.InProgress


def _worker(rank, world, port, samples, all_ranks, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tracer = FakeTracer()
        buf = bt.Buffer(5, 3)
        buf.data[..., :3] = 2.0                     # the caller's buffer already holds earlier passes
        buf._samples = 8
        rc = bt.RenderConfig.with_samples_subsample(samples, bt.Subsample.subpixel(2))
        st = bt.render_sharded(tracer, None, 0, rc, buf, all_ranks=all_ranks, sample_base=10)
        out[rank] = (buf.data.copy(), buf.samples(), int(st), tracer.calls)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,samples,all_ranks", [(2, 8, False), (2, 7, True), (3, 2, False), (2, 1, True)])
def test_render_sharded_gloo(world, samples, all_ranks):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), samples, all_ranks, out), nprocs=world, join=True)
    pattern = np.arange(5 * 3 * 3, dtype=np.float32).reshape(3, 5, 3) / 7.0
    expect = 2.0 + sum((g + 1) * 4 for g in range(10, 10 + samples)) * pattern
    covered = []
    for rank in range(world):
        data, n, status, calls = out[rank]
        covered += [g for base, cnt in calls for g in range(base, base + cnt)]
        assert status == int(bt.Status.InProgress)
        if all_ranks or rank == 0:
            assert np.allclose(data[..., :3], expect, rtol=1e-6) and n == 8 + samples * 4
            assert (data[..., 3] == 1.0).all()                       # alpha is not summed
        else:
            assert (data[..., :3] == 2.0).all()                      # non-root images untouched ...
            assert n == 8 + samples * 4                              # ... but every rank's pass counter advances (it keys the next call)
    assert sorted(covered) == list(range(10, 10 + samples))          # disjoint, complete pass ranges


def _progressive_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rc = bt.RenderConfig.with_samples_subsample(3, bt.Subsample.subpixel(2))
        tracer, buf = FakeTracer(), bt.Buffer(5, 3)
        for _ in range(2):                                   # the reference's progressive loop (main.rs:245-254): default sample_base
            bt.render_sharded(tracer, None, 0, rc, buf, all_ranks=True)
        once_tracer, once = FakeTracer(), bt.Buffer(5, 3)
        bt.render_sharded(once_tracer, None, 0, bt.RenderConfig.with_samples_subsample(6, bt.Subsample.subpixel(2)), once, all_ranks=True)
        out[rank] = (buf.data.copy(), buf.samples(), once.data.copy(), once.samples(), tracer.calls)
    finally:
        dist.destroy_process_group()


def test_render_sharded_progressive_calls_draw_new_passes():
    """two calls of 3 passes into one buffer == one call of 6 passes: the default sample_base continues at the
    buffer's pass count on every rank, so repeated calls never re-render the same sample sets"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_progressive_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    covered = []
    for rank in range(2):
        twice, n2, once, n1, calls = out[rank]
        assert n2 == n1 == 24
        assert np.allclose(twice, once, rtol=1e-6)
        covered += [g for base, cnt in calls for g in range(base, base + cnt)]
    assert sorted(covered) == list(range(6))                 # global passes 0..5, each rendered exactly once
