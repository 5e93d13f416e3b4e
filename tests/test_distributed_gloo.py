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
            assert (data[..., :3] == 2.0).all() and n == 8           # non-root buffers untouched
    assert sorted(covered) == list(range(10, 10 + samples))          # disjoint, complete pass ranges
