"""Multi-GPU: disjoint spp slices per rank + ONE framebuffer reduce.

Samples are i.i.d. and the buffer is a running sum (reference src/tracer/buffer.rs:159-164), so the
path shards by global pass index with no data-path exchange until the final sum.  The per-path
RNG is keyed by the GLOBAL pass index, so the union of the ranks' sample sets is exactly the
1-GPU sample set; the images agree up to f32 summation order.
"""
import numpy as np


def shard_passes(samples, world_size, rank):
    """Contiguous slice [lo, hi) of the global pass range [0, samples) owned by `rank`."""
    lo = samples * rank // world_size
    hi = samples * (rank + 1) // world_size
    return lo, hi


def render_sharded(tracer, scene, camera, render_config, buffer, *, group=None, dst=0, all_ranks=False,
                   sample_base=None):
    """Each rank renders its slice of `render_config.samples` passes into a private zeroed buffer of
    `buffer`'s shape; the slices are summed with one reduce (all_reduce if all_ranks) and added to
    `buffer` on `dst` (a GLOBAL rank; on every rank if all_ranks).  Works with nccl (device buffers)
    and gloo (host buffers).  Returns the Status of the local render.

    `sample_base` is the global pass index of this call's first pass and must be the same on every
    rank.  Default: the passes `sample_base_hint` says the frame already holds -- the caller's running
    pass count, which every rank tracks in `buffer.samples()` (the counter is bumped on every rank, the
    image only where the sum lands), so that the reference's progressive loop (src/main.rs:245-254)
    keeps drawing NEW sample sets call after call, exactly as Tracer.render does."""
    import dataclasses

    import torch
    import torch.distributed as dist

    from .api import Buffer, Status

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if sample_base is None:
        sample_base = buffer.samples() // render_config.subsample.subpixel_count()
    lo, hi = shard_passes(render_config.samples, world, rank)
    local = Buffer(buffer.width(), buffer.height(), buffer.color_space, device=buffer.device)
    local.data[..., 3] = 0.0  # alpha is not a sum: keep the caller's (buffer.rs:159-164 never writes it)
    status = Status.Done
    if hi > lo:
        rc = dataclasses.replace(render_config, samples=hi - lo)
        status = tracer.render(scene, camera, rc, local, sample_base=sample_base + lo)
    t = local.data if not isinstance(local.data, np.ndarray) else torch.from_numpy(local.data)
    if all_ranks:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if all_ranks or dist.get_rank() == dst:   # dst is a global rank (torch.distributed.reduce's convention)
        if isinstance(buffer.data, np.ndarray):
            buffer.data += t.numpy() if not isinstance(local.data, np.ndarray) else local.data
        else:
            buffer.data += t
    # the pass counter advances on every rank (it keys the next call's sample_base); the image only where the sum lands
    buffer._samples += render_config.samples * render_config.subsample.subpixel_count()
    if render_config.samples == 0:
        return Status.Done
    return Status.InProgress if status == Status.Done and hi == lo else status
