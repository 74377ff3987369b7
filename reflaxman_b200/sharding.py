"""Multi-GPU partitioning of the hot path.  One process per GPU; pixels and frames are independent GIVEN
their position in the reference's serial random stream, so the only things to agree on are who renders what and where
in the stream it starts.  No data-path collective is needed for frame sharding; a split frame is gathered by storing
straight into the gathering GPU's framebuffer (peer mapping over NVLink) — see split_frame().

Pure functions first (testable on CPU with gloo), then thin drivers over :class:`reflaxman_b200.capi.Context`.
"""
from __future__ import annotations


def calls_per_frame(width, height, samples):
    """Scene::trace calls (= randDir draws) one frame consumes (reference Render.cpp:152-186)."""
    if samples > 0:
        return width * height * samples * samples
    a = -samples
    return ((width + a - 1) // a) * ((height + a - 1) // a)


def frame_shard(n_frames, world, rank):
    """Config 5: frames dealt round-robin; rank r renders frames r, r + world, ..."""
    return list(range(rank, n_frames, world))


def contiguous_frame_shard(n_frames, world, rank):
    """Weak-scaling batches: rank r owns the r-th contiguous block of frames."""
    per = (n_frames + world - 1) // world
    return list(range(rank * per, min(n_frames, (rank + 1) * per)))


def strip_ranges(width, height, world, rank, strip_rows=16):
    """Config 3: interleaved strips of ``strip_rows`` rows dealt round-robin (contiguous bands are badly balanced:
    sky rows take 1 bounce, floor rows 2-20).  Returns increasing [(p0, p1)] linear pixel ranges."""
    out = []
    n_strips = (height + strip_rows - 1) // strip_rows
    for s in range(rank, n_strips, world):
        y0, y1 = s * strip_rows, min(height, (s + 1) * strip_rows)
        out.append((y0 * width, y1 * width))
    return out


def render_frames_sharded(ctx, cams, frames, refl, samples, out, first_frame_in_stream=0):
    """Render the given (increasing) global frame indices of a camera path on this rank.  The context's randDir stream
    must stand at global frame ``first_frame_in_stream``; frames this rank does not own are skipped in the stream."""
    import numpy as np
    from . import capi
    pos = first_frame_in_stream
    per = calls_per_frame(ctx.W, ctx.H, samples)
    for k, f in enumerate(frames):
        if f > pos:
            ctx.skip_samples((f - pos) * per)
        ctx.render_frames(capi.pack_cameras([cams[f]]), refl, samples, out=out[k:k + 1])
        pos = f + 1
    return pos


def split_frame(ctx, cam, refl, samples, world, rank, gather_ptr, strip_rows=16, stream=0):
    """Render this rank's strips of ONE frame, storing ARGB straight into ``gather_ptr`` (a full-frame buffer that may
    be peer memory on the gathering GPU).  Leaves the random stream where a full-frame render would."""
    ctx.set_camera(cam)
    ctx.render_begin(refl, samples, False)
    ctx.render_strips(strip_rows, world, rank, gather_ptr, stream)   # one K1 pass + one K2 launch over this rank's rows


def split_frame_by_ranges(ctx, cam, refl, samples, world, rank, gather_ptr, strip_rows=16, stream=0):
    """The same through the general range API (one launch per strip) — kept as the cross-check of render_strips."""
    ctx.set_camera(cam)
    ctx.render_begin(refl, samples, False)
    for p0, p1 in strip_ranges(ctx.W, ctx.H, world, rank, strip_rows):
        ctx.render_range(p0, p1, gather_ptr, stream)
    ctx.render_finish()
