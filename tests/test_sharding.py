"""Multi-GPU partitioning.  CPU: pure partition properties + a world_size-2 gloo run in which each rank renders its
frames of a camera path with the ORACLE as the engine (skipping the random stream over the other rank's frames) and
the gathered result equals the single-process sequence.  GPU: the same through the C ABI, plus a split frame
assembled from two contexts' strips equals the unsplit frame bit for bit."""
import os
import socket

import numpy as np
import pytest

from reflaxman_b200 import scenes as S, sharding as P

W, H, REFL = 48, 36, 8


def test_partitions_cover_exactly_once():
    for world in (1, 2, 3, 4, 8):
        fr = sorted(f for r in range(world) for f in P.frame_shard(240, world, r))
        assert fr == list(range(240))
        fr = sorted(f for r in range(world) for f in P.contiguous_frame_shard(30, world, r))
        assert fr == list(range(30))
        for (w, h, rows) in ((7680, 4320, 16), (1920, 1080, 16), (100, 37, 5)):
            px = np.zeros(w * h, np.uint8)
            for r in range(world):
                last = -1
                for p0, p1 in P.strip_ranges(w, h, world, r, rows):
                    assert p0 > last and p0 % w == 0 and p1 % w == 0
                    px[p0:p1] += 1
                    last = p0
            assert px.min() == 1 and px.max() == 1
    assert P.calls_per_frame(10, 7, 3) == 10 * 7 * 9 and P.calls_per_frame(10, 7, -4) == 3 * 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as O
    cams = S.orbit_cameras(6)
    mine = P.frame_shard(len(cams), world, rank)
    r = O.OracleRender(S.default_scene(), W, H, seed=77, nthreads=1)
    per = P.calls_per_frame(W, H, 1)
    pos = 0
    out = torch.zeros((len(cams), H, W), dtype=torch.int64)
    for f in mine:
        if f > pos:                                   # skip the stream over frames owned by other ranks
            _, st = O.rand_dirs(int(r.seeds[0]), (f - pos) * per)
            r.seeds[0] = st
        out[f] = torch.from_numpy(r.render(cams[f], REFL).resolve()[1].astype(np.int64))
        pos = f + 1
    dist.all_reduce(out)                              # test-side gather only: frames are disjoint, the rest is zero
    if rank == 0:
        q.put(out.numpy())
    dist.destroy_process_group()


def test_frame_sharding_world2_gloo(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cams = S.orbit_cameras(6)
    r = oracle.OracleRender(S.default_scene(), W, H, seed=77, nthreads=2)
    for f, cam in enumerate(cams):
        want = r.render(cam, REFL).resolve()[1]
        assert np.array_equal(got[f].astype(np.uint32), want), f


@pytest.mark.gpu
def test_gpu_frame_sharding_matches_sequence(rfx_lib, oracle):
    from reflaxman_b200 import capi
    cams = S.orbit_cameras(6)
    c = capi.Context(0)
    c.load_scene(S.default_scene()); c.set_seeds(77, 77); c.set_image_size(W, H)
    seq = c.render_frames(cams, REFL)
    c.close()
    world = 3
    for rank in range(world):
        d = capi.Context(0)
        d.load_scene(S.default_scene()); d.set_seeds(77, 77); d.set_image_size(W, H)
        mine = P.frame_shard(len(cams), world, rank)
        out = np.zeros((len(mine), H, W), np.uint32)
        P.render_frames_sharded(d, cams, mine, REFL, 1, out)
        for k, f in enumerate(mine):
            assert np.array_equal(out[k], seq[f]), (rank, f)
        d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("by_ranges", [False, True], ids=["strips", "ranges"])
@pytest.mark.parametrize("samples", [1, 2, -3])
def test_gpu_split_frame_equals_whole_frame(rfx_lib, samples, by_ranges):
    """two 'ranks' (contexts) render interleaved strips of one frame straight into one shared buffer"""
    from reflaxman_b200 import capi
    w, h = 64, 52
    cam = S.default_camera()
    whole = capi.Context(0)
    whole.load_scene(S.default_scene()); whole.set_seeds(5, 5); whole.set_image_size(w, h)
    whole.render(cam, REFL, samples)
    want = whole.read_argb()
    seeds_after = whole.get_seeds()
    gather = whole.buffer_alloc(w * h * 4)
    world = 3
    for rank in range(world):
        d = capi.Context(0)
        d.load_scene(S.default_scene()); d.set_seeds(5, 5); d.set_image_size(w, h)
        (P.split_frame_by_ranges if by_ranges else P.split_frame)(d, cam, REFL, samples, world, rank, gather, strip_rows=6 if by_ranges else 8)
        d.synchronize()
        assert d.get_seeds() == seeds_after          # the stream ends where the unsplit render ends
        d.close()
    got = whole.buffer_read(gather, np.zeros((h, w), np.uint32))
    assert np.array_equal(got, want)
    whole.buffer_free(gather)
    whole.close()
