"""Secondary workloads of BASELINE.json (configs[2], [3], [4]) — measurement helpers used by `bench.py --workload ...`.
They are not the driver's headline line (that is config 2); results are recorded under profiles/."""
from __future__ import annotations

import json
import os
import time


def _dist():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return torch, dist, rank, world, local


def _maxreduce(torch, dist, world, x):
    if world > 1:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def _sumreduce(torch, dist, world, x):
    if world > 1:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    return x


def config3(args):
    """7680x4320 single frame split by interleaved 16-row strips across the ranks; every rank stores its strips straight
    into rank 0's framebuffer through a cudaIpc peer mapping (NVLink), so the gather is fused into K2's stores."""
    import numpy as np
    from reflaxman_b200 import capi, scenes as S, sharding as P
    torch, dist, rank, world, local = _dist()
    Wd, Hd, depth = 7680, 4320, 20
    ctx = capi.Context(local)
    ctx.load_scene(S.default_scene()); ctx.set_image_size(Wd, Hd); ctx.set_seeds(12345, 12345)
    cam = S.default_camera()
    handle = [None]
    if rank == 0:
        gather = ctx.buffer_alloc(Wd * Hd * 4)
        handle[0] = ctx.ipc_export(gather)
    if world > 1:
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            gather = ctx.ipc_import(handle[0])
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    K, Wm = args.steps, max(args.warmup, 3)

    def step():
        P.split_frame(ctx, cam, depth, 1, world, rank, gather, strip_rows=16, stream=stream.cuda_stream)

    for _ in range(Wm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ctx.stats_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = _maxreduce(torch, dist, world, e0.elapsed_time(e1))
    st = ctx.stats()
    rays = _sumreduce(torch, dist, world, float(st["rays"]))
    if rank == 0:
        img = ctx.buffer_read(gather, np.zeros((Hd, Wd), np.uint32))
        print(json.dumps({"workload": "config3: default scene 7680x4320 depth 20, one frame split by 16-row interleaved strips, peer-store gather to GPU 0",
                          "n_gpus": world, "steps": K, "ms_per_frame": ms / K, "frames_per_s": K / (ms * 1e-3), "Mrays_per_s": rays / (ms * 1e-3) / 1e6,
                          "scaling": "strong", "gathered_nonzero_fraction": float((img != 0).mean()),
                          "nvlink_bytes_per_frame": int(Wd * Hd * 4 * (world - 1) / world)}))
    if world > 1:
        dist.barrier()
        if rank != 0:
            ctx.ipc_close(gather)
        dist.barrier()
        dist.destroy_process_group()


def config5(args):
    """240-frame orbiting camera path at 1920x1080, frames dealt round-robin to the ranks, no communication."""
    import numpy as np
    from reflaxman_b200 import capi, scenes as S, sharding as P
    torch, dist, rank, world, local = _dist()
    Wd, Hd, depth, NF = 1920, 1080, 20, 240
    ctx = capi.Context(local)
    ctx.load_scene(S.default_scene()); ctx.set_image_size(Wd, Hd)
    cams = S.orbit_cameras(NF)
    mine = P.frame_shard(NF, world, rank)
    packed = capi.pack_cameras([cams[f] for f in mine])
    out = torch.empty((len(mine), Hd, Wd), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    per = P.calls_per_frame(Wd, Hd, 1)
    K, Wm = args.steps, max(args.warmup, 3)

    def step():
        # the path restarts from the pinned seed every step; frames owned by other ranks are skipped in the stream
        ctx.set_seeds(12345, 12345)
        pos = 0
        for k, f in enumerate(mine):
            if f > pos:
                ctx.skip_samples((f - pos) * per)
            ctx.render_frames_device(packed[k:k + 1], depth, 1, out[k].data_ptr(), stream.cuda_stream)
            pos = f + 1

    for _ in range(Wm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ctx.stats_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = _maxreduce(torch, dist, world, e0.elapsed_time(e1) * 1e-3)
    st = ctx.stats()
    rays = _sumreduce(torch, dist, world, float(st["rays"]))
    if rank == 0:
        print(json.dumps({"workload": "config5: 240-frame orbit of the default scene, 1920x1080 depth 20, frames round-robin over ranks, no communication",
                          "n_gpus": world, "steps": K, "ms_per_path": 1e3 * dt / K, "frames_per_s": K * NF / dt, "Mrays_per_s": rays / dt / 1e6,
                          "scaling": "strong", "timing": "CUDA events on the launching stream, max over ranks; stream skips of the other ranks' frames included"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config4(args):
    """1024 random reflective spheres + textured floor/back wall at 3840x2160, depth sweep 1..8 (shared-memory kernel,
    bounding-volume hierarchy over the spheres; `--frames 0` measures the reference's brute-force list walk instead)."""
    from reflaxman_b200 import capi, scenes as S
    torch, dist, rank, world, local = _dist()
    Wd, Hd = 3840, 2160
    ctx = capi.Context(local)
    ctx.load_scene(S.synthetic_scene(32, floor=S.synthetic_texture(1024, 1024, 11), skybox=S.synthetic_texture(2048, 1536, 7)))
    ctx.set_image_size(Wd, Hd); ctx.set_seeds(12345, 12345)
    brute = args.frames == 0
    ctx.set_bvh_mode(2 if brute else 0)
    if os.environ.get("RFX_FORCE_PATH"):
        ctx.force_path(int(os.environ["RFX_FORCE_PATH"]))   # 3 = general blob kernel only (A/B against the batch kernel)
    out = torch.empty((1, Hd, Wd), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    cam = capi.pack_cameras([S.default_camera()])
    res = []
    for depth in (1, 2, 4, 8):
        ctx.render_frames_device(cam, depth, 1, out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        ctx.stats_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            ctx.render_frames_device(cam, depth, 1, out.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        st = ctx.stats()
        res.append({"depth": depth, "ms_per_frame": ms, "Mrays_per_s": st["rays"] / args.steps / (ms * 1e-3) / 1e6, "rays_per_frame": st["rays"] // args.steps})
    if rank == 0:
        print(json.dumps({"workload": "config4: 1024 spheres + textured floor/wall, 3840x2160, " + ("brute-force list walk" if brute else "BVH over the spheres (results identical to brute force)"),
                          "n_gpus": 1, "sweep": res}))
