"""Secondary workloads of BASELINE.json — configs[2] (8K frame split by strips, NVLink gather), configs[3] (1024 spheres, depth
sweep) and configs[4] (240-frame orbit sharded over the GPUs).  bench.py runs them after the headline measurement and prints
them in the `secondary` object of its one JSON line; `bench.py --workload configN` runs one alone.

Every function takes an `Env` (torch, torch.distributed, rank, world, local) whose process group already exists, times on the
device with CUDA events on the launching stream, reduces with MAX over ranks, and returns a dict on every rank."""
from __future__ import annotations

import json
import os

CENSUS_C4 = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "census_c4_r2.json")


class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world > 1:
            t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
            self.dist.all_reduce(t, op=op)
            return float(t.item())
        return float(x)

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def close(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.barrier()
            self.dist.destroy_process_group()


def _timed(env, stream, step, steps, warmup):
    """warmup untimed steps, then `steps` steps between two events on `stream`, barrier + synchronize on both sides; ms, max over ranks"""
    torch = env.torch
    for _ in range(warmup):
        step()
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    env.barrier()
    return env.max(e0.elapsed_time(e1))


def config3(env, steps=10, warmup=3, gather_mode="peer", strip_rows=16):
    """7680x4320 single frame split by interleaved strips across the ranks.  gather_mode "peer": every rank's K2 stores its
    strips straight into rank 0's framebuffer through a cudaIpc peer mapping (NVLink), the gather is fused into the kernel's
    stores.  "local": every rank stores into a buffer of its own (no gather; the A/B that isolates the cost of the peer stores).
    Also times the unsplit frame on rank 0 alone, so the strong-scaling efficiency comes from one run."""
    import numpy as np
    from reflaxman_b200 import capi, scenes as S, sharding as P
    torch, dist = env.torch, env.dist
    Wd, Hd, depth = 7680, 4320, 20
    ctx = capi.Context(env.local)
    ctx.load_scene(S.default_scene()); ctx.set_image_size(Wd, Hd); ctx.set_seeds(12345, 12345)
    cam = S.default_camera()
    handle = [None]
    own = env.rank == 0 or gather_mode == "local"
    if own:
        gather = ctx.buffer_alloc(Wd * Hd * 4)
        if env.rank == 0:
            handle[0] = ctx.ipc_export(gather)
    if env.world > 1 and gather_mode == "peer":
        dist.broadcast_object_list(handle, src=0)
        if env.rank != 0:
            gather = ctx.ipc_import(handle[0])
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    # the unsplit frame on one GPU (rank 0; the others wait): the denominator of the strong-scaling efficiency
    single_ms = 0.0
    if env.world > 1:
        if env.rank == 0:
            single_ms = _timed_local(torch, stream, lambda: P.split_frame(ctx, cam, depth, 1, 1, 0, gather, strip_rows=strip_rows, stream=stream.cuda_stream), steps, warmup) / steps
        env.barrier()
        single_ms = env.max(single_ms)
        ctx.set_seeds(12345, 12345)

    def step():
        P.split_frame(ctx, cam, depth, 1, env.world, env.rank, gather, strip_rows=strip_rows, stream=stream.cuda_stream)

    ctx.stats_reset()
    ms = _timed(env, stream, step, steps, warmup)
    st = ctx.stats()
    rays = env.sum(float(st["rays"])) * steps / (steps + warmup)
    res = {"workload": "config3: default scene 7680x4320 depth 20, one frame split by %d-row interleaved strips, %s" %
                       (strip_rows, "peer-store gather into GPU 0's framebuffer over NVLink" if gather_mode == "peer" else "each GPU keeps its strips (no gather: A/B arm)"),
           "n_gpus": env.world, "steps": steps, "warmup": warmup, "ms_per_frame": ms / steps, "frames_per_s": steps / (ms * 1e-3), "Mrays_per_s": rays / (ms * 1e-3) / 1e6,
           "scaling": "strong", "gather": gather_mode,
           "nvlink_bytes_per_frame": int(Wd * Hd * 4 * (env.world - 1) / env.world) if gather_mode == "peer" else 0}
    if env.world > 1:
        res["single_gpu_ms_per_frame"] = single_ms
        res["speedup_vs_1gpu"] = single_ms / (ms / steps)
        res["strong_scaling_efficiency"] = single_ms / (ms / steps) / env.world
    if env.rank == 0 and gather_mode == "peer":
        img = ctx.buffer_read(gather, np.zeros((Hd, Wd), np.uint32))
        res["gathered_nonzero_fraction"] = float((img != 0).mean())
    env.barrier()
    if not own:
        ctx.ipc_close(gather)
    env.barrier()
    if own:
        ctx.buffer_free(gather)
    torch.cuda.set_stream(torch.cuda.default_stream())
    ctx.close()
    return res


def _timed_local(torch, stream, step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def config5(env, steps=3, warmup=3):
    """240-frame orbiting camera path at 1920x1080, frames dealt round-robin to the ranks, no communication."""
    from reflaxman_b200 import capi, scenes as S, sharding as P
    torch = env.torch
    Wd, Hd, depth, NF = 1920, 1080, 20, 240
    ctx = capi.Context(env.local)
    ctx.load_scene(S.default_scene()); ctx.set_image_size(Wd, Hd)
    if os.environ.get("RFX_TILE_ORDER_PERIOD"):
        ctx.set_option("tile_order_period", int(os.environ["RFX_TILE_ORDER_PERIOD"]))   # A/B knob (default 8)
    cams = S.orbit_cameras(NF)
    mine = P.frame_shard(NF, env.world, env.rank)
    packed = capi.pack_cameras([cams[f] for f in mine])
    out = torch.empty((len(mine), Hd, Wd), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    per = P.calls_per_frame(Wd, Hd, 1)

    def step():
        # the path restarts from the pinned seed every step; frames owned by other ranks are skipped in the stream
        ctx.set_seeds(12345, 12345)
        pos = 0
        for k, f in enumerate(mine):
            if f > pos:
                ctx.skip_samples((f - pos) * per)
            ctx.render_frames_device(packed[k:k + 1], depth, 1, out[k].data_ptr(), stream.cuda_stream)
            pos = f + 1

    ctx.stats_reset()
    ms = _timed(env, stream, step, steps, warmup)
    st = ctx.stats()
    rays = env.sum(float(st["rays"])) * steps / (steps + warmup)
    res = {"workload": "config5: 240-frame orbit of the default scene, 1920x1080 depth 20, frames round-robin over the GPUs, no communication",
           "n_gpus": env.world, "steps": steps, "warmup": warmup, "ms_per_path": ms / steps, "frames_per_s": steps * NF / (ms * 1e-3), "Mrays_per_s": rays / (ms * 1e-3) / 1e6,
           "scaling": "strong", "timing": "CUDA events on the launching stream, max over ranks; stream skips of the other ranks' frames included"}
    torch.cuda.set_stream(torch.cuda.default_stream())
    del out
    ctx.close()
    return res


def config4(env, steps=5, brute=False, depths=(1, 2, 3, 4, 5, 6, 7, 8), fp32_peak_tflops=None):
    """1024 random reflective spheres + textured floor / back wall / sky at 3840x2160, reflection depth sweep (one GPU: every rank
    would render the same frame).  BVH over the spheres (results identical to the reference's brute-force list walk, which
    `brute` times instead).  `roofline`: census flops of the REFERENCE's algorithm (list walk: profiles/census_c4_r2.json, counted
    at 480x270) over the kernel time — the hierarchy skips ~99 % of those sphere tests, so the fraction measures algorithmic
    saving, not pipe utilisation; the ncu digest under profiles/ has the pipe numbers."""
    from reflaxman_b200 import capi, scenes as S
    torch = env.torch
    Wd, Hd = 3840, 2160
    ctx = capi.Context(env.local)
    ctx.load_scene(S.synthetic_scene(32, floor=S.synthetic_texture(1024, 1024, 11), skybox=S.synthetic_texture(2048, 1536, 7)))
    ctx.set_image_size(Wd, Hd); ctx.set_seeds(12345, 12345)
    ctx.set_bvh_mode(2 if brute else 0)
    if os.environ.get("RFX_BLOB_WAVEFRONT"):
        ctx.set_option("blob_wavefront", int(os.environ["RFX_BLOB_WAVEFRONT"]))   # 0 = the single tile kernel (A/B against the wavefront), k = segments in tiles
    if os.environ.get("RFX_BLOB_SMEM_BVH"):
        ctx.set_option("blob_smem_bvh", int(os.environ["RFX_BLOB_SMEM_BVH"]))
    if os.environ.get("RFX_FORCE_PATH"):
        ctx.force_path(int(os.environ["RFX_FORCE_PATH"]))   # 3 = general blob kernel only (A/B against the batch kernel)
    if os.environ.get("RFX_EYE_GRID"):
        ctx.set_option("eye_grid", int(os.environ["RFX_EYE_GRID"]))         # 0 = first queries walk the hierarchy (A/B)
    if os.environ.get("RFX_LIGHT_GRIDS"):
        ctx.set_option("light_grids", int(os.environ["RFX_LIGHT_GRIDS"]))   # 0 = shadow queries walk the hierarchy (A/B)
    out = torch.empty((1, Hd, Wd), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    cam = capi.pack_cameras([S.default_camera()])
    census = {}
    try:
        census = json.load(open(CENSUS_C4))["per_depth"]
    except Exception:
        pass
    sweep = []
    for depth in depths:
        ctx.render_frames_device(cam, depth, 1, out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        ctx.stats_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            ctx.render_frames_device(cam, depth, 1, out.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        st = ctx.stats()
        row = {"depth": depth, "ms_per_frame": ms, "Mrays_per_s": st["rays"] / steps / (ms * 1e-3) / 1e6, "rays_per_frame": st["rays"] // steps}
        c = census.get(str(depth))
        if c and fp32_peak_tflops:
            ach = c["flop_per_pixel"] * Wd * Hd / (ms * 1e-3) / 1e12
            row["roofline"] = {"bound": "fp32", "achieved": ach, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": ach / fp32_peak_tflops,
                               "census_flop_per_pixel": c["flop_per_pixel"], "basis": "reference list-walk census (all 1028 objects per ray)"}
        sweep.append(row)
    res = {"workload": "config4: 1024 spheres + textured floor/wall/sky, 3840x2160, " + ("brute-force list walk" if brute else "BVH over the spheres (results identical to brute force)"),
           "n_gpus": 1, "steps": steps, "kernel": "k_trace_blob" if not os.environ.get("RFX_FORCE_PATH") else "forced path " + os.environ["RFX_FORCE_PATH"], "sweep": sweep}
    torch.cuda.set_stream(torch.cuda.default_stream())
    del out
    ctx.close()
    return res


def main(args):
    """bench.py --workload config3|config4|config5: one secondary workload alone, one JSON line on rank 0"""
    env = Env()
    if args.workload == "config3":
        res = config3(env, steps=args.steps, warmup=max(args.warmup, 3), gather_mode=os.environ.get("RFX_C3_GATHER", "peer"),
                      strip_rows=int(os.environ.get("RFX_C3_STRIP_ROWS", "16")))
    elif args.workload == "config5":
        res = config5(env, steps=args.steps, warmup=max(args.warmup, 3))
    else:
        res = config4(env, steps=args.steps, brute=args.frames == 0, depths=(1, 2, 4, 8))
    if env.rank == 0:
        print(json.dumps(res))
    env.close()
