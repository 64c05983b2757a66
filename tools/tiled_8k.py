"""BASELINE configs[4]: ONE 1920x1080 -> 7680x4320 MewZoom-4X-Ctrl frame, cut into halo-padded LR tiles, one tile per
rank (one process per GPU, no collective on the data path; NCCL only for the timing barrier / max-reduce).

    python tools/tiled_8k.py                                                   # 1 GPU, 1 tile (= the whole frame)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/tiled_8k.py

Every rank runs the whole network on its haloed tile (halo = 2L+1 = 81 LR pixels, exact) and puts the HR core of the
tile into the frame assembled on rank 0's GPU: a one-sided 2-D copy over NVLink into rank 0's buffer, mapped once
through a CUDA IPC handle (sharding.share_frame): by default the head kernel of the tile stores its core straight into
that buffer (run_tile_into -> mz_upscale_window: the stores are the transfer), `--copy-put` writes the tile locally and
copies the core (put_core), `--no-stitch` leaves the cores on their GPUs; `--io-uint8` runs 8-bit frames.  Strong
scaling: value = 33.2 output Mpx / max-over-ranks device time, puts included.  Rank 0 checks the assembled frame
against the un-tiled result.  Prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ultrazoom_b200 import MODEL_CONFIGS, MewZoom  # noqa: E402
from ultrazoom_b200.sharding import (best_grid, frames_for_rank, halo_radius, plan_tiles, put_core, run_tile,  # noqa: E402
                                     run_tile_into, share_frame)


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    steps, warmup = 5, 3
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = MODEL_CONFIGS["MewZoom-4X-Ctrl"]
    r, L, H, W = cfg["upscale_ratio"], cfg["num_encoder_layers"], 1080, 1920
    torch.manual_seed(0)
    model = MewZoom(**cfg).to(dev).eval()
    g = torch.Generator().manual_seed(1234)           # every rank holds the same LR frame
    x = torch.rand(1, 3, H, W, generator=g).to(dev)
    io8 = "--io-uint8" in sys.argv                     # 8-bit frame in and out: a quarter of the bytes put over NVLink
    if io8:
        x = (x * 255).to(torch.uint8)
    c = torch.tensor([[0.5, 0.2, 0.3]], device=dev)
    rows, cols = best_grid(H, W, world, halo_radius(L), align_w=128)   # columns sized for the kernel's 128-pixel tiles
    plan = plan_tiles(H, W, rows, cols, halo_radius(L), align_w=128)
    mine = [plan[i] for i in frames_for_rank(len(plan), rank, world)]

    stitched = "--no-stitch" not in sys.argv
    frame = None
    if stitched:                                       # the assembled 8K frame lives on rank 0's GPU
        shared = share_frame((1, 3, H * r, W * r), torch.uint8 if io8 else torch.float32, 0, rank, dev)
        frame = shared.tensor

    copy_put = "--copy-put" in sys.argv

    def step():
        if stitched and not copy_put:
            for t in mine:
                run_tile_into(model, x, c, t, r, frame)
            return None
        cores = [run_tile(model.upscale, x, c, t, r) for t in mine]
        if stitched:
            for t, core in zip(mine, cores):
                put_core(frame, core, t, r)
        return cores

    for _ in range(warmup):
        cores = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        cores = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    err = None
    if rank == 0:                                      # exactness of the tiling against the un-tiled frame
        full = model.upscale(x, c)
        t0 = mine[0]
        if stitched:                                   # every rank's puts landed before the barrier above returned
            err = float((full.float() - frame.float()).abs().max())
        else:
            err = float((full[:, :, t0.y0 * r:t0.y1 * r, t0.x0 * r:t0.x1 * r].float() - cores[0].float()).abs().max())
        executed = sum((t.hy1 - t.hy0) * (t.hx1 - t.hx0) for t in plan) / (H * W)
        print(json.dumps({
            "metric": "output_mpx_per_s", "value": H * r * W * r / (ms * 1e-3) / 1e6, "unit": "Mpx/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "config": {"workload": "MewZoom-4X-Ctrl 96ch/40L, one 1920x1080->7680x4320 frame, halo-tiled "
                                   f"{rows}x{cols} (BASELINE configs[4])", "halo_lr_px": halo_radius(L), "image_io": "uint8" if io8 else "float32",
                       "executed_over_algorithmic_work": executed,
                       "stitch": ("none" if not stitched else "2-D copies of the cores into rank 0's frame (CUDA IPC peer mapping)"
                                  if copy_put else "head kernel stores the core into rank 0's frame (CUDA IPC peer mapping)")},
            "max_abs_diff_vs_untiled": err}), flush=True)
    if world > 1:
        dist.barrier()
    if stitched:
        frame = None
        shared.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
