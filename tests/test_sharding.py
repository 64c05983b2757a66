"""CPU: host-side partitioning -- halo tiling is exact, frame-stream assignment is a partition, and the
world_size-2 path works over gloo.  The compute function here is the oracle (tests may use it)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import make_oracle, max_abs_err
from ultrazoom_b200.sharding import best_grid, frames_for_rank, halo_radius, plan_tiles, upscale_tiled

CFG = dict(upscale_ratio=2, num_channels=16, hidden_ratio=2, num_encoder_layers=3, control_features=3)


def test_frames_for_rank_is_a_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in frames_for_rank(64, r, world))
        assert seen == list(range(64))
    assert frames_for_rank(5, 3, 4) == [3]
    assert frames_for_rank(2, 3, 4) == []          # ragged: more ranks than frames
    with pytest.raises(AssertionError):
        frames_for_rank(4, 4, 4)


def test_plan_tiles_covers_image_and_clips_halo():
    tiles = plan_tiles(1080, 1920, 2, 4, halo_radius(40))
    assert len(tiles) == 8 and halo_radius(40) == 81
    area = sum((t.y1 - t.y0) * (t.x1 - t.x0) for t in tiles)
    assert area == 1080 * 1920
    t0 = tiles[0]
    assert (t0.hy0, t0.hx0) == (0, 0) and t0.hy1 == 540 + 81 and t0.hx1 == 480 + 81
    executed = sum((t.hy1 - t.hy0) * (t.hx1 - t.hx0) for t in tiles) / area
    assert 1.40 < executed < 1.48                   # 4 cols x 2 rows -> 1.44x
    assert best_grid(1080, 1920, 8, 81) == (2, 4)
    with pytest.raises(AssertionError):
        plan_tiles(4, 4, 8, 1, 1)


def test_columns_sized_for_the_128_pixel_kernel_tiles():
    """align_w = 128: every haloed tile of the 1080p frame on a 2 x 4 grid is at most 640 pixels (5 kernel tiles) wide,
    where the equal split makes the two middle columns 642 (6 tiles); the cores still partition the image."""
    halo = halo_radius(40)
    eq = plan_tiles(1080, 1920, 2, 4, halo)
    al = plan_tiles(1080, 1920, 2, 4, halo, align_w=128)
    assert max(-(-(t.hx1 - t.hx0) // 128) for t in eq) == 6
    assert max(-(-(t.hx1 - t.hx0) // 128) for t in al) == 5
    assert sum((t.y1 - t.y0) * (t.x1 - t.x0) for t in al) == 1080 * 1920
    xs = sorted({(t.x0, t.x1) for t in al})
    assert xs[0][0] == 0 and xs[-1][1] == 1920 and all(a[1] == b[0] for a, b in zip(xs, xs[1:]))
    for n in (1, 2, 4, 8):
        r_, c_ = best_grid(1080, 1920, n, halo, align_w=128)
        assert r_ * c_ == n
    # small images: one kernel tile per column is enough, the cores still partition the width
    small = plan_tiles(8, 40, 1, 4, 3, align_w=128)
    assert small[0].x0 == 0 and small[-1].x1 == 40 and all(a.x1 == b.x0 and a.x1 > a.x0 for a, b in zip(small, small[1:]))


@pytest.mark.parametrize("rows,cols,align", [(1, 1, 1), (2, 2, 1), (3, 2, 1), (2, 3, 16)])
def test_tiled_inference_is_exact(rows, cols, align):
    m = make_oracle(CFG, seed=5)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 3, 37, 29, generator=g)
    c = torch.rand(2, 3, generator=g)
    full = m.upscale(x, c)
    tiled = upscale_tiled(m.upscale, x, c, 2, CFG["num_encoder_layers"], rows, cols, align_w=align)
    assert max_abs_err(full, tiled) <= 2e-6         # fp reassociation only


def test_halo_one_short_is_not_exact():
    import ultrazoom_b200.sharding as sh

    m = make_oracle(CFG, seed=5)
    x = torch.rand(1, 3, 40, 40, generator=torch.Generator().manual_seed(2))
    c = torch.tensor([0.5, 0.2, 0.3])
    full = m.upscale(x, c)
    plan = plan_tiles(40, 40, 2, 2, halo_radius(3) - 1)
    out = torch.zeros_like(full)
    for t in plan:
        sh.stitch(out, sh.run_tile(m.upscale, x, c, t, 2), t, 2)
    assert max_abs_err(full, out) > 1e-5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = make_oracle(CFG, seed=5)
        g = torch.Generator().manual_seed(9)
        x = torch.rand(1, 3, 30, 34, generator=g)
        c = torch.tensor([0.5, 0.2, 0.3])
        mine = frames_for_rank(4, rank, world)
        # each rank fills only its own tiles; the (test-only) sum over ranks stitches them
        part = upscale_tiled(m.upscale, x, c, 2, CFG["num_encoder_layers"], 2, 2, tiles=mine)
        dist.all_reduce(part)
        full = m.upscale(x, c)
        # frame-stream partition: every frame processed exactly once
        count = torch.zeros(7)
        for i in frames_for_rank(7, rank, world):
            count[i] += 1
        dist.all_reduce(count)
        ret[rank] = (max_abs_err(full, part), count.tolist())
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        for r in range(world):
            err, count = ret[r]
            assert err <= 2e-6
            assert count == [1.0] * 7


@pytest.mark.parametrize("H,W,rows,cols,halo,r,align", [(45, 290, 2, 2, 7, 2, 128), (1080, 1920, 2, 4, 81, 4, 128),
                                                       (33, 70, 3, 2, 5, 3, 1)])
def test_run_tile_into_windows_tile_the_frame_exactly_once(H, W, rows, cols, halo, r, align):
    """Host logic of the fused stitch (MewZoom.upscale_into): for every tile the window handed to the kernel is the
    tile's core in tile-local coordinates and lands at the core's HR position; together the windows cover the frame
    exactly once.  A stand-in model records what it is asked to do and "upscales" by nearest-neighbour replication."""
    from ultrazoom_b200.sharding import plan_tiles, run_tile_into

    class Recorder:
        def __init__(self):
            self.calls = []

        def upscale_into(self, x, c, frame, window, at):
            y0, y1, x0, x1 = window
            fy, fx = at
            assert 0 <= y0 < y1 <= x.shape[2] and 0 <= x0 < x1 <= x.shape[3]
            core = x[:, :, y0:y1, x0:x1].repeat_interleave(r, 2).repeat_interleave(r, 3)
            frame[:, :, fy:fy + core.shape[2], fx:fx + core.shape[3]] += core
            self.calls.append((window, at))

    x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(0)) + 1.0
    frame = torch.zeros(1, 3, H * r, W * r)
    rec = Recorder()
    plan = plan_tiles(H, W, rows, cols, halo, align)
    for t in plan:
        run_tile_into(rec, x, None, t, r, frame)
    assert len(rec.calls) == len(plan)
    assert torch.equal(frame, x.repeat_interleave(r, 2).repeat_interleave(r, 3))   # every HR pixel written exactly once


def test_refresh_rectangles_cover_every_halo_exactly_once():
    """Periodic halo refresh (sharding.upscale_tiled_refresh): what tile i receives from its neighbours -- the parts of
    its haloed region inside their cores -- covers its halo exactly once, never its core; tile rows start on multiples
    of ROW_ALIGN (the fused block's accumulator ring is tied to the image row: bit-exact tiling)."""
    import numpy as np

    from ultrazoom_b200.sharding import ROW_ALIGN, plan_tiles, refresh_rects

    for (H, W, rows, cols, hw, aw) in [(1080, 1920, 2, 4, 21, 128), (70, 300, 2, 2, 5, 1), (45, 290, 3, 1, 7, 128), (64, 64, 1, 1, 9, 1)]:
        plan = plan_tiles(H, W, rows, cols, hw, aw)
        rects = refresh_rects(plan)
        for t in plan:
            assert t.hy0 % ROW_ALIGN == 0 and t.hy0 <= max(0, t.y0 - hw)
            cover = np.zeros((H, W), int)
            for (i, j), r in rects.items():
                if i == t.index:
                    o = plan[j]
                    assert o.y0 <= r[0] < r[1] <= o.y1 and o.x0 <= r[2] < r[3] <= o.x1      # inside the sender's core
                    cover[r[0]:r[1], r[2]:r[3]] += 1
            want = np.zeros((H, W), int)
            want[t.hy0:t.hy1, t.hx0:t.hx1] = 1
            want[t.y0:t.y1, t.x0:t.x1] = 0
            assert (cover == want).all()


def test_bind_process_to_gpu_is_a_harmless_hint_without_a_gpu():
    """`sharding.bind_process_to_gpu` pins a rank to the CPUs next to its GPU (node-local pinned buffers for the
    end-to-end path); without NVML / a device it changes nothing and returns None instead of raising."""
    import os

    import torch

    from ultrazoom_b200.sharding import bind_process_to_gpu

    before = os.sched_getaffinity(0)
    got = bind_process_to_gpu(0)
    if not torch.cuda.is_available():
        assert got is None and os.sched_getaffinity(0) == before
    else:
        assert got is None or set(got) == os.sched_getaffinity(0)
        os.sched_setaffinity(0, before)

