"""Multi-GPU parity worker: run under torchrun with one rank per GPU (tests/test_mgpu_gpu.py launches it; also
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_worker.py`).

Every rank renders its share straight into rank 0's delivery buffer through rxc_mgpu_* (peer writes over NVLink, or
the grouped-NCCL fallback when RXC_MGPU_FORCE_NCCL=1); rank 0 then renders the same frames alone into ordinary device
memory and compares BIT FOR BIT:
  1. frame sharding: a camera sweep, contiguous blocks of frames per rank, three steps through a two-slot ring with
     the release hand-shake;
  2. row bands and 3. column bands of one frame, written in place into a full frame.
Prints one line `MGPU_WORKER_OK mode=<peer|nccl|local> world=N` on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from rusterix_b200 import DeviceContext, Rasterizer, mgpu, scenes

    ctx = DeviceContext.get(local)
    dl = mgpu.Delivery(ctx, rank, world)

    # ---- 1. frame sharding
    W, H, F = 640, 360, 3
    cfg = scenes.sweep(width=W, height=H, tile_size=40, n_frames=64, logo_size=128)
    fb = W * H * 4
    slots = 2
    buf = dl.target(slots * world * F * fb)
    mode = dl.status()["mode"]
    for step in range(3):
        ids = [(step * world * F + rank * F + i) % 64 for i in range(F)]   # contiguous block of this rank
        batch = Rasterizer.prepare_batch([cfg.rasterizer(i).on_device(local) for i in ids], cfg.scene, W, H, cfg.tile_size, cfg.assets, device=local)
        slot = step % slots
        if step >= slots:
            dl.release()   # rank 0 has checked the step that used this slot
        dl.render(batch, (slot * world + rank) * F * fb)
        dl.deliver(mgpu.frame_regions(world, F, fb, slot * world * F * fb))
        if rank == 0:
            ctx.synchronize()
            got = buf[slot * world * F * fb:(slot + 1) * world * F * fb].reshape(world * F, H, W, 4)
            all_ids = [(step * world * F + k) % 64 for k in range(world * F)]
            ref = torch.empty((world * F, H, W, 4), dtype=torch.uint8, device=dev)
            rb = Rasterizer.prepare_batch([cfg.rasterizer(i).on_device(local) for i in all_ids], cfg.scene, W, H, cfg.tile_size, cfg.assets, device=local)
            rb.run(ref, sync=True)
            assert torch.equal(got, ref), f"frame sharding step {step}: delivered frames differ from the single-GPU render"
            assert int(ref.max()) > 0
    ctx.synchronize()

    # ---- 2./3. one frame split into row bands, then column bands, written in place
    W, H = 1024, 800
    cfg = scenes.map_config(width=W, height=H, tile_size=40, logo_size=128)
    rast = cfg.rasterizer().on_device(local)
    buf = dl.target(W * H * 4)
    ref = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
    if rank == 0:
        rast.rasterize(cfg.scene, ref, W, H, cfg.tile_size, cfg.assets)
    for kind in ("rows", "columns"):
        if kind == "rows":
            bands = [mgpu.band_for_rank(H, r, world) + (0, W) for r in range(world)]
        else:
            bands = [(0, H) + mgpu.column_band_for_rank(W, r, world) for r in range(world)]
        if rank == 0:
            buf.zero_()
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        y0, y1, x0, x1 = bands[rank]
        if y1 > y0 and x1 > x0:
            batch = Rasterizer.prepare_batch([rast], cfg.scene, W, H, cfg.tile_size, cfg.assets, band=(y0, y1, x0, x1), device=local)
            dl.render(batch, (y0 * W + x0) * 4, pitch_bytes=W * 4)
        dl.deliver(mgpu.band_regions(bands, W))
        if rank == 0:
            ctx.synchronize()
            assert torch.equal(buf[: W * H * 4].reshape(H, W, 4), ref), f"{kind} bands written in place differ from the single-GPU frame"
        dl.release()
    ctx.synchronize()
    st = dl.status()
    assert st["timeouts"] == 0, st
    dl.close()
    modes = [mode]
    if world > 1:
        modes = [None] * world
        dist.all_gather_object(modes, mode)
    if rank == 0:
        print(f"MGPU_WORKER_OK mode={modes[-1]} world={world}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
