"""Bare host->device copy ceiling of the box, per GPU and aggregate, with N ranks copying AT THE SAME TIME (run under torchrun):
every rank repeatedly copies a pinned 1.26 GB buffer (256 pages 1280x1280 RGB, bench.py's raw-RGB e2e upload) to its GPU with plain
cudaMemcpyAsync — no kernels, no library of this repo.  bench.py's raw-RGB e2e figure is printed as a fraction of this number in
DESIGN.md; the JPEG e2e uploads 25x fewer bytes and is not bound by it.

    python tools/measure_h2d.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/measure_h2d.py
Prints one JSON line on rank 0."""
import json
import os
import time

import torch


def main():
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = 256 * 1280 * 1280 * 3
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    out = {}
    for name, chunk in (("one_copy", n), ("page_sized_copies", 1280 * 1280 * 3)):
        for _ in range(2):
            with torch.cuda.stream(st):
                for off in range(0, n, chunk):
                    d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
            st.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(reps):
                for off in range(0, n, chunk):
                    d[off:off + chunk].copy_(h[off:off + chunk], non_blocking=True)
            e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1)
        gbs = n * reps / (ms * 1e-3) / 1e9
        t = torch.tensor([gbs, ms], device="cuda", dtype=torch.float64)
        if dist is not None:
            mn, mx = t.clone(), t.clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            agg = n * reps * world / (float(mx[1]) * 1e-3) / 1e9
            out[name] = {"per_gpu_gbs_min": float(mn[0]), "per_gpu_gbs_max": float(mx[0]), "aggregate_gbs": agg}
        else:
            out[name] = {"per_gpu_gbs_min": gbs, "per_gpu_gbs_max": gbs, "aggregate_gbs": gbs}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy_round": n, "host_cores": os.cpu_count(), "h2d": out,
                          "pages_1280_rgb_per_s_ceiling": out["page_sized_copies"]["aggregate_gbs"] * 1e9 / (1280 * 1280 * 3)}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
