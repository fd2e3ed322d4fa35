"""Side benchmark for BASELINE.json configs[4] on ONE GPU: mixed-size pages (long side log-uniform in [640, 4096] px,
aspect in [0.5, 1]) through retto_b200_run_pages — exercises resize_both (> 2000 px), resize_either up-scaling
(< 736 px) and ragged det tensors.  A pool of unique rendered pages is cycled; forwards are replayed (bench.ReplayWorker).
Prints one JSON line (pages/s device-resident and host-resident + per-kernel ms).  bench.py measures the headline config."""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_mixed(n_unique, seed0=500):
    from oracle import oracle as O
    from tools.synth import gen_page, probmap_from_rects
    rng = np.random.default_rng(seed0)
    pages, probs = [], []
    for i in range(n_unique):
        long_side = int(round(np.exp(rng.uniform(np.log(640), np.log(4096)))))
        short = max(64, int(round(long_side * rng.uniform(0.5, 1.0))))
        h, w = (long_side, short) if rng.random() < 0.5 else (short, long_side)
        img, rects = gen_page(seed0 + i, h, w, n_lines=(max(4, h // 90), max(6, h // 45)))
        ah, aw = O.resize_both_plan(h, w)[-1] if O.resize_both_plan(h, w) else (h, w)
        dh, dw = O.resize_either_plan(ah, aw)
        sx, sy = dw / w, dh / h
        probs.append(probmap_from_rects(seed0 + i, [(r[0] * sx, r[1] * sy, r[2] * sx, r[3] * sy, r[4]) for r in rects], dh, dw))
        pages.append(img)
    return pages, probs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=128)
    ap.add_argument("--unique", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from bench import ReplayWorker
    from retto_b200._lib import Page, Results
    from retto_b200.api import Context
    from retto_b200.shard import shard_indices
    from tools.synth import synth_dict_text
    # under torchrun: `--pages` per GPU; the global list of pages * world pages is LPT-sharded by H*W over the ranks
    # (retto_b200.shard, the page-parallel driver's rule), no collective on the data path; times are the max over ranks
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        saved = os.dup(1); os.dup2(2, 1)   # NCCL banner off stdout
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            dist.barrier()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    ctx = Context(local_rank)
    ctx.dict_load(synth_dict_text())
    L, H = ctx._L, ctx._h
    pages_np, probs_np = make_mixed(args.unique)
    U = len(pages_np)
    mine = shard_indices([pages_np[i % U].shape[0] * pages_np[i % U].shape[1] for i in range(args.pages * world)], world)[rank]
    P = len(mine)
    pages_dev = [torch.from_numpy(pages_np[i % U]).cuda() for i in mine]
    probs_dev = [torch.from_numpy(probs_np[i % U]).cuda() for i in mine]
    sizes = [pages_np[i % U].shape[:2] for i in mine]
    pages_np = [pages_np[i % U] for i in mine]
    U = P
    total_bytes = sum(h * w * 3 for h, w in sizes)
    hp = C.c_void_p()
    ctx._check(L.retto_b200_host_alloc(H, total_bytes + 64 * P, C.byref(hp)))
    offs, o = [], 0
    for i in range(P):
        h, w = sizes[i]
        C.memmove(hp.value + o, pages_np[i % U].ctypes.data, h * w * 3)
        offs.append(o)
        o += (h * w * 3 + 63) & ~63
    pg_dev = (Page * P)(*[Page(pages_dev[i].data_ptr(), sizes[i][0], sizes[i][1], 1) for i in range(P)])
    pg_host = (Page * P)(*[Page(hp.value + offs[i], sizes[i][0], sizes[i][1], 0) for i in range(P)])
    worker = ReplayWorker(torch, f"cuda:{local_rank}", probs_dev, seed=rank)
    res = Results()

    def step(pg):
        worker.begin_step()
        st = L.retto_b200_run_pages(H, pg, P, worker.cb, None, C.byref(res))
        if worker.err is not None:
            raise worker.err
        ctx._check(st)

    stream = ctx.torch_stream()

    def timed(pg, steps, kernels=False):
        for _ in range(3):
            step(pg)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        if kernels:
            ctx.enable_kernel_timing(True); ctx.reset_kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step(pg)
        e1.record(stream)
        torch.cuda.synchronize()
        kt = None
        if kernels:
            kt = {k: v[1] / steps for k, v in ctx.kernel_times().items() if v[0]}
            ctx.enable_kernel_timing(False)
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, kt

    ms_dev, _ = timed(pg_dev, args.steps)
    ms_k, kt = timed(pg_dev, args.steps, kernels=True)
    ms_host, _ = timed(pg_host, args.steps)
    PT = args.pages * world
    if rank == 0:
        print(json.dumps({"config": f"{PT} mixed-size pages per step (long side logU[640,4096], aspect U[0.5,1]; {args.unique} unique), LPT-sharded by H*W over {world} GPU(s)",
                          "n_gpus": world, "pages_rank0": P, "pixels_per_step_rank0": int(sum(h * w for h, w in sizes)), "lines_per_step_rank0": int(res.n_lines),
                          "ms_device_resident": ms_dev, "pages_per_s_device_resident": PT / ms_dev * 1e3,
                          "ms_host_resident": ms_host, "pages_per_s_host_resident": PT / ms_host * 1e3, "h2d_bytes_per_step_rank0": total_bytes,
                          "kernels_ms_rank0": dict(sorted(kt.items(), key=lambda kv: -kv[1]))}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
