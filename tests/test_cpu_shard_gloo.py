"""N>1 host path on CPU: world_size-2 gloo processes shard a page list (no data-path collective), run
the CPU oracle on their shard, and gather results by page index — the order equals the sequential run."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _work(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from retto_b200.shard import gather_by_page, shard_indices
    from tools.synth import gen_probmap
    shapes = [(96, 128), (64, 64), (128, 160), (96, 96), (64, 200)]
    mine = shard_indices([h * w for h, w in shapes], world)[rank]
    res = []
    for i in mine:
        h, w = shapes[i]
        r = O.det_postprocess(gen_probmap(40 + i, h, w, k_range=(1, 3)), h, w)
        res.append((i, r.boxes.tobytes()))
    full = gather_by_page(res, mine, len(shapes))
    if rank == 0:
        ret.put([f[0] for f in full] + [len(f[1]) for f in full])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_preserves_page_order():
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from tools.synth import gen_probmap
    O.lib()   # build before forking workers
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_work, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shapes = [(96, 128), (64, 64), (128, 160), (96, 96), (64, 200)]
    seq = [len(O.det_postprocess(gen_probmap(40 + i, h, w, k_range=(1, 3)), h, w).boxes.tobytes()) for i, (h, w) in enumerate(shapes)]
    assert got[:5] == [0, 1, 2, 3, 4] and got[5:] == seq
