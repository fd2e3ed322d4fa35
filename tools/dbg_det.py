"""debug aid: per-contour trace of the CUDA det_postprocess vs the oracle (prints mismatches only)"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from retto_b200.api import Context
from oracle import oracle as O
from tools.synth import gen_probmap


def compare(ctx, probs, verbose=False):
    O.set_libm(1)
    gs = [torch.from_numpy(p).cuda() for p in probs]; torch.cuda.synchronize()
    out = ctx.det_postprocess(gs, [p.shape for p in probs])
    for i, p in enumerate(probs):
        tr = ctx.fetch_trace(i)
        r1, ss, sc, st = O.det_trace(p, *p.shape)
        order = [o for o in np.argsort(tr['key'], kind='stable') if tr['status'][o] != 6]
        print("page", i, "status", out.page_status[i], "gpu comps", len(order), "oracle contours", len(st), "holes", tr['n_holes'])
        for j, o in enumerate(order):
            g = (tr['rect1'][o].tolist(), float(tr['sside1'][o]), float(tr['score'][o]), int(tr['status'][o]))
            if j < len(st):
                r = (r1[j].tolist(), float(ss[j]), float(sc[j]), int(st[j]))
                same = (g[0] == r[0] and g[3] == r[3] and (g[2] == r[2] or (g[2] != g[2] and r[2] != r[2])))
                if verbose or not same:
                    print("  ", j, "key", tr['key'][o], g, r, "" if same else "  <<<<")
        b, s = out.page(i)
        ref = O.det_postprocess(p, *p.shape)
        if not np.array_equal(b, ref.boxes):
            print("BOX DIFF page", i); print(b.reshape(-1, 8)); print(ref.boxes.reshape(-1, 8))


if __name__ == "__main__":
    ctx = Context(0); ctx.enable_trace(True)
    rng = np.random.default_rng(3)
    p = (rng.random((200, 300)) < 0.02).astype(np.float32) * 0.9
    p[50:90, 40:200] = 0.8
    compare(ctx, [p])
