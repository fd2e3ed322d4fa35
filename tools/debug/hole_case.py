import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import oracle as O
from tools.demo_worker import StatelessWorker
from tools.synth import gen_page
from retto_b200.api import Context
w = StatelessWorker()
ctx = Context(0)
for s in (7, 21, 25, 26, 28, 30):
    img = gen_page(4 + s, 1280, 1280)[0]
    pred = np.ascontiguousarray(w.det(O.det_preprocess(img))[0, 0])
    ref = O.det_postprocess(pred, 1280, 1280, want_bitmap=True)
    g = torch.from_numpy(pred).cuda(); torch.cuda.synchronize()
    ctx.enable_trace(True)
    out = ctx.det_postprocess([g], [(1280, 1280)])
    boxes, scores = out.page(0)
    tr = ctx.fetch_trace(0)
    print("page", s, "ref", len(ref.boxes), "gpu", len(boxes), "n_holes", tr["n_holes"], "status", out.page_status[0])
    A = {tuple(b.reshape(-1).tolist()) for b in ref.boxes}; B = {tuple(b.reshape(-1).tolist()) for b in boxes}
    print("  only ref", sorted(A - B)[:5]); print("  only gpu", sorted(B - A)[:5])
    contours = O.find_contours(ref.bitmap)
    n_hole = sum(h for _, h in contours)
    print("  oracle contours", len(contours), "holes", n_hole, " true holes(quirk0)", sum(h for _, h in O.find_contours(ref.bitmap, quirk_x0=0)))
    for bx in sorted(B - A)[:3]:
        xs, ys = bx[0::2], bx[1::2]
        x0, x1, y0, y1 = int(min(xs)) - 3, int(max(xs)) + 4, int(min(ys)) - 3, int(max(ys)) + 4
        x0, y0 = max(x0, 0), max(y0, 0)
        print("  bitmap around extra gpu box", bx)
        for y in range(y0, min(y1, 1280)):
            print("   ", "".join("#" if ref.bitmap[y, x] else "." for x in range(x0, min(x1, 1280))))
