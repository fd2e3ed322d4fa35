import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from retto_b200.api import Context
ctx = Context(0)
rng = np.random.default_rng(3)
p = (rng.random((200, 300)) < 0.02).astype(np.float32) * 0.9
p[50:90, 40:200] = 0.8
g = torch.from_numpy(p).cuda(); torch.cuda.synchronize()
out = ctx.det_postprocess([g], [p.shape])
lab = ctx.fetch_labels(0, *p.shape)
n = 4096
buf = np.zeros((n, 8), np.int32)
ctx._L.retto_b200_debug_comps.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]
ctx._L.retto_b200_debug_comps(ctx._h, 0, buf.ctypes.data, n)
roots = np.unique(lab[lab >= 0])
print("n roots", len(roots))
W = 300
bad = 0
off = 0
for i, r in enumerate(roots):
    ys, xs = np.nonzero(lab == r)
    exp = (int(r), int(ys.max()), int(xs.min()), int(xs.max()), off)
    got = tuple(int(v) for v in buf[i][:5])
    off += int(ys.max()) - int(r) // W + 1
    if exp != got:
        bad += 1
        if bad < 12: print(i, "exp(root,ymax,xmin,xmax,row_off)", exp, "got", got, "key", buf[i][5])
print("bad", bad)
nrows = off
rt = np.zeros((nrows, 2), np.int32)
ctx._L.retto_b200_debug_rowtab.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32]
ctx._L.retto_b200_debug_rowtab(ctx._h, 0, rt.ctypes.data, nrows)
bad = 0; off = 0
for i, r in enumerate(roots):
    ys, xs = np.nonzero(lab == r)
    y0 = int(r) // W
    for y in range(y0, int(ys.max()) + 1):
        e = (int(xs[ys == y].min()), int(xs[ys == y].max()))
        g = tuple(int(v) for v in rt[off + y - y0])
        if e != g:
            bad += 1
            if bad < 10: print("comp", i, "row", y, "exp", e, "got", g)
    off += int(ys.max()) - y0 + 1
print("rowtab bad", bad, "of", nrows)
