import sys, numpy as np, torch
sys.path.insert(0, '.')
from retto_b200.api import Context
from oracle import oracle as O
ctx = Context(0); ctx.enable_trace(True)
a = np.full((128, 192), 0.05, np.float32)
a[10:60, 10:180] = 0.9
a[25:45, 30:160] = 0.05
a[80:110, 20:80] = 0.9
a[93:96, 40:43] = 0.1
a[93:96, 60:64] = 0.1
g = torch.from_numpy(a).cuda(); torch.cuda.synchronize()
out = ctx.det_postprocess([g], [a.shape])
tr = ctx.fetch_trace(0)
print(tr)
print(out.page(0))
r = O.det_postprocess(a, *a.shape); print(r.boxes.reshape(-1, 8), r.status)
print(O.det_trace(a, *a.shape))
