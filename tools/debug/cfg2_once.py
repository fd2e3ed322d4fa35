import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from retto_b200.api import Context
from tools.synth import gen_probmap
ctx = Context(0)
uniq = [torch.from_numpy(gen_probmap(2000 + i, 960, 960)).cuda() for i in range(16)]
maps = [uniq[i % 16].clone() for i in range(256)]
torch.cuda.synchronize()
for _ in range(2):
    ctx.det_postprocess(maps, [(960, 960)] * 256, max_boxes_total=256 * 80)
