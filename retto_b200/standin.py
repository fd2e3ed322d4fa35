"""Stand-in forward passes of PP-OCRv4-mobile I/O shape for machines without
onnxruntime / the ONNX weights: random-init torch modules that are really EXECUTED on the tensors the forward seam hands over
(`retto_b200_forward_fn`: device pointers into the context's own buffers), so the seam — zero-copy binding, stream ordering, the
launch pattern of a worker between the pre- and post-processing kernels — is exercised with a realistic amount of GPU work in
the middle.  They are NOT the product and NOT a model: the reference runs the real networks through ONNX Runtime
(retto-core/src/worker/ort_worker.rs:188-221) and an IoBinding worker plugs into the same seam.

    det  [1,3,H,W]    -> [1,1,H,W]     sigmoid   (DBNet: light backbone to 1/32, FPN-lite at 1/4, two transposed convs back to full size)
    cls  [n,3,48,192] -> [n,2]         softmax
    rec  [n,3,48,W]   -> [n,W/8,6625]  softmax   (height collapsed, width / 8, linear to the dictionary size)
"""
from __future__ import annotations


def _dw(nn, cin, cout, stride):
    return nn.Sequential(nn.Conv2d(cin, cin, 3, stride, 1, groups=cin, bias=False), nn.Conv2d(cin, cout, 1, bias=True), nn.Hardswish())


class StandInNets:
    def __init__(self, torch, device, n_classes=6625, seed=0):
        nn = torch.nn
        self.torch = torch
        g = torch.Generator().manual_seed(seed)  # noqa: F841  (module init uses the global RNG; seeded below)
        torch.manual_seed(seed)

        class Det(nn.Module):
            def __init__(s):
                super().__init__()
                s.stem = nn.Sequential(nn.Conv2d(3, 16, 3, 2, 1), nn.Hardswish())
                s.s4, s.s8, s.s16, s.s32 = _dw(nn, 16, 24, 2), _dw(nn, 24, 40, 2), _dw(nn, 40, 80, 2), _dw(nn, 80, 112, 2)
                s.l4, s.l8, s.l16, s.l32 = nn.Conv2d(24, 24, 1), nn.Conv2d(40, 24, 1), nn.Conv2d(80, 24, 1), nn.Conv2d(112, 24, 1)
                s.fuse = nn.Sequential(nn.Conv2d(96, 24, 3, 1, 1), nn.ReLU())
                s.up1 = nn.Sequential(nn.ConvTranspose2d(24, 6, 2, 2), nn.ReLU())
                s.up2 = nn.ConvTranspose2d(6, 1, 2, 2)

            def forward(s, x):
                F = torch.nn.functional
                c4 = s.s4(s.stem(x))
                c8 = s.s8(c4)
                c16 = s.s16(c8)
                c32 = s.s32(c16)
                hw = c4.shape[-2:]
                f = torch.cat([s.l4(c4), F.interpolate(s.l8(c8), size=hw), F.interpolate(s.l16(c16), size=hw), F.interpolate(s.l32(c32), size=hw)], 1)
                return torch.sigmoid(s.up2(s.up1(s.fuse(f))))

        class Cls(nn.Module):
            def __init__(s):
                super().__init__()
                s.f = nn.Sequential(nn.Conv2d(3, 16, 3, 2, 1), nn.Hardswish(), _dw(nn, 16, 32, 2), _dw(nn, 32, 64, 2), _dw(nn, 64, 96, 2))
                s.fc = nn.Linear(96, 2)

            def forward(s, x):
                return torch.softmax(s.fc(s.f(x).mean((2, 3))), 1)

        class Rec(nn.Module):
            def __init__(s):
                super().__init__()
                s.f = nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1), nn.Hardswish(), _dw(nn, 32, 64, 2), _dw(nn, 64, 128, 2),
                                    nn.Conv2d(128, 192, (6, 1)), nn.Hardswish())
                s.fc = nn.Linear(192, n_classes)

            def forward(s, x):
                y = s.f(x)                                # [n,192,1,W/8]
                return torch.softmax(s.fc(y.squeeze(2).transpose(1, 2)), 2)

        self.nets = [Det().to(device).eval(), Cls().to(device).eval(), Rec().to(device).eval()]
        torch.backends.cudnn.benchmark = False

    def n_params(self):
        return {k: int(sum(p.numel() for p in n.parameters())) for k, n in zip(("det", "cls", "rec"), self.nets)}

    def run(self, stage, xs):
        """forward of every input tensor of the call (no concatenation: the inputs are read where the context put them)"""
        net = self.nets[stage]
        outs = []
        for x in xs:
            outs.append(net(x))
        return outs
