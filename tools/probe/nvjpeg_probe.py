"""Runs on the GPU box: renders pages, encodes them as JPEG (PIL / libjpeg-turbo), runs the nvJPEG probe,
compares the decoded pixels with PIL's and cv2's decoders."""
import io, os, struct, subprocess, sys
import numpy as np
from PIL import Image
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.synth import gen_page
out = "gpurun_out/nvjpeg"; os.makedirs(out, exist_ok=True)
pages = [gen_page(100 + i, 1280, 1280)[0] for i in range(8)]
for tag, kw in (("q90_420", dict(quality=90, subsampling=2)), ("q90_444", dict(quality=90, subsampling=0)), ("q95_444", dict(quality=95, subsampling=0))):
    blobs = []
    for i in range(128):
        b = io.BytesIO(); Image.fromarray(pages[i % 8]).save(b, "JPEG", **kw); blobs.append(b.getvalue())
    with open(f"{out}/{tag}.bin", "wb") as f:
        f.write(struct.pack("<I", len(blobs))); f.write(struct.pack(f"<{len(blobs)}I", *map(len, blobs)))
        for b in blobs: f.write(b)
    print(tag, "mean bytes", np.mean([len(b) for b in blobs]), flush=True)
    r = subprocess.run(["tools/probe/nvjpeg_probe", f"{out}/{tag}.bin", f"{out}/{tag}"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr[-2000:], flush=True)
    ref = np.asarray(Image.open(io.BytesIO(blobs[0])).convert("RGB")).astype(np.int32)
    import cv2
    ref2 = cv2.cvtColor(cv2.imdecode(np.frombuffer(blobs[0], np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB).astype(np.int32)
    print(" PIL vs cv2 maxdiff", np.abs(ref - ref2).max())
    for nm in sorted(os.listdir(out)):
        if nm.startswith(tag) and nm.endswith(".rgb"):
            got = np.fromfile(f"{out}/{nm}", np.uint8).reshape(1280, 1280, 3).astype(np.int32)
            d = np.abs(got - ref)
            print(f" {nm}: vs PIL maxdiff {d.max()} mean {d.mean():.4f} frac!=0 {(d > 0).mean():.4f}; vs orig page max {np.abs(got - pages[0]).max()}")
            os.remove(f"{out}/{nm}")
    os.remove(f"{out}/{tag}.bin")
