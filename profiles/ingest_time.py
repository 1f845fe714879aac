"""Device time of the frame ingest (loans_stn_ingest_u8) over rotating buffers, CUDA-graph replay: resize 384x512 -> 224x224,
the conversion alone at 224x224, and a 1080p source.  usage: ingest_time.py [path/to/libloans_stn.so] (default: the product)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from loans_b200 import _lib  # noqa: E402
from loans_b200.functions import FrameIngest  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])                       # an A/B build (profiles/ab_local.sh)

dev = torch.device("cuda", 0)
CASES = [("384x512->224x224 x64", 64, (384, 512), (224, 224)), ("224x224 convert x64", 64, (224, 224), None),
         ("1080x1920->224x224 x16", 16, (1080, 1920), (224, 224)), ("512x512->224x224 x64", 64, (512, 512), (224, 224))]
for name, b, src, dst in CASES:
    S = 4
    raw = [torch.randint(0, 256, (b,) + src + (3,), dtype=torch.uint8, device=dev) for _ in range(S)]
    oh, ow = dst or src
    outs = [torch.empty((b, 3, oh, ow), device=dev) for _ in range(S)]
    ing = FrameIngest(b, src, dst, device=dev)
    ing(raw[0], out=outs[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(S):
            ing(raw[i], out=outs[i])
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (20 * S)
    nb = b * src[0] * src[1] * 3 + 4 * b * 3 * oh * ow
    print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "case": name, "us": round(us, 2), "algorithmic_bytes": nb, "gbs": round(nb / us / 1e3, 1)}), flush=True)
