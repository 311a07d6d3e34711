"""Device log-softmax (kab_softmax.cuh) alone: config-2-sized rows, CUDA events, GB/s against the HBM peak.
    python tools/softmax_bench.py [rows] [V]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from kokoro_align_b200 import align  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_758_460
V = int(sys.argv[2]) if len(sys.argv) > 2 else 39
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
x = torch.randn(rows, V, device="cuda") * 3
y = torch.empty_like(x)
for inplace in (False, True):
    dst = x.clone() if inplace else y
    src = dst if inplace else x
    for _ in range(3):
        align.log_softmax_torch(src, out=dst)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        align.log_softmax_torch(src, out=dst)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gbs = 8.0 * rows * V / ms / 1e6
    print(json.dumps({"rows": rows, "V": V, "in_place": inplace, "ms": ms, "algorithmic_gbs": gbs, "hbm_frac": gbs / peak}))
ref = torch.log_softmax(x.double(), -1)
print("max |device - fp64 log_softmax| =", float((y.double() - ref).abs().max()))
