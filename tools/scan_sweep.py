#!/usr/bin/env python3
"""GPU tuning aid: times the scan / decode kernels of the bench workload under different
DCSB_SCAN_LANES settings (streams per warp in the frame-boundary scan)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import torch
import dcsexplorer_b200 as dx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
streams, n_unique, src = bench.build_corpus(n, 10.0, 0)
ctx = dx.Context(0)
batch = ctx.batch(streams, os_version=dx.OS94, master_volume=255, mixing_level=0x64, tail_frames=2)
d_pcm = torch.empty(batch.total_samples, dtype=torch.int16, device="cuda")
st = torch.cuda.current_stream()
ref = None
for overlap, ctas in ((0, 3), (1, 1)):
  ctx.set_overlap(overlap)
  os.environ["DCSB_DECODE_CTAS"] = str(ctas)
  for lanes in sys.argv[2:] or ["2"]:
    if "s" in lanes:                  # "<solo>s<lanes>": slots with a warp of their own + streams per shared warp
        solo, lanes = lanes.split("s")
        os.environ["DCSB_SCAN_SOLO"] = solo
        lanes_label = solo + "s" + lanes
    else:
        os.environ.pop("DCSB_SCAN_SOLO", None)
        lanes_label = lanes
    if lanes == "0":
        os.environ.pop("DCSB_SCAN_LANES", None)
    else:
        os.environ["DCSB_SCAN_LANES"] = lanes
    ks, kd, kt = [], [], []
    for i in range(6):
        batch.decode(d_pcm.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        if i >= 2:
            ks.append(batch.kernel_ms(0)); kd.append(batch.kernel_ms(1)); kt.append(batch.kernel_ms(2))
    res = batch.results(st.cuda_stream)
    x = 0
    for r in res:
        x ^= r["checksum"]
    ref = x if ref is None else ref
    print("decode CTAs/SM %d " % ctas, end="")
    print("overlap=%d lanes=%s scan %.3f ms decode %.3f ms step %.3f ms  xor %016x %s" % (
        overlap, lanes_label, np.mean(ks), np.mean(kd), np.mean(kt), x, "OK" if x == ref else "MISMATCH"), flush=True)
