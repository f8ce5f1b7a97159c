#!/bin/bash
# Run on the GPU box (under gpurun): parity tests, bench line, one `--set full` capture with
# source of the scan / decode kernels, pinned copy bandwidths.  usage: tools/gpu_quick.sh <tag>
set -u
TAG=${1:-run}; shift || true
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"dcsb_(scan|decode)" -s 6 -c 2 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 "$@" > $OUT/${TAG}_ncu_full.log 2>&1
python - <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h), ("D2H", h, d)):
    a.copy_(b); torch.cuda.synchronize(); t = time.perf_counter(); a.copy_(b); torch.cuda.synchronize()
    print("pinned %s %.1f GB/s" % (name, n / (time.perf_counter() - t) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("pinned H2D+D2H concurrent: %.1f GB/s each way" % (n / dt / 1e9))
PY
nproc; ls -la $OUT | tail -8
