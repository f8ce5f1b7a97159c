#!/bin/bash
# Run on the GPU box (under gpurun): parity tests, bench line, ncu launch list and one
# `--set full` capture of each kernel.  Everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_profile.sh <tag> [bench args...]
set -u
TAG=${1:-run}; shift || true
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3500 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
# launch list (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 0 "$@" > $OUT/${TAG}_ncu_launch.log 2>&1
# full capture of the scan and the 1994-layout decode kernel themselves: the bench's last steps run them one after
# the other (7 scans beside the persistent decode kernel come first and are skipped)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"dcsb_(scan|decode94)_kernel" -s 7 -c 3 -f -o $OUT/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 0 "$@" > $OUT/${TAG}_ncu_full.log 2>&1
python - <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h), ("D2H", h, d)):
    a.copy_(b); torch.cuda.synchronize(); t = time.perf_counter(); a.copy_(b); torch.cuda.synchronize()
    print("pinned %s %.1f GB/s" % (name, n / (time.perf_counter() - t) / 1e9))
PY
ls -la $OUT | tail -12
