#!/bin/bash
# Tuning build of the library with the scan's per-phase clock64 probes (-DDCSB_SCAN_DEBUG) into
# gpurun_dbg/libdcsb200_dbg.so (git-ignored; travels to the GPU box).  Used by tools/scan_probe.py.
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_dbg
C=dcsexplorer_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DDCSB_SCAN_DEBUG \
    -o gpurun_dbg/libdcsb200_dbg.so $C/dcsb_kernels.cu $C/dcsb_api.cu $C/dcsb_player.cu $C/dcsb_host.cpp $C/dcsb_rom.cpp -lz
ls -la gpurun_dbg/libdcsb200_dbg.so
