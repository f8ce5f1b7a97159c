// dcsb200 kernels (sm_100a).
//   K1 dcsb_scan_kernel    -- frame-boundary scan: one thread walks one stream's bit stream
//                             (lengths only) and emits a checkpoint {bit offset, band types}
//                             per frame, which is what makes frames independently decodable.
//   K2 dcsb_decode_kernel  -- one warp per tile of 31 consecutive output frames of a stream:
//                             lanes decode one frame each from its checkpoint into 16-bit
//                             frequency bins in shared memory, then the warp runs the exact
//                             fixed-point inverse transform frame by frame, applies the volume
//                             shift + 16-sample overlap-add, and writes PCM with coalesced
//                             32-bit stores.  Bins never touch HBM.
// The arithmetic lives in dcsb_core.cuh (shared with the CPU-side kernel simulator).
#include <cuda_runtime.h>
#include <stdint.h>
#include "dcsb_core.cuh"

#define DCSB_WARPS_PER_CTA 4

__device__ __forceinline__ void dcsb_load_lut(uint16_t *s_lut, const DcsbTables *tab)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(tab->lut);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_lut);
    for (int i = threadIdx.x; i < DCSB_LUT_WORDS / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// ------------------------------------------------------------------------------------
// K1: frame-boundary scan
__global__ void __launch_bounds__(128)
dcsb_scan_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams, int nstreams,
                 const DcsbTables *__restrict__ tab, DcsbScanOut out)
{
    __shared__ uint16_t s_lut[DCSB_LUT_WORDS];
    dcsb_load_lut(s_lut, tab);
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nstreams) return;
    dcsb_scan_stream(slab, streams, si, tab, s_lut, out);
}

// ------------------------------------------------------------------------------------
// K2: decode + transform + overlap + PCM writeback.  T93 selects the 1993 transform
// (512-bin rows) and the 1993/OS93a walkers; otherwise the 1994 path (256-bin rows).
template <bool T93>
__global__ void __launch_bounds__(DCSB_WARPS_PER_CTA * 32)
dcsb_decode_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams,
                   const DcsbTile *__restrict__ tiles, int ntiles, const DcsbTables *__restrict__ tab,
                   DcsbScanOut scan, int16_t *__restrict__ pcm, unsigned long long *__restrict__ checksums)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(smem);
    uint32_t *s_rows = smem + DCSB_LUT_WORDS / 2;
    dcsb_load_lut(s_lut, tab);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * DCSB_WARPS_PER_CTA + warp;
    if (tile >= ntiles) return;
    const DcsbTile tl = tiles[tile];
    unsigned long long csum = dcsb_decode_tile<T93>(slab, streams, tl, tab, s_lut, scan, pcm,
                                                    s_rows + warp * DcsbWarpSmem<T93>::WORDS);
    if (checksums) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
        if (lane == 0) atomicAdd(checksums + tl.stream, csum);
    }
}

// ------------------------------------------------------------------------------------
cudaError_t dcsb_launch_scan(const uint8_t *slab, const DcsbStreamRec *streams, int nstreams,
                             const DcsbTables *tables, DcsbScanOut out, cudaStream_t st)
{
    if (nstreams <= 0) return cudaSuccess;
    const int threads = 128;      // few threads per CTA spreads the latency-bound walkers over all SMs
    dcsb_scan_kernel<<<(nstreams + threads - 1) / threads, threads, 0, st>>>(slab, streams, nstreams, tables, out);
    return cudaGetLastError();
}

template <bool T93>
static cudaError_t launch_decode_t(const uint8_t *slab, const DcsbStreamRec *streams, const DcsbTile *tiles,
                                   int ntiles, const DcsbTables *tables, DcsbScanOut scan, int16_t *pcm,
                                   unsigned long long *checksums, cudaStream_t st)
{
    if (ntiles <= 0) return cudaSuccess;
    const size_t smem = (DCSB_LUT_WORDS / 2 + DCSB_WARPS_PER_CTA * DcsbWarpSmem<T93>::WORDS) * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(dcsb_decode_kernel<T93>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (ntiles + DCSB_WARPS_PER_CTA - 1) / DCSB_WARPS_PER_CTA;
    dcsb_decode_kernel<T93><<<grid, DCSB_WARPS_PER_CTA * 32, smem, st>>>(slab, streams, tiles, ntiles, tables, scan, pcm, checksums);
    return cudaGetLastError();
}

// tiles[0..ntiles94) are 1994-family tiles, tiles[ntiles94..ntiles) 1993-family tiles
cudaError_t dcsb_launch_decode(const uint8_t *slab, const DcsbStreamRec *streams, const DcsbTile *tiles,
                               int ntiles94, int ntiles93, const DcsbTables *tables, DcsbScanOut scan,
                               int16_t *pcm, unsigned long long *checksums, cudaStream_t st)
{
    cudaError_t e = launch_decode_t<false>(slab, streams, tiles, ntiles94, tables, scan, pcm, checksums, st);
    if (e != cudaSuccess) return e;
    return launch_decode_t<true>(slab, streams, tiles + ntiles94, ntiles93, tables, scan, pcm, checksums, st);
}
