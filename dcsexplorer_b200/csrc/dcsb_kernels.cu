// dcsb200 kernels (sm_100a).
//   K1 dcsb_scan_kernel           -- frame-boundary scan: one thread walks one stream's bit stream
//                                    (lengths only) and emits a checkpoint {bit offset, band types}
//                                    per frame, which is what makes frames independently decodable
//                                    (dcsb_scan94.cuh for the 1994+ layout, dcsb_core.cuh for the others).
//   K2+K3 dcsb_decode94_kernel /  -- 1994+ layout: one warp per work item of consecutive output frames,
//         dcsb_decode94_queue_kernel one LANE per frame: bit unpack from the checkpoint, dequantise, exact
//                                    fixed-point inverse transform in the lane's shared-memory row,
//                                    overlap-add, coalesced PCM stores (dcsb_fast94.cuh).  The queue
//                                    variant is persistent and takes its items from the scan running
//                                    beside it.
//         dcsb_decode_kernel<true>  -- 1993 layouts: one warp per tile of 31 output frames, lanes decode
//                                    a frame each into shared memory, the warp transforms frame by frame.
//   K4 dcsb_mix94/93_kernel       -- track playback: the channels of the host's mix schedule decoded in
//                                    order into the same bins, one transform per output frame (dcsb_mix.cuh).
// Bins never touch HBM.  The arithmetic lives in the .cuh files, shared with the CPU-side kernel simulator.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "dcsb_core.cuh"
#include "dcsb_fast94.cuh"
#include "dcsb_scan94.cuh"
#include "dcsb_mix.cuh"
#include "dcsb_seq.cuh"

#define DCSB_WARPS_PER_CTA 2      // 1993 layouts: 33 KB of rows per warp; 70 KB CTAs: three per SM, or one beside a scan CTA

__device__ __forceinline__ void dcsb_load_lut(uint16_t *s_lut, const DcsbTables *tab)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(tab->lut);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_lut);
    for (int i = threadIdx.x; i < DCSB_LUT_WORDS / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// ------------------------------------------------------------------------------------
// K1: frame-boundary scan.  A warp takes 32 streams of the scan order (alike streams side by
// side, dcsb_scan_order) and walks them in LOCK STEP, one lane per stream (dcsb_scan94.cuh); a
// CTA holds `warps` such warps (two in the single-wave case: the CTA then fills its SM's shared
// memory and the SMs it does not need run three decode CTAs each, dcsb_scan_shape) that share
// the tables and take stream groups grid-stride.  Streams of the 1993 layouts
// are walked by their lane alone first (dcsb_scan_stream), the lane then idles in the lock-step
// walk of its warp.
// Shared memory: the length table tx (96 KB) sits on a 16 KB boundary of the shared window
// (table | index is one LOP3), the rings follow it (1 KB each, 1 KB aligned: ring | offset), then
// 18 band entries of 16 bytes per lane; the peek LUTs, the band-descriptor table and the zero
// word go into the alignment gap in front of the table or behind everything, whichever is large enough.
#define DCSB_TX_BYTES (6 * DCSB_T8_CB * 4)
#define DCSB_SCAN_LUT_BYTES ((DCSB_LUT_WORDS * 2 + 15) & ~15)
#define DCSB_SCAN_SMALL (DCSB_SCAN_LUT_BYTES + DCSB_DTAB_WORDS * 2 + 16)
#define DCSB_SCAN_ENT_STRIDE (19u * 16u)        // 18 entries + 16 bytes: lanes start 12 banks apart, LDS.128 conflict-free
#define DCSB_SCAN_WARP_BYTES(ring) ((ring ? 32u * DCSB_RING_BYTES : 0u) + 32u * DCSB_SCAN_ENT_STRIDE)
#define DCSB_SCAN_SMEM(warps, ring) (16384u + DCSB_TX_BYTES + (warps) * DCSB_SCAN_WARP_BYTES(ring))
static_assert(DCSB_SCAN_SMALL <= 8192, "the small tables must fit the smaller alignment gap");
static_assert(DCSB_SCAN_SMEM(DCSB_SCAN_MAXWARPS, 1) <= 232448, "scan CTA exceeds the shared memory of an SM");
static_assert(DCSB_SCAN_SMEM(DCSB_SCAN_MAXWARPS_DIRECT, 0) <= 232448, "scan CTA exceeds the shared memory of an SM");

// RING: the lanes' stream bytes staged in shared-memory rings (single wave) / read from global memory
// through L1 (many waves: more warps per SM), see DcsbWinT.
template <bool RING>
__global__ void __launch_bounds__((RING ? DCSB_SCAN_MAXWARPS : DCSB_SCAN_MAXWARPS_DIRECT) * 32, 1)
dcsb_scan_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams, const uint32_t *__restrict__ order,
                 int nstreams, const DcsbTables *__restrict__ tab, DcsbScanOut out, uint32_t f0, uint32_t f1)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint8_t *sm8 = reinterpret_cast<uint8_t *>(smem);
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tx = (s_base + 16383u) & ~16383u;
    const uint32_t tx_off = tx - s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const uint32_t rings_off = tx_off + DCSB_TX_BYTES;                          // 16 KB aligned in the shared window
    const uint32_t ents_off = rings_off + (RING ? (uint32_t)warps * 32u * DCSB_RING_BYTES : 0u);
    const uint32_t small_off = tx_off >= DCSB_SCAN_SMALL ? 0u : ents_off + (uint32_t)warps * 32u * DCSB_SCAN_ENT_STRIDE;
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(sm8 + small_off);
    uint16_t *s_dtab = reinterpret_cast<uint16_t *>(sm8 + small_off + DCSB_SCAN_LUT_BYTES);
    const uint32_t zero_off = small_off + DCSB_SCAN_LUT_BYTES + DCSB_DTAB_WORDS * 2;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(tab->tx);
        uint4 *dst = reinterpret_cast<uint4 *>(sm8 + tx_off);
        for (int i = threadIdx.x; i < DCSB_TX_BYTES / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x == 0) {
        *reinterpret_cast<uint4 *>(sm8 + zero_off) = make_uint4(0, 0, 0, 0);
        if (out.started) atomicAdd(out.started, 1u);       // this CTA is resident (dcsb_gate_kernel)
    }
    dcsb_load_lut(s_lut, tab);
    for (int i = threadIdx.x; i < DCSB_DTAB_WORDS; i += blockDim.x) s_dtab[i] = (uint16_t)dcsb_dtab_entry(s_lut, i);
    __syncthreads();
    const uint32_t ring_off = rings_off + ((uint32_t)warp * 32u + (uint32_t)lane) * DCSB_RING_BYTES;
    const uint32_t ent_off = ents_off + ((uint32_t)warp * 32u + (uint32_t)lane) * DCSB_SCAN_ENT_STRIDE;
#if DCSB_DEVICE_PASS
    const DcsbRingPtr ring = s_base + ring_off;
    const DcsbTxBase txb = tx;
    const DcsbSA ents = s_base + ent_off, zero = s_base + zero_off;
#else       // (nvcc's host pass only type-checks this body)
    const DcsbRingPtr ring = reinterpret_cast<DcsbSA>(sm8 + ring_off);
    const DcsbTxBase txb = reinterpret_cast<DcsbSA>(sm8 + tx_off);
    const DcsbSA ents = reinterpret_cast<DcsbSA>(sm8 + ent_off), zero = reinterpret_cast<DcsbSA>(sm8 + zero_off);
#endif
    const int ngroups = (nstreams + 31) >> 5;
    for (int g = blockIdx.x * warps + warp; g < ngroups; g += gridDim.x * warps) {
        const int k = g * 32 + lane;
        int si = -1;
        if (k < nstreams) si = order ? (int)order[k] : k;
        if (si >= 0 && streams[si].fmt != DCSB_FMT_94) {
            dcsb_scan_stream(slab, streams, si, tab, s_lut, out, f0, f1);
            si = -1;
        }
        __syncwarp();
        dcsb_scan94_stream<RING>(slab, streams, si, tab, s_lut, txb, s_dtab, ring, ents, zero, out, f0, f1);
        __syncwarp();
    }
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
// K2, 1994 layout: one warp per work item, one lane per frame (dcsb_fast94.cuh)
#define DCSB_WARPS94 4
struct DcsbSmem94 {
    uint16_t lut[DCSB_LUT_WORDS];
    DcsbTw94 tw;
    uint8_t hdr[DCSB_WARPS94][16];
    uint32_t rows[DCSB_WARPS94][DCSB_WARP94_WORDS];
};

__global__ void __launch_bounds__(DCSB_WARPS94 * 32, 3)
dcsb_decode94_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams,
                     const DcsbTile *__restrict__ items, int nitems, const DcsbTables *__restrict__ tab,
                     DcsbScanOut scan, int16_t *__restrict__ pcm, unsigned long long *__restrict__ checksums)
{
    extern __shared__ __align__(16) uint32_t smem[];
    DcsbSmem94 &sm = *reinterpret_cast<DcsbSmem94 *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * DCSB_WARPS94 + warp;
    {
        for (int i = threadIdx.x; i < 64; i += blockDim.x) dcsb_tw94_fill(&sm.tw, tab, i);
        if (item < nitems && lane < 16) sm.hdr[warp][lane] = streams[items[item].stream].hdr[lane];
    }
    dcsb_load_lut(sm.lut, tab);
    if (item >= nitems) return;
    const DcsbTile it = items[item];
    // checkpoints first-1 .. first+count (a frame's own band types are the NEXT entry)
    const uint32_t nfr = streams[it.stream].nframes;
    const bool fin = dcsb_await(scan.progress, it.stream, it.first + it.count + 1 < nfr + 1 ? it.first + it.count + 1 : nfr + 1, scan.qctl ? scan.qctl + 2 : nullptr);
    const uint32_t nplay = fin ? __ldcg(scan.nplay + it.stream) : nfr;
    const int stopband = fin ? __ldcg(scan.stopband + it.stream) : 0xFF;
    unsigned long long csum = dcsb_decode94_item(slab, streams, it, tab, sm.lut, &sm.tw, sm.hdr[warp], scan, nplay, stopband, pcm, sm.rows[warp]);
    if (checksums) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
        if (lane == 0) atomicAdd(checksums + it.stream, csum);
    }
}

// K2, 1994 layout, overlapped mode: persistent warps take work items from the queue the scan
// running beside them fills (dcsb_queue_push), in the order the checkpoints become available.
__global__ void __launch_bounds__(DCSB_WARPS94 * 32, 3)
dcsb_decode94_queue_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams, int nitems,
                           const DcsbTables *__restrict__ tab, DcsbScanOut scan, int16_t *__restrict__ pcm,
                           unsigned long long *__restrict__ checksums)
{
    extern __shared__ __align__(16) uint32_t smem[];
    DcsbSmem94 &sm = *reinterpret_cast<DcsbSmem94 *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        for (int i = threadIdx.x; i < 64; i += blockDim.x) dcsb_tw94_fill(&sm.tw, tab, i);
    }
    dcsb_load_lut(sm.lut, tab);
#ifdef DCSB_SCAN_DEBUG
    unsigned long long t_first = 0, t_last = 0, t_wait = 0, n_items = 0;
#define DCSB_NOW(v) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v))
#endif
    for (;;) {
        unsigned long long e = 0;
#ifdef DCSB_SCAN_DEBUG
        unsigned long long t_a, t_b;
        DCSB_NOW(t_a);
#endif
        if (lane == 0) {
            const uint32_t q = atomicAdd(scan.qctl + 1, 1u);
            if (q < (uint32_t)nitems) {
                const long long t0 = clock64();
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(e) : "l"(scan.queue + q) : "memory");
                    if (e & DCSB_Q_VALID) break;
                    if (clock64() - t0 > 4000000000ll) {     // never hang the GPU on a lost producer; the host reports DCSB_E_CUDA
                        atomicOr(scan.qctl + 2, DCSB_DEV_E_QUEUE);
                        e = 0;
                        break;
                    }
                    __nanosleep(256);
                }
            }
        }
        e = __shfl_sync(0xffffffffu, e, 0);
#ifdef DCSB_SCAN_DEBUG
        DCSB_NOW(t_b);
        if (!(e & DCSB_Q_VALID) && lane == 0 && scan.dbg) {
            unsigned long long *d = reinterpret_cast<unsigned long long *>(scan.dbg) + 4ull * (blockIdx.x * DCSB_WARPS94 + warp);
            d[0] = t_first; d[1] = t_last; d[2] = t_wait; d[3] = n_items;
        }
        if (e & DCSB_Q_VALID) { t_wait += t_b - t_a; if (!t_first) t_first = t_b; ++n_items; }
#endif
        if (!(e & DCSB_Q_VALID)) return;
        DcsbTile it;
        it.stream = (uint32_t)((e >> 24) & 0x3FFFFFFFull);
        it.first = (uint32_t)(e & 0xFFFFFFull);
        const DcsbStreamRec *sp = streams + it.stream;
        it.count = sp->out_frames - it.first < DCSB_QITEM ? sp->out_frames - it.first : DCSB_QITEM;
        const bool fin = (e & DCSB_Q_FINAL) != 0;
        const uint32_t nplay = fin ? __ldcg(scan.nplay + it.stream) : sp->nframes;
        const int stopband = fin ? __ldcg(scan.stopband + it.stream) : 0xFF;
        if (lane < 16) sm.hdr[warp][lane] = sp->hdr[lane];
        __syncwarp();
        unsigned long long csum = dcsb_decode94_item(slab, streams, it, tab, sm.lut, &sm.tw, sm.hdr[warp], scan, nplay, stopband, pcm, sm.rows[warp]);
        if (checksums) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            if (lane == 0) atomicAdd(checksums + it.stream, csum);
        }
        __syncwarp();
#ifdef DCSB_SCAN_DEBUG
        DCSB_NOW(t_last);
#endif
    }
}

static int decode_queue_grid(int nitems)
{
    int grid = (nitems + DCSB_WARPS94 - 1) / DCSB_WARPS94;
    // persistent CTAs per SM: as many as fit (three of 70 KB; beside a resident scan CTA of 150 KB the
    // SM takes one, the others follow when the scan CTA has left)
    const int sms = dcsb_num_sms();
    int per_sm = 3;
    if (const char *e = getenv("DCSB_DECODE_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= 3) per_sm = v; }   // tuning override
    if (grid > sms * per_sm) grid = sms * per_sm;
    return grid;
}

void dcsb_decode_shapes(int nitems, int *grid_items, int *grid_queue, int *block)
{
    *grid_items = (nitems + DCSB_WARPS94 - 1) / DCSB_WARPS94;
    *grid_queue = nitems > 0 ? decode_queue_grid(nitems) : 0;
    *block = DCSB_WARPS94 * 32;
}

cudaError_t dcsb_launch_decode_queue(const uint8_t *slab, const DcsbStreamRec *streams, int nstreams, int nitems,
                                     const DcsbTables *tables, DcsbScanOut scan, int16_t *pcm,
                                     unsigned long long *checksums, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    const size_t smem = sizeof(DcsbSmem94);
    cudaError_t e = cudaFuncSetAttribute(dcsb_decode94_queue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // keep the SM at its largest shared-memory split, so that CTAs of the other kernel can join
    // this one (the split only changes on an idle SM)
    e = cudaFuncSetAttribute(dcsb_decode94_queue_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    (void)nstreams;
    const int grid = decode_queue_grid(nitems);
    dcsb_decode94_queue_kernel<<<grid, DCSB_WARPS94 * 32, smem, st>>>(slab, streams, nitems, tables, scan, pcm, checksums);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K4: channel mix for track playback (dcsb_mix.cuh).  Same shapes as the single-stream decode
// kernels: 1994 layout one lane per output frame, 1993 layouts one warp-cooperative tile.
struct DcsbSmemMix94 {
    uint16_t lut[DCSB_LUT_WORDS];
    DcsbTw94 tw;
    uint8_t hdr[DCSB_WARPS94][32][16];
    uint32_t rows[DCSB_WARPS94][DCSB_WARP94_WORDS];
};

__global__ void __launch_bounds__(DCSB_WARPS94 * 32, 3)
dcsb_mix94_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams,
                  const DcsbMixItem *__restrict__ items, int nitems, DcsbMixSched sched, const DcsbTables *__restrict__ tab,
                  DcsbScanOut scan, int16_t *__restrict__ pcm, unsigned long long *__restrict__ checksums)
{
    extern __shared__ __align__(16) uint32_t smem[];
    DcsbSmemMix94 &sm = *reinterpret_cast<DcsbSmemMix94 *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * DCSB_WARPS94 + warp;
    {
        for (int i = threadIdx.x; i < 64; i += blockDim.x) dcsb_tw94_fill(&sm.tw, tab, i);
    }
    dcsb_load_lut(sm.lut, tab);
    if (item >= nitems) return;
    const DcsbMixItem it = items[item];
    unsigned long long csum = dcsb_mix94_item(slab, streams, it, sched, tab, sm.lut, &sm.tw, scan, pcm, sm.rows[warp], &sm.hdr[warp][0][0]);
    if (checksums) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
        if (lane == 0) atomicAdd(checksums + it.timeline, csum);
    }
}

__global__ void __launch_bounds__(DCSB_WARPS_PER_CTA * 32)
dcsb_mix93_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams,
                  const DcsbMixItem *__restrict__ items, int nitems, DcsbMixSched sched, const DcsbTables *__restrict__ tab,
                  DcsbScanOut scan, int16_t *__restrict__ pcm, unsigned long long *__restrict__ checksums)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(smem);
    uint32_t *s_rows = smem + DCSB_LUT_WORDS / 2;
    dcsb_load_lut(s_lut, tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * DCSB_WARPS_PER_CTA + warp;
    if (item >= nitems) return;
    const DcsbMixItem it = items[item];
    unsigned long long csum = dcsb_mix93_tile(slab, streams, it, sched, tab, s_lut, scan, pcm, s_rows + warp * DcsbWarpSmem<true>::WORDS);
    if (checksums) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
        if (lane == 0) atomicAdd(checksums + it.timeline, csum);
    }
}

cudaError_t dcsb_launch_mix(bool family93, const uint8_t *slab, const DcsbStreamRec *streams, const void *items, int nitems,
                            const void *frames, const void *entries, const DcsbTables *tables, DcsbScanOut scan,
                            int16_t *pcm, unsigned long long *checksums, cudaStream_t st)
{
    if (nitems <= 0) return cudaSuccess;
    DcsbMixSched sc{ static_cast<const DcsbSchedFrame *>(frames), static_cast<const DcsbSchedEntry *>(entries) };
    const DcsbMixItem *it = static_cast<const DcsbMixItem *>(items);
    if (family93) {
        const size_t smem = (DCSB_LUT_WORDS / 2 + DCSB_WARPS_PER_CTA * DcsbWarpSmem<true>::WORDS) * sizeof(uint32_t);
        cudaError_t e = cudaFuncSetAttribute(dcsb_mix93_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // keep the SM at its largest shared-memory split, so that CTAs of the other kernel can join
        // this one (the split only changes on an idle SM)
        e = cudaFuncSetAttribute(dcsb_mix93_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        const int grid = (nitems + DCSB_WARPS_PER_CTA - 1) / DCSB_WARPS_PER_CTA;
        dcsb_mix93_kernel<<<grid, DCSB_WARPS_PER_CTA * 32, smem, st>>>(slab, streams, it, nitems, sc, tables, scan, pcm, checksums);
    } else {
        const size_t smem = sizeof(DcsbSmemMix94);
        cudaError_t e = cudaFuncSetAttribute(dcsb_mix94_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // keep the SM at its largest shared-memory split, so that CTAs of the other kernel can join
        // this one (the split only changes on an idle SM)
        e = cudaFuncSetAttribute(dcsb_mix94_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        const int grid = (nitems + DCSB_WARPS94 - 1) / DCSB_WARPS94;
        dcsb_mix94_kernel<<<grid, DCSB_WARPS94 * 32, smem, st>>>(slab, streams, it, nitems, sc, tables, scan, pcm, checksums);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K5: device-side track interpreter.  One thread = one decoder instance: it runs the
// reference's MainLoop control plane (dcsb_seq.cuh: command queue, track byte code, mixer, fades, data-port
// state machine, gain staging) frame after frame against the ROM image in HBM and writes the mix schedule K4
// renders -- dcsb_render_timelines needs no host sequencer threads and no schedule upload.  The instance's
// state (about 4 KB) lives in the thread's local memory; the 32 instances of a warp run their own track
// programs (they diverge where their programs do).
__global__ void __launch_bounds__(32)
dcsb_seq_kernel(DcsbRomView rv, const DcsbSeqTimeline *__restrict__ tls, int n, const dcsb_port_write *__restrict__ writes,
                DcsbSchedFrame *__restrict__ frames, DcsbSchedEntry *__restrict__ entries, uint32_t *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const DcsbSeqTimeline tl = tls[t];
    DcsbSeqState st;
    dcsb_seq_init(st);
    dcsb_seq_set_master_volume(st, (int)tl.master_volume);
    uint32_t w = 0;
    for (uint32_t f = 0; f < tl.n_frames; ++f) {
        while (w < tl.n_writes && writes[tl.first_write + w].frame <= f) dcsb_seq_write_port(st, rv, writes[tl.first_write + w++].byte);
        DcsbSchedFrame fr;
        DcsbSchedEntry e[DCSB_MAX_CHANNELS];
        dcsb_seq_frame(st, rv, &fr, e);
        const uint32_t gf = tl.first_frame + f;
        fr.first_entry = gf * DCSB_MAX_CHANNELS;
        for (int i = 0; i < fr.n_entries; ++i) entries[fr.first_entry + i] = e[i];
        frames[gf] = fr;
        st.host_n = 0;
    }
    out[2 * t] = st.fatal ? 1u : 0u;
    out[2 * t + 1] = st.host_total;
}

cudaError_t dcsb_launch_seq(const DcsbRomView *view, const DcsbSeqTimeline *tls, int n, const void *writes,
                            void *frames, void *entries, uint32_t *out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    // The interpreter is a chain of dependent loads per timeline and the lanes of a warp serialise wherever their
    // programs part: narrow CTAs spread the timelines over every SM's schedulers and keep the serialisation short.
    static const int block = [] { const char *e = getenv("DCSB_SEQ_BLOCK"); const int b = e ? atoi(e) : 4; return b < 1 ? 1 : (b > 32 ? 32 : b); }();
    dcsb_seq_kernel<<<(n + block - 1) / block, block, 0, st>>>(*view, tls, n, static_cast<const dcsb_port_write *>(writes),
                                                  static_cast<DcsbSchedFrame *>(frames), static_cast<DcsbSchedEntry *>(entries), out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
static void scan_shape(int nstreams, int concurrent, int &warps, int &grid) { dcsb_scan_shape(nstreams, concurrent, &warps, &grid); }

// K1 for the 1993 layouts: one lane walks one stream by itself (dcsb_scan_stream: the four layouts share no
// control flow worth keeping in step), `lanes` streams to a warp -- the lanes of a warp serialise wherever their
// streams part, so a warp holds few of them and a CTA holds several such warps around one copy of the peek tables.
#define DCSB_SCAN93_MAXWARPS 16
__global__ void __launch_bounds__(DCSB_SCAN93_MAXWARPS * 32)
dcsb_scan93_kernel(const uint8_t *__restrict__ slab, const DcsbStreamRec *__restrict__ streams, const uint32_t *__restrict__ order,
                   int nstreams, int lanes, const DcsbTables *__restrict__ tab, DcsbScanOut out, uint32_t f0, uint32_t f1)
{
    __shared__ __align__(16) uint16_t s_lut[DCSB_LUT_WORDS];
    if (threadIdx.x == 0 && out.started) atomicAdd(out.started, 1u);       // this CTA is resident (dcsb_gate_kernel)
    dcsb_load_lut(s_lut, tab);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    if (lane >= lanes) return;
    for (int k = (blockIdx.x * warps + warp) * lanes + lane; k < nstreams; k += gridDim.x * warps * lanes)
        dcsb_scan_stream(slab, streams, order ? (int)order[k] : k, tab, s_lut, out, f0, f1);
}

static void scan93_shape(int n93, int &lanes, int &warps, int &grid)
{
    static const int env = [] { const char *e = getenv("DCSB_SCAN93_LANES"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 32 ? 32 : v); }();
    lanes = env;
    const int sms = dcsb_num_sms(), groups = (n93 + lanes - 1) / lanes;
    warps = (groups + sms - 1) / sms;
    warps = warps < 1 ? 1 : (warps > DCSB_SCAN93_MAXWARPS ? DCSB_SCAN93_MAXWARPS : warps);
    const int want = (groups + warps - 1) / warps;                             // one CTA per SM: all resident (the gate waits for them)
    grid = n93 > 0 ? (want < sms ? want : sms) : 0;
}

int dcsb_scan_grid(int nstreams, int n94, int concurrent)
{
    int warps, grid = 0, lanes, warps93, grid93;
    if (n94 > 0) scan_shape(n94, concurrent, warps, grid);
    scan93_shape(nstreams - n94, lanes, warps93, grid93);
    return grid + grid93;
}

// The decode kernel may only start filling the SMs once every scan CTA is resident: its warps
// wait for scan progress, and a scan CTA that could not get onto the chip behind them would
// never deliver it.  One thread polls the counter the scan CTAs bump on entry.
__global__ void dcsb_gate_kernel(const uint32_t *started, uint32_t ctas, uint32_t *errw)
{
    const long long t0 = clock64();
    uint32_t v;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(started) : "memory");
        if (v >= ctas) break;
        if (clock64() - t0 > 4000000000ll) { atomicOr(errw, DCSB_DEV_E_GATE); break; }
        __nanosleep(128);
    }
}

cudaError_t dcsb_launch_gate(DcsbScanOut scan, int ctas, cudaStream_t st)
{
    if (!scan.started || !scan.qctl || ctas <= 0) return cudaSuccess;
    dcsb_gate_kernel<<<1, 1, 0, st>>>(scan.started, (uint32_t)ctas, scan.qctl + 2);
    return cudaGetLastError();
}

template <bool RING>
static cudaError_t launch_scan_t(const uint8_t *slab, const DcsbStreamRec *streams, const uint32_t *order, int nstreams, int warps, int grid,
                                 const DcsbTables *tables, DcsbScanOut out, cudaStream_t st, uint32_t f0, uint32_t f1)
{
    const size_t smem = DCSB_SCAN_SMEM((size_t)warps, RING);
    cudaError_t e = cudaFuncSetAttribute(dcsb_scan_kernel<RING>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)DCSB_SCAN_SMEM(RING ? DCSB_SCAN_MAXWARPS : DCSB_SCAN_MAXWARPS_DIRECT, RING));
    if (e != cudaSuccess) return e;
    // keep the SM at its largest shared-memory split, so that CTAs of the other kernel can join
    // this one (the split only changes on an idle SM)
    e = cudaFuncSetAttribute(dcsb_scan_kernel<RING>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             RING ? (int)cudaSharedmemCarveoutMaxShared : (int)((smem + 1024) * 100 / (228 * 1024) + 4));
    if (e != cudaSuccess) return e;
    dcsb_scan_kernel<RING><<<grid, warps * 32, smem, st>>>(slab, streams, order, nstreams, tables, out, f0, f1);
    return cudaGetLastError();
}

cudaError_t dcsb_launch_scan(const uint8_t *slab, const DcsbStreamRec *streams, const uint32_t *order, int nstreams, int n94, int concurrent,
                             const DcsbTables *tables, DcsbScanOut out, cudaStream_t st, uint32_t f0, uint32_t f1)
{
    if (nstreams <= 0) return cudaSuccess;
    if (n94 < 0 || n94 > nstreams || (!order && n94 > 0 && n94 < nstreams)) return cudaErrorInvalidValue;
    if (n94 > 0) {
        int warps, grid;
        scan_shape(n94, concurrent, warps, grid);
        const cudaError_t e = dcsb_scan_direct(n94, concurrent) ? launch_scan_t<false>(slab, streams, order, n94, warps, grid, tables, out, st, f0, f1)
                                                               : launch_scan_t<true>(slab, streams, order, n94, warps, grid, tables, out, st, f0, f1);
        if (e != cudaSuccess) return e;
    }
    if (nstreams > n94) {
        int lanes, warps, grid;
        scan93_shape(nstreams - n94, lanes, warps, grid);
        // keep the SM at its largest shared-memory split, so that CTAs of the decode kernel can join this one
        const cudaError_t e = cudaFuncSetAttribute(dcsb_scan93_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        dcsb_scan93_kernel<<<grid, warps * 32, 0, st>>>(slab, streams, order ? order + n94 : nullptr, nstreams - n94, lanes, tables, out, f0, f1);
    }
    return cudaGetLastError();
}

static cudaError_t launch_decode93(const uint8_t *slab, const DcsbStreamRec *streams, const DcsbTile *tiles,
                                   int ntiles, const DcsbTables *tables, DcsbScanOut scan, int16_t *pcm,
                                   unsigned long long *checksums, cudaStream_t st)
{
    if (ntiles <= 0) return cudaSuccess;
    const size_t smem = (DCSB_LUT_WORDS / 2 + DCSB_WARPS_PER_CTA * DcsbWarpSmem<true>::WORDS) * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(dcsb_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // keep the SM at its largest shared-memory split, so that CTAs of the other kernel can join
    // this one (the split only changes on an idle SM)
    e = cudaFuncSetAttribute(dcsb_decode_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    const int grid = (ntiles + DCSB_WARPS_PER_CTA - 1) / DCSB_WARPS_PER_CTA;
    dcsb_decode_kernel<true><<<grid, DCSB_WARPS_PER_CTA * 32, smem, st>>>(slab, streams, tiles, ntiles, tables, scan, pcm, checksums);
    return cudaGetLastError();
}

// tiles[0..ntiles94) are 1994-family work items, tiles[ntiles94..ntiles94+ntiles93) 1993-family tiles
cudaError_t dcsb_launch_decode(const uint8_t *slab, const DcsbStreamRec *streams, const DcsbTile *tiles,
                               int ntiles94, int ntiles93, const DcsbTables *tables, DcsbScanOut scan,
                               int16_t *pcm, unsigned long long *checksums, cudaStream_t st)
{
    if (ntiles94 > 0) {
        const size_t smem = sizeof(DcsbSmem94);
        cudaError_t e = cudaFuncSetAttribute(dcsb_decode94_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // keep the SM at its largest shared-memory split, so that CTAs of the other kernel can join
        // this one (the split only changes on an idle SM)
        e = cudaFuncSetAttribute(dcsb_decode94_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        const int grid = (ntiles94 + DCSB_WARPS94 - 1) / DCSB_WARPS94;
        dcsb_decode94_kernel<<<grid, DCSB_WARPS94 * 32, smem, st>>>(slab, streams, tiles, ntiles94, tables, scan, pcm, checksums);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return launch_decode93(slab, streams, tiles + ntiles94, ntiles93, tables, scan, pcm, checksums, st);
}
