// Device-side plumbing for the vkvg_b200 pipeline: error checking, growable device buffers, block/warp
// prefix scans and a stable LSD radix sort.  All hand-written (no CUB/Thrust) so every launch in the
// timed region is one of ours.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define VKB_CUDA_OK(expr)                                                                                    \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            fprintf(stderr, "vkvg_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e__), __FILE__,      \
                    __LINE__, cudaGetErrorString(e__));                                                      \
            vkb_note_cuda_error(e__);                                                                        \
        }                                                                                                    \
    } while (0)

void vkb_note_cuda_error(cudaError_t e);  // pipeline.cu: makes the device whose entry point is running on this thread sticky-failed
extern unsigned long long g_vkb_launches; // number of kernels launched by this library (bench: gpu_launches)
#define VKB_LAUNCHED() (++g_vkb_launches)

static inline uint32_t vkb_div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// A device allocation that only ever grows (by doubling) and is reused flush after flush, so the steady
// state does no cudaMalloc.  Contents are NOT preserved across a growth unless keep=true.
extern unsigned long long g_vkb_alloc_generation;  // bumped whenever a DevBuf changes address: cached CUDA graphs bake pointers in
struct DevBuf {
    void  *p   = nullptr;
    size_t cap = 0;
    void   ensure(size_t bytes, cudaStream_t s, bool keep = false) {
        if (bytes <= cap) return;
        g_vkb_alloc_generation++;
        size_t ncap = cap ? cap : 4096;
        while (ncap < bytes) ncap *= 2;
        void *np = nullptr;
        if (cudaMalloc(&np, ncap) != cudaSuccess) {  // out of memory: keep what there is (the device turns sticky-failed and launches nothing more)
            cudaGetLastError();
            fprintf(stderr, "vkvg_b200: cudaMalloc of %zu bytes failed\n", ncap);
            vkb_note_cuda_error(cudaErrorMemoryAllocation);
            return;
        }
        if (p) {
            if (keep) VKB_CUDA_OK(cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, s));
            VKB_CUDA_OK(cudaStreamSynchronize(s));
            VKB_CUDA_OK(cudaFree(p));
        }
        p   = np;
        cap = ncap;
    }
    void release() {
        if (p) g_vkb_alloc_generation++;
        if (p) cudaFree(p);
        p   = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return (T *)p; }
};

// ----------------------------------------------------------------------------------------------------
// Counts that live on the device.
// How many points a batch flattens to, how many stroke vertices, path-tiles, tile edges ... is only known after the
// kernel that produces them.  Reading each one back would stall the stream eight times per flush, so the host never
// does: it sizes buffers and grids from CAPACITIES (what earlier flushes needed, or a guess), kernels read the real
// count from this block, and a one-thread commit kernel after each producing scan checks it against the capacity.
// If a count does not fit, `overflow` is set and every later kernel of the flush returns at once — nothing is written
// past a buffer and the surface is left untouched; the host sees the flag when it next synchronises, grows the
// capacities from `need` and replays the batch (which is still resident).  Steady state: zero host round trips.
// ----------------------------------------------------------------------------------------------------
enum {
    VKC_POINTS = 0,  // flattened points
    VKC_FILL,        // fill / clip work items: the points of the filled sub-paths (one polygon edge each)
    VKC_SITEMS,      // stroke work items (points of stroked sub-paths)
    VKC_VERTS,       // stroke vertices
    VKC_INDS,        // stroke indices
    VKC_TRIS,        // stroke triangles
    VKC_EDGES,       // all device-space edges
    VKC_PT,          // path-tiles (draw x tile of its rectangle)
    VKC_ROWS,        // path-tile rows
    VKC_NE,          // non-empty path-tiles
    VKC_TE,          // edges copied into tile lists
    VKC_FEDGES,      // fill / clip polygon edges: the items, plus the pieces NON_ZERO draws are split into at their self-intersections
    VKC_N
};
struct vkb_counts {
    uint32_t n[16];     // committed counts (0 for one that overflowed)
    uint32_t cap[16];   // capacities this flush was launched with
    uint32_t need[16];  // raw totals, valid for every count up to and including the first one that overflowed
    uint32_t overflow;  // bit i: count i did not fit
    uint32_t pad[3];
};
__device__ __forceinline__ void vkc_commit(vkb_counts *C, int idx, uint32_t raw) {
    C->need[idx] = raw;
    if (raw > C->cap[idx]) { C->overflow |= 1u << idx; C->n[idx] = 0; }
    else C->n[idx] = raw;
}

// ----------------------------------------------------------------------------------------------------
// warp / block scans
// ----------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T warp_incl_scan(T v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += n;
    }
    return v;
}
// exclusive scan across a block of BLOCK threads; `total` receives the block sum (all threads)
template <class T, int BLOCK> __device__ __forceinline__ T block_excl_scan(T v, T &total) {
    __shared__ T   warp_sums[BLOCK / 32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T              incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < BLOCK / 32 ? warp_sums[lane] : T(0);
        T s = warp_incl_scan(w);
        if (lane < BLOCK / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    T base = warp ? warp_sums[warp - 1] : T(0);
    total  = warp_sums[BLOCK / 32 - 1];
    __syncthreads();
    return base + incl - v;
}

// ----------------------------------------------------------------------------------------------------
// device-wide exclusive scan: reduce per chunk -> scan chunk sums (one block) -> scan chunks.
//   out[i] = sum_{j<i} in[j];  *total (device) = sum of all.  in may alias out.
// ----------------------------------------------------------------------------------------------------
#define VKB_SCAN_BLOCK 256
#define VKB_SCAN_ITEMS 8
#define VKB_SCAN_CHUNK (VKB_SCAN_BLOCK * VKB_SCAN_ITEMS)

// n = n_add (+ C->n[idx] when C is given: a count that lives on the device, see vkb_counts)
// The block that finishes LAST (a ticket from *ticket, which it hands back as zero: the next scan - or the next replay of a captured graph -
// starts from a clean counter) goes on to scan the chunk sums, in chunk order, so the middle launch of the classic three is gone.
// Cw: the counts an overflow of which turns the scan into a no-op, and where the total is committed (commit_idx >= 0).
template <class TI, class T>
__global__ void __launch_bounds__(VKB_SCAN_BLOCK) scan_reduce_k(const TI *in, T *sums, uint64_t n, const vkb_counts *C, int idx, uint32_t *ticket, T *total_out, vkb_counts *Cw,
                                                                 int commit_idx) {
    if (Cw && Cw->overflow) return;   // (uniform over the grid: nothing sets the flag while this kernel runs but its own last block, at the very end)
    if (C) n += C->n[idx];
    uint64_t base = (uint64_t)blockIdx.x * VKB_SCAN_CHUNK;
    T        acc  = 0;
#pragma unroll
    for (int k = 0; k < VKB_SCAN_ITEMS; k++) {
        uint64_t i = base + (uint64_t)k * VKB_SCAN_BLOCK + threadIdx.x;
        if (i < n) acc += (T)in[i];
    }
    T total;
    block_excl_scan<T, VKB_SCAN_BLOCK>(acc, total);
    __shared__ bool is_last;
    __shared__ T    carry;
    if (threadIdx.x == 0) {
        sums[blockIdx.x] = total;
        __threadfence();   // the sum is visible before the ticket is
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        carry   = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const uint32_t  m = gridDim.x;
    volatile T     *vs = sums;   // (written by other blocks: not through this SM's L1)
    for (uint32_t b0 = 0; b0 < m; b0 += VKB_SCAN_BLOCK) {
        uint32_t i = b0 + threadIdx.x;
        T        v = i < m ? vs[i] : T(0), tot;
        T        e = block_excl_scan<T, VKB_SCAN_BLOCK>(v, tot);
        if (i < m) vs[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *ticket = 0;
        if (total_out) *total_out = carry;
        if (commit_idx >= 0) vkc_commit(Cw, commit_idx, (uint32_t)carry);  // the total is itself a device-side count
    }
}
// short inputs: one block does the whole scan in one launch.  Up to ONE pass of the block (8192 items): the 32 k path-tile and 16 k histogram
// scans of a tiger frame are faster as three launches of many blocks than as four sequential passes of one (measured, threshold 64 k / 16 k /
// 8 k: C1 0.333 / 0.318 / 0.305 ms per frame, C5b 49.2 / 45.4 / 45.4 ms)
#define VKB_SCAN_SMALL (1024 * 8)
template <class TI, class T> __global__ void __launch_bounds__(1024) scan_small_k(const TI *in, T *out, uint64_t n, T *total_out, vkb_counts *C, int idx, int commit_idx) {
    if (C) { if (C->overflow) return; if (idx >= 0) n += C->n[idx]; }
    __shared__ T carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024 * 8) {
        const uint64_t b0 = base + (uint64_t)threadIdx.x * 8;
        T v[8], acc = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k] = b0 + k < n ? (T)in[b0 + k] : T(0);
            acc += v[k];
        }
        T tot;
        T e = block_excl_scan<T, 1024>(acc, tot) + carry_s;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (b0 + k < n) out[b0 + k] = e;
            e += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = carry_s;
        if (commit_idx >= 0) vkc_commit(C, commit_idx, (uint32_t)carry_s);
    }
}
// two independent short scans (each at most VKB_SCAN_SMALL items, lengths known to the host) in one launch: block 0 the first, block 1 the second
template <class T> __global__ void __launch_bounds__(1024) scan_small2_k(T *a, uint32_t na, T *total_a, T *b, uint32_t nb, T *total_b) {
    T             *p = blockIdx.x ? b : a, *total_out = blockIdx.x ? total_b : total_a;
    const uint32_t n = blockIdx.x ? nb : na, b0 = threadIdx.x * 8;
    T              v[8], acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = b0 + k < n ? p[b0 + k] : T(0);
        acc += v[k];
    }
    T tot;
    T e = block_excl_scan<T, 1024>(acc, tot);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (b0 + k < n) p[b0 + k] = e;
        e += v[k];
    }
    if (threadIdx.x == 0) *total_out = tot;
}
template <class TI, class T> __global__ void __launch_bounds__(VKB_SCAN_BLOCK) scan_apply_k(const TI *in, T *out, const T *sums, uint64_t n, const vkb_counts *C, int idx) {
    if (C) { if (C->overflow) return; n += C->n[idx]; }   // (an overflow the scan's own commit raised included: the offsets would be of no use)
    // each thread owns VKB_SCAN_ITEMS consecutive items so the order of summation is the input order
    uint64_t base = (uint64_t)blockIdx.x * VKB_SCAN_CHUNK + (uint64_t)threadIdx.x * VKB_SCAN_ITEMS;
    T        v[VKB_SCAN_ITEMS], acc = 0;
#pragma unroll
    for (int k = 0; k < VKB_SCAN_ITEMS; k++) {
        uint64_t i = base + k;
        v[k]       = i < n ? (T)in[i] : T(0);
        acc += v[k];
    }
    T total;
    T e = block_excl_scan<T, VKB_SCAN_BLOCK>(acc, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < VKB_SCAN_ITEMS; k++) {
        uint64_t i = base + k;
        if (i < n) out[i] = e;
        e += v[k];
    }
}
// scan_reduce_k sums items strided; scan_apply_k sums them blocked — both give the same chunk total for
// integers; for floating point the chunk totals come from scan_reduce_k only, and the within-chunk order
// from scan_apply_k, which is deterministic run to run.
struct ScanScratch {
    DevBuf sums;
    DevBuf ticket;  // one zeroed word: which block of scan_reduce_k finishes last
};
// the zeroed words the "last block" kernels take their tickets from (word 0: scan_reduce_k, word 1: sort_hist_k); each hands its word back as zero
static inline uint32_t *vkb_scan_ticket(ScanScratch &sc, cudaStream_t s) {
    if (!sc.ticket.p) {
        sc.ticket.ensure(256, s);
        if (sc.ticket.p) VKB_CUDA_OK(cudaMemsetAsync(sc.ticket.p, 0, 256, s));
    }
    return sc.ticket.as<uint32_t>();
}
// host-known length n; with C the length is n + C->n[idx] and `cap` bounds it (grid and scratch are sized for cap)
// commit_idx >= 0: the total is committed as count commit_idx of C (checked against its capacity)
template <class TI, class T>
static inline void vkb_exclusive_scan(const TI *in, T *out, uint64_t n, T *total_dev, ScanScratch &sc, cudaStream_t s, vkb_counts *C = nullptr,
                                      int idx = -1, uint64_t cap = 0, int commit_idx = -1) {
    const uint64_t bound = (C && idx >= 0) ? cap : n;
    if (bound == 0) {
        if (total_dev) VKB_CUDA_OK(cudaMemsetAsync(total_dev, 0, sizeof(T), s));
        return;  // (a committed count keeps the zero the reset gave it)
    }
    if (bound <= VKB_SCAN_SMALL) {
        scan_small_k<TI, T><<<1, 1024, 0, s>>>(in, out, n, total_dev, C, idx, commit_idx);
        VKB_LAUNCHED();
        return;
    }
    const vkb_counts *Cn = (C && idx >= 0) ? C : nullptr;
    uint32_t chunks = vkb_div_up(bound, VKB_SCAN_CHUNK);
    sc.sums.ensure((size_t)chunks * sizeof(T), s);
    uint32_t *ticket = vkb_scan_ticket(sc, s);
    if (!sc.sums.p || !ticket) return;  // out of memory: the device is sticky-failed already
    scan_reduce_k<TI, T><<<chunks, VKB_SCAN_BLOCK, 0, s>>>(in, sc.sums.as<T>(), n, Cn, idx, ticket, total_dev, C, commit_idx);
    VKB_LAUNCHED();
    scan_apply_k<TI, T><<<chunks, VKB_SCAN_BLOCK, 0, s>>>(in, out, sc.sums.as<T>(), n, Cn, idx);
    VKB_LAUNCHED();
}

// ----------------------------------------------------------------------------------------------------
// stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass.
// Each block owns a contiguous chunk of 4096 pairs; each warp a contiguous 512 of those, walked in 16
// rounds of 32, so "earlier in memory" == "earlier (warp, round, lane)" and ranks are stable.
// ----------------------------------------------------------------------------------------------------
// ROUNDS = 16 for large inputs (few, long blocks); 2 for short ones, where the sequential rounds of a block are pure latency
#define VKB_SORT_BLOCK 256

template <int ROUNDS> __device__ __forceinline__ void sort_warp_hist(const uint32_t *keys, uint64_t n, uint64_t warp_base, int shift, uint32_t *wc) {
    const unsigned lane = threadIdx.x & 31;
    for (int r = 0; r < ROUNDS; r++) {
        uint64_t i     = warp_base + (uint64_t)r * 32 + lane;
        uint32_t d     = i < n ? ((keys[i] >> shift) & 0xFF) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d < 256 && (peers & ((1u << lane) - 1)) == 0) wc[d] += __popc(peers);
        __syncwarp();
    }
}
// ticket != null (short inputs: at most 32 blocks): the block that finishes last also turns the histogram into its exclusive scan - thread d
// walks the row of digit d - so that the pass is two launches, not three
template <int ROUNDS>
static __global__ void __launch_bounds__(VKB_SORT_BLOCK) sort_hist_k(const uint32_t *keys, uint64_t n, int shift, uint32_t *hist, uint32_t nblocks,
                                                                     const vkb_counts *C, int idx, uint32_t *ticket) {
    constexpr int VKB_SORT_ROUNDS = ROUNDS, VKB_SORT_CHUNK = VKB_SORT_BLOCK * ROUNDS;
    if (C) { if (C->overflow) return; n = C->n[idx]; }
    __shared__ uint32_t wc[VKB_SORT_BLOCK / 32][256];
    for (int i = threadIdx.x; i < (VKB_SORT_BLOCK / 32) * 256; i += VKB_SORT_BLOCK) (&wc[0][0])[i] = 0;
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5;
    sort_warp_hist<ROUNDS>(keys, n, (uint64_t)blockIdx.x * VKB_SORT_CHUNK + (uint64_t)warp * 32 * VKB_SORT_ROUNDS, shift, wc[warp]);
    __syncthreads();
    uint32_t sum = 0;
    for (int w = 0; w < VKB_SORT_BLOCK / 32; w++) sum += wc[w][threadIdx.x];
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = sum;  // digit-major so one scan yields global offsets
    if (!ticket) return;
    __shared__ bool is_last;
    __threadfence();   // this thread's count is visible before the block's ticket is
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    volatile uint32_t *row = hist + (uint64_t)threadIdx.x * nblocks;   // (written by other blocks: not through this SM's L1)
    uint32_t           rsum = 0, tot;
    for (uint32_t b = 0; b < nblocks; b++) rsum += row[b];
    uint32_t base = block_excl_scan<uint32_t, VKB_SORT_BLOCK>(rsum, tot);
    for (uint32_t b = 0; b < nblocks; b++) {
        const uint32_t c = row[b];
        row[b] = base;
        base += c;
    }
    if (threadIdx.x == 0) *ticket = 0;
}
template <int ROUNDS>
static __global__ void __launch_bounds__(VKB_SORT_BLOCK) sort_scatter_k(const uint32_t *keys, const uint32_t *vals, uint32_t *okeys, uint32_t *ovals,
                                                                uint64_t n, int shift, const uint32_t *hist_scan, uint32_t nblocks, const vkb_counts *C, int idx) {
    constexpr int VKB_SORT_ROUNDS = ROUNDS, VKB_SORT_CHUNK = VKB_SORT_BLOCK * ROUNDS;
    if (C) { if (C->overflow) return; n = C->n[idx]; }
    __shared__ uint32_t wc[VKB_SORT_BLOCK / 32][256];
    for (int i = threadIdx.x; i < (VKB_SORT_BLOCK / 32) * 256; i += VKB_SORT_BLOCK) (&wc[0][0])[i] = 0;
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t warp_base = (uint64_t)blockIdx.x * VKB_SORT_CHUNK + (uint64_t)warp * 32 * VKB_SORT_ROUNDS;
    sort_warp_hist<ROUNDS>(keys, n, warp_base, shift, wc[warp]);
    __syncthreads();
    {   // thread d: turn per-warp counts into global start offsets for (digit d, this block, warp w)
        uint32_t run = hist_scan[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
        for (int w = 0; w < VKB_SORT_BLOCK / 32; w++) {
            uint32_t c          = wc[w][threadIdx.x];
            wc[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int r = 0; r < VKB_SORT_ROUNDS; r++) {
        uint64_t i     = warp_base + (uint64_t)r * 32 + lane;
        bool     ok    = i < n;
        uint32_t k     = ok ? keys[i] : 0, v = ok ? vals[i] : 0;
        uint32_t d     = ok ? ((k >> shift) & 0xFF) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned rank  = __popc(peers & ((1u << lane) - 1));
        uint32_t base  = ok ? wc[warp][d] : 0;
        __syncwarp();
        if (ok) {
            okeys[base + rank] = k;
            ovals[base + rank] = v;
            if (rank == 0) wc[warp][d] = base + __popc(peers);
        }
        __syncwarp();
    }
}
struct SortScratch {
    DevBuf      hist, k2, v2;
    ScanScratch scan;
};
// sorts in place (result ends in keys/vals); `bits` = number of significant key bits.  With C the number of pairs is
// C->n[idx] (<= n, which then is the capacity the launch is sized for).
static inline void vkb_radix_sort(uint32_t *keys, uint32_t *vals, uint64_t n, int bits, SortScratch &sc, cudaStream_t s, vkb_counts *C = nullptr,
                                  int idx = 0) {
    if (n < 2) return;
    const bool small   = n <= 65536;
    uint32_t   nblocks = vkb_div_up(n, VKB_SORT_BLOCK * (small ? 2 : 16));
    sc.hist.ensure((size_t)256 * nblocks * 4, s);
    sc.k2.ensure(n * 4, s);
    sc.v2.ensure(n * 4, s);
    uint32_t *ka = keys, *va = vals, *kb = sc.k2.as<uint32_t>(), *vb = sc.v2.as<uint32_t>();
    int       passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    uint32_t *ticket = (small && nblocks <= 32) ? vkb_scan_ticket(sc.scan, s) : nullptr;   // (then the histogram kernel scans its own output)
    if (ticket) ticket += 1;
    for (int p = 0; p < passes; p++) {
        if (small) sort_hist_k<2><<<nblocks, VKB_SORT_BLOCK, 0, s>>>(ka, n, p * 8, sc.hist.as<uint32_t>(), nblocks, C, idx, ticket);
        else sort_hist_k<16><<<nblocks, VKB_SORT_BLOCK, 0, s>>>(ka, n, p * 8, sc.hist.as<uint32_t>(), nblocks, C, idx, nullptr);
        VKB_LAUNCHED();
        if (!ticket) vkb_exclusive_scan<uint32_t, uint32_t>(sc.hist.as<uint32_t>(), sc.hist.as<uint32_t>(), (uint64_t)256 * nblocks, nullptr, sc.scan, s);
        if (small) sort_scatter_k<2><<<nblocks, VKB_SORT_BLOCK, 0, s>>>(ka, va, kb, vb, n, p * 8, sc.hist.as<uint32_t>(), nblocks, C, idx);
        else sort_scatter_k<16><<<nblocks, VKB_SORT_BLOCK, 0, s>>>(ka, va, kb, vb, n, p * 8, sc.hist.as<uint32_t>(), nblocks, C, idx);
        VKB_LAUNCHED();
        uint32_t *t;
        t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    if (ka != keys) {
        VKB_CUDA_OK(cudaMemcpyAsync(keys, ka, n * 4, cudaMemcpyDeviceToDevice, s));
        VKB_CUDA_OK(cudaMemcpyAsync(vals, va, n * 4, cudaMemcpyDeviceToDevice, s));
    }
}
