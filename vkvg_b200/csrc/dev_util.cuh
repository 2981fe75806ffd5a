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

void vkb_note_cuda_error(cudaError_t e);  // pipeline.cu: makes the owning device sticky-failed
extern unsigned long long g_vkb_launches; // number of kernels launched by this library (bench: gpu_launches)
#define VKB_LAUNCHED() (++g_vkb_launches)

static inline uint32_t vkb_div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// A device allocation that only ever grows (by doubling) and is reused flush after flush, so the steady
// state does no cudaMalloc.  Contents are NOT preserved across a growth unless keep=true.
struct DevBuf {
    void  *p   = nullptr;
    size_t cap = 0;
    void   ensure(size_t bytes, cudaStream_t s, bool keep = false) {
        if (bytes <= cap) return;
        size_t ncap = cap ? cap : 4096;
        while (ncap < bytes) ncap *= 2;
        void *np = nullptr;
        VKB_CUDA_OK(cudaMalloc(&np, ncap));
        if (p) {
            if (keep) VKB_CUDA_OK(cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, s));
            VKB_CUDA_OK(cudaStreamSynchronize(s));
            VKB_CUDA_OK(cudaFree(p));
        }
        p   = np;
        cap = ncap;
    }
    void release() {
        if (p) cudaFree(p);
        p   = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return (T *)p; }
};

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

template <class TI, class T> __global__ void __launch_bounds__(VKB_SCAN_BLOCK) scan_reduce_k(const TI *in, T *sums, uint64_t n) {
    uint64_t base = (uint64_t)blockIdx.x * VKB_SCAN_CHUNK;
    T        acc  = 0;
#pragma unroll
    for (int k = 0; k < VKB_SCAN_ITEMS; k++) {
        uint64_t i = base + (uint64_t)k * VKB_SCAN_BLOCK + threadIdx.x;
        if (i < n) acc += (T)in[i];
    }
    T total;
    block_excl_scan<T, VKB_SCAN_BLOCK>(acc, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
template <class T> __global__ void __launch_bounds__(1024) scan_sums_k(T *sums, uint32_t m, T *total_out) {
    __shared__ T carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < m; base += 1024) {
        uint32_t i = base + threadIdx.x;
        T        v = i < m ? sums[i] : T(0), tot;
        T        e = block_excl_scan<T, 1024>(v, tot);
        if (i < m) sums[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
template <class TI, class T> __global__ void __launch_bounds__(VKB_SCAN_BLOCK) scan_apply_k(const TI *in, T *out, const T *sums, uint64_t n) {
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
};
template <class TI, class T>
static inline void vkb_exclusive_scan(const TI *in, T *out, uint64_t n, T *total_dev, ScanScratch &sc, cudaStream_t s) {
    if (n == 0) {
        if (total_dev) VKB_CUDA_OK(cudaMemsetAsync(total_dev, 0, sizeof(T), s));
        return;
    }
    uint32_t chunks = vkb_div_up(n, VKB_SCAN_CHUNK);
    sc.sums.ensure((size_t)chunks * sizeof(T), s);
    scan_reduce_k<TI, T><<<chunks, VKB_SCAN_BLOCK, 0, s>>>(in, sc.sums.as<T>(), n);
    VKB_LAUNCHED();
    scan_sums_k<T><<<1, 1024, 0, s>>>(sc.sums.as<T>(), chunks, total_dev);
    VKB_LAUNCHED();
    scan_apply_k<TI, T><<<chunks, VKB_SCAN_BLOCK, 0, s>>>(in, out, sc.sums.as<T>(), n);
    VKB_LAUNCHED();
}

// ----------------------------------------------------------------------------------------------------
// stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass.
// Each block owns a contiguous chunk of 4096 pairs; each warp a contiguous 512 of those, walked in 16
// rounds of 32, so "earlier in memory" == "earlier (warp, round, lane)" and ranks are stable.
// ----------------------------------------------------------------------------------------------------
#define VKB_SORT_BLOCK 256
#define VKB_SORT_ROUNDS 16
#define VKB_SORT_CHUNK (VKB_SORT_BLOCK * VKB_SORT_ROUNDS)

__device__ __forceinline__ void sort_warp_hist(const uint32_t *keys, uint64_t n, uint64_t warp_base, int shift, uint32_t *wc) {
    const unsigned lane = threadIdx.x & 31;
    for (int r = 0; r < VKB_SORT_ROUNDS; r++) {
        uint64_t i     = warp_base + (uint64_t)r * 32 + lane;
        uint32_t d     = i < n ? ((keys[i] >> shift) & 0xFF) : 256u;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d < 256 && (peers & ((1u << lane) - 1)) == 0) wc[d] += __popc(peers);
        __syncwarp();
    }
}
static __global__ void __launch_bounds__(VKB_SORT_BLOCK) sort_hist_k(const uint32_t *keys, uint64_t n, int shift, uint32_t *hist, uint32_t nblocks) {
    __shared__ uint32_t wc[VKB_SORT_BLOCK / 32][256];
    for (int i = threadIdx.x; i < (VKB_SORT_BLOCK / 32) * 256; i += VKB_SORT_BLOCK) (&wc[0][0])[i] = 0;
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5;
    sort_warp_hist(keys, n, (uint64_t)blockIdx.x * VKB_SORT_CHUNK + (uint64_t)warp * 32 * VKB_SORT_ROUNDS, shift, wc[warp]);
    __syncthreads();
    uint32_t sum = 0;
    for (int w = 0; w < VKB_SORT_BLOCK / 32; w++) sum += wc[w][threadIdx.x];
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = sum;  // digit-major so one scan yields global offsets
}
static __global__ void __launch_bounds__(VKB_SORT_BLOCK) sort_scatter_k(const uint32_t *keys, const uint32_t *vals, uint32_t *okeys, uint32_t *ovals,
                                                                uint64_t n, int shift, const uint32_t *hist_scan, uint32_t nblocks) {
    __shared__ uint32_t wc[VKB_SORT_BLOCK / 32][256];
    for (int i = threadIdx.x; i < (VKB_SORT_BLOCK / 32) * 256; i += VKB_SORT_BLOCK) (&wc[0][0])[i] = 0;
    __syncthreads();
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t warp_base = (uint64_t)blockIdx.x * VKB_SORT_CHUNK + (uint64_t)warp * 32 * VKB_SORT_ROUNDS;
    sort_warp_hist(keys, n, warp_base, shift, wc[warp]);
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
// sorts in place (result ends in keys/vals); `bits` = number of significant key bits
static inline void vkb_radix_sort(uint32_t *keys, uint32_t *vals, uint64_t n, int bits, SortScratch &sc, cudaStream_t s) {
    if (n < 2) return;
    uint32_t nblocks = vkb_div_up(n, VKB_SORT_CHUNK);
    sc.hist.ensure((size_t)256 * nblocks * 4, s);
    sc.k2.ensure(n * 4, s);
    sc.v2.ensure(n * 4, s);
    uint32_t *ka = keys, *va = vals, *kb = sc.k2.as<uint32_t>(), *vb = sc.v2.as<uint32_t>();
    int       passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    for (int p = 0; p < passes; p++) {
        sort_hist_k<<<nblocks, VKB_SORT_BLOCK, 0, s>>>(ka, n, p * 8, sc.hist.as<uint32_t>(), nblocks);
        VKB_LAUNCHED();
        vkb_exclusive_scan<uint32_t, uint32_t>(sc.hist.as<uint32_t>(), sc.hist.as<uint32_t>(), (uint64_t)256 * nblocks, nullptr, sc.scan, s);
        sort_scatter_k<<<nblocks, VKB_SORT_BLOCK, 0, s>>>(ka, va, kb, vb, n, p * 8, sc.hist.as<uint32_t>(), nblocks);
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
