// radix_sort.cuh -- device-side interface of the radix sort for kernels that PRODUCE the keys.
//
// The sort needs all per-pass digit histograms of its input before the first onesweep pass.  Instead of
// re-reading the keys in a histogram kernel, the kernel that writes them (the face-preprocess kernels for the
// depth sort, duplicate_kernel for the tile sort) accumulates the histograms in shared memory as a by-product,
// flushes them with one global atomic per non-empty bin, and the LAST block to finish also computes the
// sort plan (exclusive scans, skipped passes, ping-pong assignment) -- the stand-alone histogram and plan
// kernels and their launches disappear from the pipeline (they remain for dmr_sort_pairs()).
#pragma once
#include "common.cuh"

namespace dmr {

#define RS_MAX_PASS 8

// control block living in the temp buffer (zeroed before every sort)
struct SortCtl {
    uint32_t ticket[RS_MAX_PASS];
    uint32_t exec[RS_MAX_PASS];
    uint32_t src[RS_MAX_PASS];   // 0 = input, 1 = output, 2 = temp
    uint32_t dst[RS_MAX_PASS];
    uint32_t producers_done;     // blocks of the producing kernel that have flushed their histogram
    uint32_t pad[3];
};

// handle passed to a producing kernel (by value)
struct SortPre {
    uint32_t* hist;   // [npass][256], zeroed
    SortCtl* ctl;     // zeroed
    uint32_t n;       // number of keys the producer emits in total (n_dev == nullptr) / capacity of the buffers
    const uint32_t* n_dev;   // when non-null: the number of keys lives in device memory (min(*n_dev, n) are sorted)
    int npass, end_bit;
};

// number of keys of a sort whose count may live on the device
__device__ __forceinline__ uint32_t rs_count(uint32_t n_cap, const uint32_t* n_dev)
{
    if (!n_dev) return n_cap;
    const uint32_t v = *n_dev;
    return v < n_cap ? v : n_cap;
}

// exclusive scans + pass skipping + buffer assignment; blockDim.x == 256, all threads of ONE block
__device__ inline void rs_plan_block(uint32_t* __restrict__ hist, SortCtl* __restrict__ ctl, uint32_t n, int npass,
                                     uint32_t* s_scan /* 256 */, uint32_t* s_skip /* RS_MAX_PASS */)
{
    const int tid = threadIdx.x;
    for (int p = 0; p < npass; p++) {
        uint32_t c = __ldcg(&hist[p * 256 + tid]);
        if (tid == 0) s_skip[p] = 0;
        __syncthreads();
        if (c == n) s_skip[p] = 1;   // every key has the same digit -> identity pass
        s_scan[tid] = c;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            uint32_t t = (tid >= d) ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += t;
            __syncthreads();
        }
        hist[p * 256 + tid] = s_scan[tid] - c;   // exclusive
        __syncthreads();
    }
    if (tid == 0) {
        int nexec = 0;
        for (int p = 0; p < npass; p++) nexec += s_skip[p] ? 0 : 1;
        if (nexec == 0) { s_skip[0] = 0; nexec = 1; }   // always move input -> output
        int k = 0;
        uint32_t cur = 0;   // where the data currently lives
        for (int p = 0; p < npass; p++) {
            if (s_skip[p]) { ctl->exec[p] = 0; continue; }
            uint32_t dst = ((nexec - 1 - k) % 2 == 0) ? 1u : 2u;
            ctl->exec[p] = 1;
            ctl->src[p] = cur;
            ctl->dst[p] = dst;
            cur = dst;
            k++;
        }
    }
}

// Add K keys of this thread to the block's shared histograms (s_hist[npass*256], zeroed and synchronised by
// the caller).  All 32 lanes of a warp must call it together.  A digit that is the same for all 32*K keys of the
// warp -- the top depth byte, the upper tile bits of neighbouring instances -- costs one shared atomic instead
// of 32*K serialised ones (same trick as rs_hist_kernel).
template <int K>
__device__ __forceinline__ void rs_pre_add(uint32_t* s_hist, const uint32_t (&key)[K], const bool (&valid)[K], const SortPre& sp)
{
    const unsigned lane = threadIdx.x & 31;
    for (int p = 0; p < sp.npass; p++) {
        const int shift = 8 * p;
        const uint32_t mask = (sp.end_bit - shift >= 8) ? 0xffu : ((1u << (sp.end_bit - shift)) - 1u);
        uint32_t d[K];
        bool same = true;
#pragma unroll
        for (int k = 0; k < K; k++) {
            d[k] = (key[k] >> shift) & mask;
            same = same && valid[k] && d[k] == d[0];
        }
        const uint32_t d0 = __shfl_sync(0xffffffffu, d[0], 0);
        if (__all_sync(0xffffffffu, same && d[0] == d0)) {
            if (lane == 0) atomicAdd(&s_hist[p * 256 + d0], 32u * K);
        } else {
#pragma unroll
            for (int k = 0; k < K; k++)
                if (valid[k]) atomicAdd(&s_hist[p * 256 + d[k]], 1u);
        }
    }
}

// End of the producing kernel: flush the block's histograms; the last block computes the plan.
// Must be reached by all threads of every block; blockDim.x == 256.
__device__ __forceinline__ void rs_pre_finish(uint32_t* s_hist, const SortPre& sp, unsigned total_blocks)
{
    __shared__ uint32_t s_scan[256];
    __shared__ uint32_t s_skip[RS_MAX_PASS];
    __shared__ bool s_last;
    if (sp.npass == 0) return;   // the sort builds its histograms itself (large inputs, see bin_faces_begin)
    const int tid = threadIdx.x;
    __syncthreads();
    for (int i = tid; i < sp.npass * 256; i += 256) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&sp.hist[i], c);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&sp.ctl->producers_done, 1u) == total_blocks - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    rs_plan_block(sp.hist, sp.ctl, rs_count(sp.n, sp.n_dev), sp.npass, s_scan, s_skip);
}

// host side (radix_sort.cu)
size_t sort_zero_bytes(size_t n, size_t key_bytes, int end_bit);   // leading bytes of `temp` that must be zero
int sort_pre_handle(void* temp, size_t n, size_t key_bytes, int end_bit, SortPre* out);                       // no memset
int sort_pre_begin(void* temp, size_t n, size_t key_bytes, int end_bit, SortPre* out, cudaStream_t stream);   // zeroes
int sort_pairs_u32_pre(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, size_t n,
                       int end_bit, void* temp, bool profile, bool have_hist, cudaStream_t stream,
                       const uint32_t* n_dev = nullptr);   // no memset; n_dev: see SortPre

}  // namespace dmr
