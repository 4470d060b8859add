// binning.cu -- prefix sum, (tile|depth) key emission, tile ranges.
//
//   inclusive_scan_kernel   replaces cub::DeviceScan::InclusiveSum
//                           (cuda_rasterizer/rasterizer_impl.cu:278-284,
//                            cuda_renderer/renderer_impl.cu:296-302)
//   duplicate_kernel        replaces duplicateWithKeys
//                           (rasterizer_impl.cu:44-97, renderer_impl.cu:44-99)
//   tile_ranges_kernel      replaces identifyTileRanges (rasterizer_impl.cu:102-124)
//
// All three are HBM-bound; see DESIGN.md for the algorithmic bytes.
#include "common.cuh"

namespace dmr {

// ---------------------------------------------------------------------------
// Single-pass inclusive scan (decoupled look-back), uint32.
// state[0] = ticket counter, state[1] = grand total, state[32 + t] = tile
// descriptor { flag:2 | value:30 }.  state must be zero on entry.
// One read and one write of the data: 8 B per element.
// ---------------------------------------------------------------------------
#define SCAN_FLAG_AGG  (1u << 30)
#define SCAN_FLAG_INCL (2u << 30)
#define SCAN_VAL_MASK  ((1u << 30) - 1u)

__global__ void __launch_bounds__(DMR_SCAN_THREADS) inclusive_scan_kernel(
    const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ state,
    int32_t* __restrict__ total_mapped)
{
    __shared__ uint32_t s_warp[DMR_SCAN_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(&state[0], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t)tile * DMR_SCAN_TILE + (size_t)tid * DMR_SCAN_ITEMS;

    // blocked arrangement: 16 consecutive items per thread = 4 x 128-bit loads
    uint32_t v[DMR_SCAN_ITEMS];
    if (base + DMR_SCAN_ITEMS <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
        for (int q = 0; q < DMR_SCAN_ITEMS / 4; q++) {
            uint4 t = p[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < DMR_SCAN_ITEMS; i++) v[i] = (base + i < n) ? in[base + i] : 0u;
    }
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < DMR_SCAN_ITEMS; i++) { sum += v[i]; v[i] = sum; }

    // warp inclusive scan of the thread sums
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < DMR_SCAN_THREADS / 32; w++) {
        uint32_t t = s_warp[w];
        if (w < warp) warp_off += t;
        tile_total += t;
    }
    uint32_t thread_excl = warp_off + incl - sum;

    // publish + decoupled look-back (warp 0)
    if (warp == 0) {
        uint32_t* desc = state + 32;
        if (lane == 0) st_volatile_u32(&desc[tile], (tile == 0 ? SCAN_FLAG_INCL : SCAN_FLAG_AGG) | tile_total);
        uint32_t excl = 0;
        if (tile > 0) {
            long long t = (long long)tile - 1;
            while (true) {
                long long idx = t - lane;
                uint32_t d = SCAN_FLAG_INCL;   // virtual predecessor of tile 0: inclusive 0
                if (idx >= 0) {
                    do { d = ld_volatile_u32(&desc[idx]); } while ((d >> 30) == 0);
                }
                unsigned incl_mask = __ballot_sync(0xffffffffu, (d >> 30) == 2u);
                int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
                uint32_t c = (lane <= first) ? (d & SCAN_VAL_MASK) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                if (incl_mask) break;
                t -= 32;
            }
            if (lane == 0) st_volatile_u32(&desc[tile], SCAN_FLAG_INCL | (excl + tile_total));
        }
        if (lane == 0) {
            s_excl = excl;
            size_t ntile = (n + DMR_SCAN_TILE - 1) / DMR_SCAN_TILE;
            if ((size_t)tile == ntile - 1) {
                state[1] = excl + tile_total;
                if (total_mapped) *total_mapped = (int32_t)(excl + tile_total);
            }
        }
    }
    __syncthreads();
    const uint32_t add = s_excl + thread_excl;

    if (base + DMR_SCAN_ITEMS <= n) {
        uint4* p = reinterpret_cast<uint4*>(out + base);
#pragma unroll
        for (int q = 0; q < DMR_SCAN_ITEMS / 4; q++)
            p[q] = make_uint4(v[4 * q] + add, v[4 * q + 1] + add, v[4 * q + 2] + add, v[4 * q + 3] + add);
    } else {
#pragma unroll
        for (int i = 0; i < DMR_SCAN_ITEMS; i++)
            if (base + i < n) out[base + i] = v[i] + add;
    }
}

int inclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* state, int32_t* total_host,
                       cudaStream_t stream)
{
    if (n == 0) {
        if (total_host) *total_host = 0;
        return 0;
    }
    size_t ntile = (n + DMR_SCAN_TILE - 1) / DMR_SCAN_TILE;
    {
    ProfScope prof(ST_SCAN, stream);
    inclusive_scan_kernel<<<(unsigned)ntile, DMR_SCAN_THREADS, 0, stream>>>(in, out, n, state, nullptr);
    DMR_LAUNCH_CHECK("inclusive_scan_kernel");
    }
    if (total_host)
        DMR_CUDA(cudaMemcpyAsync(total_host, state + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    return 0;
}

// ---------------------------------------------------------------------------
// Key emission.  The reference runs one thread per face with a serial loop
// over its tile rectangle (strided 12-byte writes, load imbalance on large
// triangles).  Here a block owns 256 consecutive faces and its 256 threads
// walk the block's contiguous OUTPUT range, locating the owning face by
// binary search in shared memory -> every key/value store is coalesced.
// Emission order (face-major, then y, then x) and key layout are those of
// rasterizer_impl.cu:84-96:  key = (tile + tiles*b) << 32 | depth_bits,
// value = face id within the view.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void emit_one(uint32_t o, uint32_t start, int nface, const uint32_t* s_incl, const uint2* s_rect,
                                         const uint32_t* s_depth, const uint32_t* s_tile0, const uint32_t* s_fid,
                                         int tiles_x, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    int lo = 0, hi = nface - 1;   // smallest i with s_incl[i] > o
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s_incl[mid] > o) hi = mid; else lo = mid + 1;
    }
    const int i = lo;
    const uint32_t excl = (i == 0) ? start : s_incl[i - 1];
    const uint32_t k = o - excl;
    const uint2 r = s_rect[i];
    const uint32_t x0 = r.x & 0xffffu, w = (r.x >> 16) - x0, y0 = r.y & 0xffffu;
    const uint32_t q = k / w;
    keys[o] = ((uint64_t)((y0 + q) * (uint32_t)tiles_x + x0 + (k - q * w) + s_tile0[i]) << 32) | (uint64_t)s_depth[i];
    vals[o] = s_fid[i];
}

__global__ void __launch_bounds__(256) duplicate_kernel(
    size_t BF, int F, int tiles_x, int tiles_per_view,
    const uint32_t* __restrict__ offsets, const uint2* __restrict__ rect, const uint32_t* __restrict__ depth_key,
    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    __shared__ uint32_t s_incl[256];
    __shared__ uint2 s_rect[256];
    __shared__ uint32_t s_depth[256];
    __shared__ uint32_t s_tile0[256];   // tiles_per_view * view of the face (per-instance 64-bit divisions hoisted)
    __shared__ uint32_t s_fid[256];     // face id inside its view
    const int tid = threadIdx.x;
    const size_t f0 = (size_t)blockIdx.x * 256;
    const size_t f = f0 + tid;
    const uint32_t start = (f0 == 0) ? 0u : offsets[f0 - 1];
    uint32_t my_incl = 0xffffffffu;
    if (f < BF) {
        my_incl = offsets[f];
        s_rect[tid] = rect[f];
        s_depth[tid] = depth_key[f];
        const uint32_t bview = (uint32_t)((uint32_t)f / (uint32_t)F);   // B*F < 2^31 (checked by the caller)
        s_tile0[tid] = (uint32_t)tiles_per_view * bview;
        s_fid[tid] = (uint32_t)f - bview * (uint32_t)F;
    }
    s_incl[tid] = my_incl;
    __syncthreads();
    const int nface = (int)((BF - f0 < 256) ? (BF - f0) : 256);
    const uint32_t end = s_incl[nface - 1];

    // Each thread emits 4 consecutive instances per step: one binary search, then cheap "same face or
    // next face" advances; the 4 keys / 4 values leave as two 16-byte + one 16-byte store, so a warp
    // writes 1 KB + 512 B contiguous.  Head/tail elements that break 16-byte alignment go out scalar.
    const uint32_t first4 = (start + 3u) & ~3u;                      // first 4-aligned output index of the block
    for (uint32_t o = start + tid; o < min(first4, end); o += 256) emit_one(o, start, nface, s_incl, s_rect, s_depth, s_tile0, s_fid, tiles_x, keys, vals);
    for (uint32_t o4 = first4 + 4u * tid; o4 < end; o4 += 4u * 256u) {
        if (o4 + 4u > end) {
            for (uint32_t o = o4; o < end; o++) emit_one(o, start, nface, s_incl, s_rect, s_depth, s_tile0, s_fid, tiles_x, keys, vals);
            break;
        }
        int lo = 0, hi = nface - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (s_incl[mid] > o4) hi = mid; else lo = mid + 1;
        }
        int i = lo;
        uint64_t k[4];
        uint32_t v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t o = o4 + e;
            while (s_incl[i] <= o) i++;                               // skips faces with no instances
            const uint32_t excl = (i == 0) ? start : s_incl[i - 1];
            const uint32_t kk = o - excl;
            const uint2 r = s_rect[i];
            const uint32_t x0 = r.x & 0xffffu, w = (r.x >> 16) - x0, y0 = r.y & 0xffffu;
            const uint32_t q = kk / w;
            k[e] = ((uint64_t)((y0 + q) * (uint32_t)tiles_x + x0 + (kk - q * w) + s_tile0[i]) << 32) | (uint64_t)s_depth[i];
            v[e] = s_fid[i];
        }
        uint4* kd = reinterpret_cast<uint4*>(keys + o4);
        kd[0] = make_uint4((uint32_t)k[0], (uint32_t)(k[0] >> 32), (uint32_t)k[1], (uint32_t)(k[1] >> 32));
        kd[1] = make_uint4((uint32_t)k[2], (uint32_t)(k[2] >> 32), (uint32_t)k[3], (uint32_t)(k[3] >> 32));
        *reinterpret_cast<uint4*>(vals + o4) = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

int duplicate_with_keys(size_t BF, int F, int tiles_x, int tiles_y, const uint32_t* offsets, const uint2* rect,
                        const uint32_t* depth_key, uint64_t* keys, uint32_t* vals, size_t R, cudaStream_t stream)
{
    if (BF == 0 || R == 0) return 0;
    unsigned nblk = (unsigned)((BF + 255) / 256);
    ProfScope prof(ST_DUPLICATE, stream);
    duplicate_kernel<<<nblk, 256, 0, stream>>>(BF, F, tiles_x, tiles_x * tiles_y, offsets, rect, depth_key, keys, vals);
    DMR_LAUNCH_CHECK("duplicate_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// Tile ranges: boundaries of equal tile-id runs in the sorted key list.
// ranges must be zero on entry (tiles without instances keep (0,0),
// rasterizer_impl.cu:330).
// ---------------------------------------------------------------------------
#define TR_KPT 8
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint64_t* __restrict__ keys, size_t L,
                                                          uint2* __restrict__ ranges)
{
    // 8 consecutive keys per thread (4 x 16-byte loads) plus the one before them
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * TR_KPT;
    if (i0 >= L) return;
    uint32_t t[TR_KPT + 1];
    t[0] = (i0 == 0) ? 0u : (uint32_t)(keys[i0 - 1] >> 32);
    if (i0 + TR_KPT <= L) {
        const uint4* p = reinterpret_cast<const uint4*>(keys + i0);
#pragma unroll
        for (int q = 0; q < TR_KPT / 2; q++) { uint4 v = p[q]; t[1 + 2 * q] = v.y; t[2 + 2 * q] = v.w; }
    } else {
#pragma unroll
        for (int e = 0; e < TR_KPT; e++) t[1 + e] = (i0 + e < L) ? (uint32_t)(keys[i0 + e] >> 32) : 0u;
    }
#pragma unroll
    for (int e = 0; e < TR_KPT; e++) {
        const size_t idx = i0 + e;
        if (idx >= L) break;
        const uint32_t cur = t[1 + e];
        if (idx == 0) ranges[cur].x = 0;
        else if (cur != t[e]) { ranges[t[e]].y = (uint32_t)idx; ranges[cur].x = (uint32_t)idx; }
        if (idx == L - 1) ranges[cur].y = (uint32_t)L;
    }
}

int identify_tile_ranges(const uint64_t* keys_sorted, size_t R, uint2* ranges, cudaStream_t stream)
{
    if (R == 0) return 0;
    ProfScope prof(ST_RANGES, stream);
    const size_t nthreads = (R + TR_KPT - 1) / TR_KPT;
    tile_ranges_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, stream>>>(keys_sorted, R, ranges);
    DMR_LAUNCH_CHECK("tile_ranges_kernel");
    return 0;
}

// Number of key bits above the depth word.  The reference's getHigherMsb
// (rasterizer_impl.cu:25-40) is a bisection that returns the bit length of n
// (and 1 for n == 0); the closed form below gives the same value for every n.
uint32_t higher_msb(uint32_t n)
{
    uint32_t bits = 0;
    while (n) { bits++; n >>= 1; }
    return bits ? bits : 1u;
}

}  // namespace dmr
