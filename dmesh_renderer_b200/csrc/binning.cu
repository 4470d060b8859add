// binning.cu -- prefix sum, tile-key emission, tile ranges, and the two-level binning driver.
//
//   inclusive_scan_kernel   replaces cub::DeviceScan::InclusiveSum
//                           (cuda_rasterizer/rasterizer_impl.cu:278-284,
//                            cuda_renderer/renderer_impl.cu:296-302)
//   duplicate_kernel        replaces duplicateWithKeys
//                           (rasterizer_impl.cu:44-97, renderer_impl.cu:44-99)
//   tile_ranges_kernel      replaces identifyTileRanges (rasterizer_impl.cu:102-124)
//   bin_faces / bin_instances   stage sequencing, see common.cuh
//
// All of them are HBM-bound; see DESIGN.md for the algorithmic bytes.
#include "radix_sort.cuh"

namespace dmr {

// ---------------------------------------------------------------------------
// Single-pass inclusive scan (decoupled look-back), uint32 data, 64-bit running prefix.
// state[0] = ticket counter, state[1] = grand total (saturated at 2^32 - 1), then from word 32 on one 64-bit
// tile descriptor { flag:2 | value:62 } per tile.  state must be zero on entry.
// One read and one write of the data: 8 B per element.
// The prefix is carried in 64 bits so that a total beyond 2^30 (the sort's limit) or 2^32 is REPORTED -- the host
// then refuses the call (DMR_ETOOLARGE) -- instead of spilling into the flag bits of a 32-bit descriptor and
// silently corrupting the offsets and num_rendered (the reference's CUB scan is exact up to 2^31 - 1).
// ---------------------------------------------------------------------------
#define SCAN_FLAG_AGG  (1ull << 62)
#define SCAN_FLAG_INCL (2ull << 62)
#define SCAN_VAL_MASK  ((1ull << 62) - 1ull)

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(DMR_SCAN_THREADS) inclusive_scan_kernel(
    const uint32_t* __restrict__ in, const uint32_t* __restrict__ index, uint32_t* __restrict__ out, size_t n,
    uint32_t* __restrict__ state, int32_t* __restrict__ total_mapped)
{
    griddep_wait();   // programmatic dependent launch: see dmr_launch (common.cuh)
    __shared__ uint32_t s_warp[DMR_SCAN_THREADS / 32];
    __shared__ unsigned long long s_warp64[DMR_SCAN_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(&state[0], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const size_t base = (size_t)tile * DMR_SCAN_TILE + (size_t)tid * DMR_SCAN_ITEMS;

    // blocked arrangement: 16 consecutive items per thread = 4 x 128-bit loads
    uint32_t v[DMR_SCAN_ITEMS];
    if (index) {   // gather: element i of the scanned sequence is in[index[i]]
        if (base + DMR_SCAN_ITEMS <= n) {
            const uint4* p = reinterpret_cast<const uint4*>(index + base);
#pragma unroll
            for (int q = 0; q < DMR_SCAN_ITEMS / 4; q++) {
                uint4 t = p[q];
                v[4 * q] = in[t.x]; v[4 * q + 1] = in[t.y]; v[4 * q + 2] = in[t.z]; v[4 * q + 3] = in[t.w];
            }
        } else {
#pragma unroll
            for (int i = 0; i < DMR_SCAN_ITEMS; i++) v[i] = (base + i < n) ? in[index[base + i]] : 0u;
        }
    } else if (base + DMR_SCAN_ITEMS <= n) {
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
    unsigned long long sum64 = 0;   // exact, for the tile aggregate: 4096 faces that each cover a huge tile grid can exceed 32 bits
#pragma unroll
    for (int i = 0; i < DMR_SCAN_ITEMS; i++) { sum64 += v[i]; sum += v[i]; v[i] = sum; }

    // warp inclusive scan of the thread sums
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum64 += __shfl_xor_sync(0xffffffffu, sum64, o);
    if (lane == 31) { s_warp[warp] = incl; s_warp64[warp] = sum64; }
    __syncthreads();
    uint32_t warp_off = 0;
    unsigned long long tile_total = 0;
#pragma unroll
    for (int w = 0; w < DMR_SCAN_THREADS / 32; w++) {
        if (w < warp) warp_off += s_warp[w];
        tile_total += s_warp64[w];
    }
    uint32_t thread_excl = warp_off + incl - sum;

    // publish + decoupled look-back (warp 0)
    if (warp == 0) {
        unsigned long long* desc = reinterpret_cast<unsigned long long*>(state + 32);
        if (lane == 0) st_volatile_u64(&desc[tile], (tile == 0 ? SCAN_FLAG_INCL : SCAN_FLAG_AGG) | tile_total);
        unsigned long long excl = 0;
        if (tile > 0) {
            long long t = (long long)tile - 1;
            while (true) {
                long long idx = t - lane;
                unsigned long long d = SCAN_FLAG_INCL;   // virtual predecessor of tile 0: inclusive 0
                if (idx >= 0) {
                    do { d = ld_volatile_u64(&desc[idx]); } while ((d >> 62) == 0);
                }
                unsigned incl_mask = __ballot_sync(0xffffffffu, (d >> 62) == 2u);
                int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
                unsigned long long c = (lane <= first) ? (d & SCAN_VAL_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                excl += c;
                if (incl_mask) break;
                t -= 32;
            }
            if (lane == 0) st_volatile_u64(&desc[tile], SCAN_FLAG_INCL | (excl + tile_total));
        }
        if (lane == 0) {
            s_excl = (uint32_t)excl;     // offsets are 32-bit; a total beyond 2^32 - 1 is reported below and refused
            size_t ntile = (n + DMR_SCAN_TILE - 1) / DMR_SCAN_TILE;
            if ((size_t)tile == ntile - 1) {
                const unsigned long long tot = excl + tile_total;
                const uint32_t tot32 = tot > 0xffffffffull ? 0xffffffffu : (uint32_t)tot;
                state[1] = tot32;
                if (total_mapped) *total_mapped = (int32_t)tot32;
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

int inclusive_scan_u32(const uint32_t* in, const uint32_t* index, uint32_t* out, size_t n, uint32_t* state,
                       int32_t* total_host, cudaStream_t stream)
{
    if (n == 0) {
        if (total_host) *total_host = 0;
        return 0;
    }
    size_t ntile = (n + DMR_SCAN_TILE - 1) / DMR_SCAN_TILE;
    {
    ProfScope prof(ST_SCAN, stream);
    DMR_CUDA(dmr_launch(inclusive_scan_kernel, dim3((unsigned)ntile), dim3(DMR_SCAN_THREADS), 0, stream, in, index, out, n, state, nullptr));
    DMR_LAUNCH_CHECK("inclusive_scan_kernel");
    }
    if (total_host)
        DMR_CUDA(cudaMemcpyAsync(total_host, state + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    return 0;
}

// ---------------------------------------------------------------------------
// Key emission.  The reference runs one thread per face with a serial loop
// over its tile rectangle (strided 12-byte writes, load imbalance on large
// triangles).  Here a block owns 256 consecutive faces OF THE DEPTH ORDER and
// its 256 threads walk the block's contiguous OUTPUT range, locating the
// owning face by binary search in shared memory -> every key/value store is
// coalesced.  Emission order inside a face (y, then x) and the tile id are
// those of rasterizer_impl.cu:84-96:  tile + tiles*b, value = face id within
// the view; the depth word of the reference's key is implied by the order.
// ---------------------------------------------------------------------------
struct DupFace { uint32_t x0, w, y0, tile0, fid; };

__device__ __forceinline__ uint32_t emit_one(uint32_t o, uint32_t start, int nface, const uint32_t* s_incl, const uint2* s_rect,
                                         const uint32_t* s_tile0, const uint32_t* s_fid, int tiles_x,
                                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
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
    const uint32_t key = (y0 + q) * (uint32_t)tiles_x + x0 + (k - q * w) + s_tile0[i];
    keys[o] = key;
    vals[o] = s_fid[i];
    return key;
}

template <bool HIST>   // HIST: accumulate the tile sort's histograms as a by-product (small inputs only)
__global__ void __launch_bounds__(256) duplicate_kernel(
    size_t BF, int F, int tiles_x, int tiles_per_view, const uint32_t* __restrict__ order,
    const uint32_t* __restrict__ offsets, const uint2* __restrict__ rect,
    uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, SortPre sp, uint2* __restrict__ ranges, size_t n_ranges,
    const uint32_t* __restrict__ total_dev, uint32_t total_host)
{
    griddep_wait();   // programmatic dependent launch: see dmr_launch (common.cuh)
    // tiles without instances keep the empty range (0,0): zero the table here (tile_ranges_kernel runs after the
    // sort) instead of with one more memset between the num_rendered read-back and this launch
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < n_ranges; t += (size_t)gridDim.x * 256)
        ranges[t] = make_uint2(0u, 0u);
    // The host sized keys / vals for `total_host` instances -- its copy of num_rendered, or, when phase 2 is launched
    // before num_rendered has reached the host, the capacity of a speculatively sized buffer.  If the scan produced
    // more than that, emit nothing rather than write out of bounds (the host sees the real total and runs phase 2
    // again with a buffer that fits).
    if (*total_dev > total_host) return;
    __shared__ uint32_t s_incl[256];
    __shared__ uint2 s_rect[256];
    __shared__ uint32_t s_tile0[256];   // tiles_per_view * view of the face (per-instance divisions hoisted)
    __shared__ uint32_t s_fid[256];     // face id inside its view
    __shared__ uint32_t s_hist[HIST ? 4 * 256 : 1];   // digit histograms of the emitted tile ids (tile sort, radix_sort.cuh)
    const int tid = threadIdx.x;
    if (HIST) for (int i = tid; i < 4 * 256; i += 256) s_hist[i] = 0;
    const size_t i0 = (size_t)blockIdx.x * 256;
    const size_t i = i0 + tid;          // position in the depth order
    const uint32_t start = (i0 == 0) ? 0u : offsets[i0 - 1];
    uint32_t my_incl = 0xffffffffu;
    if (i < BF) {
        my_incl = offsets[i];
        const uint32_t f = order[i];    // b*F + f, B*F < 2^31 (checked by the caller)
        s_rect[tid] = rect[f];
        const uint32_t bview = f / (uint32_t)F;
        s_tile0[tid] = (uint32_t)tiles_per_view * bview;
        s_fid[tid] = f - bview * (uint32_t)F;
    }
    s_incl[tid] = my_incl;
    __syncthreads();
    const int nface = (int)((BF - i0 < 256) ? (BF - i0) : 256);
    const uint32_t end = s_incl[nface - 1];

    // Each thread emits 4 consecutive instances per step: one binary search, then cheap "same face or
    // next face" advances; 4 keys / 4 values leave as two 16-byte stores, so a warp writes 2 x 512 B
    // contiguous.  Head/tail elements that break 16-byte alignment go out scalar.  All loops have
    // block-uniform trip counts (the histogram update is a warp-collective).
    const uint32_t first4 = (start + 3u) & ~3u;                      // first 4-aligned output index of the block
    {   // head: at most 3 elements
        const uint32_t o = start + tid;
        const bool v[1] = { o < min(first4, end) };
        uint32_t k[1] = { 0u };
        if (v[0]) k[0] = emit_one(o, start, nface, s_incl, s_rect, s_tile0, s_fid, tiles_x, keys, vals);
        if (HIST) rs_pre_add<1>(s_hist, k, v, sp);
    }
    for (uint32_t base = first4; base < end; base += 4u * 256u) {
        const uint32_t o4 = base + 4u * tid;
        uint32_t k[4] = { 0u, 0u, 0u, 0u }, v[4];
        bool ok[4] = { false, false, false, false };
        if (o4 + 4u <= end) {
            int lo = 0, hi = nface - 1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (s_incl[mid] > o4) hi = mid; else lo = mid + 1;
            }
            int j = lo;
            while (s_incl[j] <= o4) j++;                                  // skips faces with no instances
            // first instance: position inside its face's rectangle by one division; the next three advance
            // incrementally (next column, next row, or the first tile of the next face that has instances)
            uint2 r = s_rect[j];
            uint32_t x0 = r.x & 0xffffu, w = (r.x >> 16) - x0, y0 = r.y & 0xffffu;
            uint32_t kk = o4 - ((j == 0) ? start : s_incl[j - 1]);
            uint32_t row = kk / w, col = kk - row * w;
            uint32_t fend = s_incl[j], base_t = s_tile0[j] + x0, fid = s_fid[j];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                k[e] = (y0 + row) * (uint32_t)tiles_x + base_t + col;
                v[e] = fid;
                ok[e] = true;
                if (e < 3) {
                    if (o4 + e + 1 < fend) {
                        if (++col == w) { col = 0; row++; }
                    } else {
                        do { j++; } while (s_incl[j] <= o4 + e + 1);
                        r = s_rect[j];
                        x0 = r.x & 0xffffu; w = (r.x >> 16) - x0; y0 = r.y & 0xffffu;
                        row = 0; col = 0;
                        fend = s_incl[j]; base_t = s_tile0[j] + x0; fid = s_fid[j];
                    }
                }
            }
            *reinterpret_cast<uint4*>(keys + o4) = make_uint4(k[0], k[1], k[2], k[3]);
            *reinterpret_cast<uint4*>(vals + o4) = make_uint4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (o4 + e < end) { k[e] = emit_one(o4 + e, start, nface, s_incl, s_rect, s_tile0, s_fid, tiles_x, keys, vals); ok[e] = true; }
        }
        if (HIST) rs_pre_add<4>(s_hist, k, ok, sp);
    }
    if (HIST) rs_pre_finish(s_hist, sp, gridDim.x);
}

// ---------------------------------------------------------------------------
// Tile ranges: boundaries of equal tile-id runs in the sorted key list.
// ranges must be zero on entry (tiles without instances keep (0,0),
// rasterizer_impl.cu:330).
// ---------------------------------------------------------------------------
#define TR_KPT 8
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ keys, size_t L_cap,
                                                          uint2* __restrict__ ranges, const uint32_t* __restrict__ total_dev)
{
    griddep_wait();   // programmatic dependent launch: see dmr_launch (common.cuh)
    // L_cap: capacity the grid was sized for; the number of instances is read on the device.  More instances than
    // capacity: nothing was emitted (duplicate_kernel), every range stays empty.
    const uint32_t total = *total_dev;
    if (total > L_cap) return;
    const size_t L = total;
    // 8 consecutive keys per thread (2 x 16-byte loads) plus the one before them
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * TR_KPT;
    if (i0 >= L) return;
    uint32_t t[TR_KPT + 1];
    t[0] = (i0 == 0) ? 0u : keys[i0 - 1];
    if (i0 + TR_KPT <= L) {
        const uint4* p = reinterpret_cast<const uint4*>(keys + i0);
        const uint4 a = p[0], b = p[1];
        t[1] = a.x; t[2] = a.y; t[3] = a.z; t[4] = a.w; t[5] = b.x; t[6] = b.y; t[7] = b.z; t[8] = b.w;
    } else {
#pragma unroll
        for (int e = 0; e < TR_KPT; e++) t[1 + e] = (i0 + e < L) ? keys[i0 + e] : 0u;
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

// Number of key bits above the depth word.  The reference's getHigherMsb
// (rasterizer_impl.cu:25-40) is a bisection that returns the bit length of n
// (and 1 for n == 0); the closed form below gives the same value for every n.
uint32_t higher_msb(uint32_t n)
{
    uint32_t bits = 0;
    while (n) { bits++; n >>= 1; }
    return bits ? bits : 1u;
}

namespace {
template <typename T>
T* at(void* base, size_t off) { return reinterpret_cast<T*>(static_cast<unsigned char*>(base) + off); }
template <typename T>
const T* at(const void* base, size_t off) { return reinterpret_cast<const T*>(static_cast<const unsigned char*>(base) + off); }
}  // namespace

int bin_faces_begin(size_t BF, void* fb, const FaceBinLayout& L, SortPre* face_sort, cudaStream_t stream)
{
    if (BF == 0) return 0;
    // scan descriptors and the face sort's control words are adjacent: one memset
    const size_t zero = (L.fsort_temp - L.scan_state) + sort_zero_bytes(BF, 4, 32);
    DMR_CUDA(cudaMemsetAsync(at<unsigned char>(fb, L.scan_state), 0, zero, stream));
    // depth keys are non-negative floats: integer order == float order.  All 32 bits take part (the
    // constant top byte of real scenes is skipped on the device by the sort's plan).
    int rc = sort_pre_handle(at<void>(fb, L.fsort_temp), BF, 4, 32, face_sort);
    if (BF > DMR_FUSED_FACE_HIST_MAX) face_sort->npass = 0;   // histogram + plan kernels of the sort itself
    return rc;
}

int bin_faces(size_t BF, void* fb, const FaceBinLayout& L, int32_t* num_rendered_host, cudaStream_t stream)
{
    if (BF == 0) { if (num_rendered_host) *num_rendered_host = 0; return 0; }
    int rc;
    {
        ProfScope prof(ST_FACE_SORT, stream, /*count=*/false);   // the sort counts its own kernels
        if ((rc = sort_pairs_u32_pre(at<uint32_t>(fb, L.depth_key), nullptr, at<uint32_t>(fb, L.depth_sorted),
                                     at<uint32_t>(fb, L.order), BF, 32, at<void>(fb, L.fsort_temp), false,
                                     BF <= DMR_FUSED_FACE_HIST_MAX, stream)))
            return rc;
    }
    return inclusive_scan_u32(at<uint32_t>(fb, L.tiles_touched), at<uint32_t>(fb, L.order), at<uint32_t>(fb, L.offsets),
                              BF, at<uint32_t>(fb, L.scan_state), num_rendered_host, stream);
}

int bin_instances(int B, int F, int W, int H, size_t R, const void* fb, const FaceBinLayout& L, void* binning_buffer,
                  uint2* ranges, cudaStream_t stream)
{
    const int tx = (W + DMR_TILE - 1) / DMR_TILE, ty = (H + DMR_TILE - 1) / DMR_TILE;
    const size_t tiles = (size_t)B * tx * ty;
    if (R == 0) { DMR_CUDA(cudaMemsetAsync(ranges, 0, sizeof(uint2) * tiles, stream)); return 0; }
    const size_t BF = (size_t)B * F;
    BinningLayout BL = BinningLayout::make(R);
    uint32_t* ku = at<uint32_t>(binning_buffer, BL.keys_unsorted);
    uint32_t* vu = at<uint32_t>(binning_buffer, BL.vals_unsorted);
    uint32_t* ks = at<uint32_t>(binning_buffer, BL.keys_sorted);
    uint32_t* vs = at<uint32_t>(binning_buffer, BL.vals_sorted);
    // the bits above the depth word of the reference's key: rasterizer_impl.cu:316-324
    const int tile_bits = (int)higher_msb((uint32_t)tiles);
    int rc;
    SortPre sp;
    const bool fused = R <= DMR_FUSED_TILE_HIST_MAX;
    const uint32_t* total_dev = at<uint32_t>(fb, L.scan_state) + 1;   // the scan's grand total, on the device
    if ((rc = sort_pre_begin(at<void>(binning_buffer, BL.sort_temp), R, 4, tile_bits, &sp, stream))) return rc;
    sp.n_dev = total_dev;
    {
        ProfScope prof(ST_DUPLICATE, stream);
        const unsigned nblk = (unsigned)((BF + 255) / 256);
        if (fused)
            DMR_CUDA(dmr_launch(duplicate_kernel<true>, dim3(nblk), dim3(256), 0, stream, BF, F, tx, tx * ty, at<uint32_t>(fb, L.order),
                                                            at<uint32_t>(fb, L.offsets), at<uint2>(fb, L.rect), ku, vu, sp,
                                                            ranges, tiles, total_dev, (uint32_t)R));
        else
            DMR_CUDA(dmr_launch(duplicate_kernel<false>, dim3(nblk), dim3(256), 0, stream, BF, F, tx, tx * ty, at<uint32_t>(fb, L.order),
                                                             at<uint32_t>(fb, L.offsets), at<uint2>(fb, L.rect), ku, vu, sp,
                                                             ranges, tiles, total_dev, (uint32_t)R));
        DMR_LAUNCH_CHECK("duplicate_kernel");
    }
    if ((rc = sort_pairs_u32_pre(ku, vu, ks, vs, R, tile_bits, at<void>(binning_buffer, BL.sort_temp), true, fused, stream,
                                 total_dev)))
        return rc;
    {
        ProfScope prof(ST_RANGES, stream);
        const size_t nthreads = (R + TR_KPT - 1) / TR_KPT;
        DMR_CUDA(dmr_launch(tile_ranges_kernel, dim3((unsigned)((nthreads + 255) / 256)), dim3(256), 0, stream, ks, R, ranges, total_dev));
        DMR_LAUNCH_CHECK("tile_ranges_kernel");
    }
    return 0;
}

}  // namespace dmr
