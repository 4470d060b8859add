// Dev tool: the hand-written onesweep of libdmesh_b200.so (dmr_sort_pairs_u32 / dmr_sort_pairs) against
// cub::DeviceRadixSort::SortPairs -- the library sort the reference calls (cuda_rasterizer/rasterizer_impl.cu:319-324)
// -- on the same renderer-like keys, same box, same process.  Results are compared element by element.
//
//   build (here, no GPU needed):  bash tools/build_sort_vs_cub.sh        -> tools/_bin/sort_vs_cub
//   run (GPU box):                tools/_bin/sort_vs_cub [n=32000000] [tiles=16384]
//
// Three measurements per n (CUDA events, median of 9 after 3 warm-ups):
//   (a) u32 tile-id keys + u32 values on bits [0, bit_length(tiles))      ours vs CUB   (the tile sort we run)
//   (b) the same with end_bit = 8 and 16                                   -> time of ONE pass = t(16) - t(8)
//   (c) u64 (tile << 32 | depth bits) keys + u32 values on 32 + bits       CUB only      (the sort the reference runs)
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../include/dmesh_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <typename F>
static float median_ms(F&& fn, int warm = 3, int iters = 9)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < warm; i++) fn();
    std::vector<float> t;
    for (int i = 0; i < iters; i++) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        fn();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

static int bit_length(uint32_t n) { int b = 0; while (n) { b++; n >>= 1; } return b ? b : 1; }

int main(int argc, char** argv)
{
    size_t n = argc > 1 ? (size_t)atof(argv[1]) : 32000000;
    uint32_t tiles = argc > 2 ? (uint32_t)atoi(argv[2]) : 16384;
    const int tile_bits = bit_length(tiles);
    printf("n=%zu tiles=%u tile_bits=%d\n", n, tiles, tile_bits);

    // renderer-like keys: instances of a face are a small rectangle of tiles (runs of adjacent ids), depth in a narrow band
    std::vector<uint32_t> hk(n), hv(n);
    std::vector<uint64_t> hk64(n);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); };
    const uint32_t tx = 1u << (tile_bits / 2);
    size_t i = 0;
    uint32_t face = 0;
    while (i < n) {
        uint32_t w = 1 + rnd() % 4, h = 1 + rnd() % 4;
        uint32_t t0 = rnd() % tiles;
        float depth = 0.73f + 0.24f * (float)(i) / (float)n;     // faces arrive in depth order
        uint32_t db; memcpy(&db, &depth, 4);
        for (uint32_t y = 0; y < h && i < n; y++)
            for (uint32_t x = 0; x < w && i < n; x++) {
                uint32_t t = (t0 + y * tx + x) % tiles;
                hk[i] = t; hv[i] = face & 0x3fffff; hk64[i] = ((uint64_t)t << 32) | db;
                i++;
            }
        face++;
    }
    uint32_t *k_in, *v_in, *k_out, *v_out, *k_ref, *v_ref;
    uint64_t *k64_in, *k64_out;
    CK(cudaMalloc(&k_in, 4 * n)); CK(cudaMalloc(&v_in, 4 * n)); CK(cudaMalloc(&k_out, 4 * n)); CK(cudaMalloc(&v_out, 4 * n));
    CK(cudaMalloc(&k_ref, 4 * n)); CK(cudaMalloc(&v_ref, 4 * n)); CK(cudaMalloc(&k64_in, 8 * n)); CK(cudaMalloc(&k64_out, 8 * n));
    CK(cudaMemcpy(k_in, hk.data(), 4 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v_in, hv.data(), 4 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(k64_in, hk64.data(), 8 * n, cudaMemcpyHostToDevice));

    size_t cub_bytes = 0, cub_bytes64 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k_in, k_ref, v_in, v_ref, (int)n, 0, 32);
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes64, k64_in, k64_out, v_in, v_ref, (int)n, 0, 64);
    void *cub_temp, *our_temp;
    CK(cudaMalloc(&cub_temp, std::max(cub_bytes, cub_bytes64)));
    CK(cudaMalloc(&our_temp, dmr_sort_temp_bytes(n)));

    auto ours = [&](int end_bit) { if (dmr_sort_pairs_u32(k_in, v_in, k_out, v_out, n, end_bit, our_temp, 0)) { printf("ours failed: %s\n", dmr_last_error()); exit(1); } };
    auto cub32 = [&](int end_bit) { size_t b = cub_bytes; cub::DeviceRadixSort::SortPairs(cub_temp, b, k_in, k_ref, v_in, v_ref, (int)n, 0, end_bit); };
    auto cub64 = [&](int end_bit) { size_t b = cub_bytes64; cub::DeviceRadixSort::SortPairs(cub_temp, b, k64_in, k64_out, v_in, v_ref, (int)n, 0, end_bit); };

    // correctness: identical permutation (both are stable sorts on the same bits)
    ours(tile_bits); cub32(tile_bits);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> a(n), b(n);
    CK(cudaMemcpy(a.data(), v_out, 4 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), v_ref, 4 * n, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t q = 0; q < n; q++) bad += a[q] != b[q];
    CK(cudaMemcpy(a.data(), k_out, 4 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), k_ref, 4 * n, cudaMemcpyDeviceToHost));
    for (size_t q = 0; q < n; q++) bad += a[q] != b[q];
    printf("ours vs CUB, sorted keys + values: %zu mismatches\n", bad);

    const double GB = 1e-9;
    float to = median_ms([&] { ours(tile_bits); }), tc = median_ms([&] { cub32(tile_bits); });
    int npass = (tile_bits + 7) / 8;
    printf("(a) u32 pairs, bits [0,%d) = %d pass(es):  ours %.3f ms   CUB %.3f ms   (ours/CUB = %.2f)\n", tile_bits, npass, to, tc, to / tc);
    float o8 = median_ms([&] { ours(8); }), o16 = median_ms([&] { ours(16); });
    float c8 = median_ms([&] { cub32(8); }), c16 = median_ms([&] { cub32(16); });
    double bytes = 16.0 * n;   // read + write of (u32 key, u32 value)
    printf("(b) one 8-bit pass over %zu pairs (16 B/pair = %.1f MB):  ours %.3f ms = %.0f GB/s   CUB %.3f ms = %.0f GB/s\n", n,
           bytes * 1e-6, o16 - o8, bytes * GB / ((o16 - o8) * 1e-3), c16 - c8, bytes * GB / ((c16 - c8) * 1e-3));
    printf("    raw: ours end_bit 8: %.3f  16: %.3f   CUB end_bit 8: %.3f  16: %.3f ms\n", o8, o16, c8, c16);
    float c64 = median_ms([&] { cub64(32 + tile_bits); });
    printf("(c) reference formulation: CUB u64 keys + u32 values on %d bits = %d passes: %.3f ms  (ours, two-level: tile sort above + a 4-pass sort of the faces)\n",
           32 + tile_bits, (32 + tile_bits + 7) / 8, c64);
    return 0;
}
