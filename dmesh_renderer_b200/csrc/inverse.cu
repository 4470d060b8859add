// inverse.cu -- inverse(mv_mats), inverse(proj_mats) for a stack of cameras in ONE launch.
//
// The reference's Python computes them with two th.inverse calls per forward (dmesh_renderer/__init__.py:62-63,
// 298-299).  On the GPU each of those is ~12 tiny library kernels (LU factorisation with partial pivoting, row
// swaps of the identity, two triangular solves) plus a device synchronisation to read LAPACK's `info`: ~45 us of
// serialised GPU time per call even when replayed from a CUDA graph -- 10% of a whole C2 step.  Here one thread
// per matrix does the same arithmetic in registers:
//     right-looking LU, first-maximum partial pivoting, multipliers = a_ik * (1 / a_kk), fused updates;
//     L y = P I  column-oriented (ascending j), fused;
//     U x = y    column-oriented (descending j), x_j = y_j / u_jj (true division), fused updates.
// That operation order reproduces torch.inverse BIT FOR BIT on this stack (torch 2.11 / CUDA 12.9 library kernels,
// B = 1 and B > 1 alike): tools/inverse_variants.py tried the 64 plausible orders on 100 k random and camera
// matrices and exactly this one matched everywhere; tests/test_gpu_tri_vs_reference.py keeps checking it.
// Every operation is a pinned intrinsic so that no compiler version can re-associate or (un)fuse it.
//
// The kernel also writes contiguous copies of the two input stacks (the API hands them over as transposed views;
// the preprocess kernels want [B,16] rows), which replaces two more torch copy kernels.
#include "common.cuh"
#include "../../include/dmesh_b200.h"

namespace dmr {

struct MatStrides { long long batch, row, col; };

__global__ void __launch_bounds__(64) camera_inverses_kernel(int B, const float* __restrict__ mv, MatStrides smv,
                                                              const float* __restrict__ proj, MatStrides spj,
                                                              float* __restrict__ out, int32_t* info)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;      // 0..B-1: mv, B..2B-1: proj
    if (m >= 2 * B) return;
    const bool is_proj = m >= B;
    const int b = is_proj ? m - B : m;
    const float* src = is_proj ? proj : mv;
    const MatStrides s = is_proj ? spj : smv;
    float a[4][4], x[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            a[r][c] = src[b * s.batch + r * s.row + c * s.col];
            x[r][c] = (r == c) ? 1.0f : 0.0f;
        }
    float4* copy = reinterpret_cast<float4*>(out + 16 * (size_t)m);               // slabs 0, 1: the inputs, contiguous
#pragma unroll
    for (int r = 0; r < 4; ++r) copy[r] = make_float4(a[r][0], a[r][1], a[r][2], a[r][3]);

    int inf = 0;                                                                   // LAPACK getrf: first zero pivot, 1-based
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int p = k;
        float best = fabsf(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) { const float v = fabsf(a[i][k]); if (v > best) { best = v; p = i; } }
#pragma unroll
        for (int i = k + 1; i < 4; ++i)
            if (p == i) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float t = a[k][c]; a[k][c] = a[i][c]; a[i][c] = t;
                    t = x[k][c]; x[k][c] = x[i][c]; x[i][c] = t;
                }
            }
        if (a[k][k] == 0.0f) { if (!inf) inf = k + 1; continue; }
        const float rcp = __frcp_rn(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) {
            const float l = __fmul_rn(a[i][k], rcp);
            a[i][k] = l;
#pragma unroll
            for (int j = k + 1; j < 4; ++j) a[i][j] = __fmaf_rn(-l, a[k][j], a[i][j]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = j + 1; i < 4; ++i) x[i][c] = __fmaf_rn(-a[i][j], x[j][c], x[i][c]);
#pragma unroll
        for (int j = 3; j >= 0; --j) {
            x[j][c] = __fdiv_rn(x[j][c], a[j][j]);
#pragma unroll
            for (int i = 0; i < j; ++i) x[i][c] = __fmaf_rn(-a[i][j], x[j][c], x[i][c]);
        }
    }
    float4* inv = reinterpret_cast<float4*>(out + 16 * (size_t)(2 * B + m));       // slabs 2, 3: the inverses
#pragma unroll
    for (int r = 0; r < 4; ++r) inv[r] = make_float4(x[r][0], x[r][1], x[r][2], x[r][3]);
    info[m] = inf;
}

}  // namespace dmr

using namespace dmr;

extern "C" int dmr_camera_inverses(int B, const float* mv_mats, int64_t mv_batch_stride, int64_t mv_row_stride,
                                   int64_t mv_col_stride, const float* proj_mats, int64_t proj_batch_stride,
                                   int64_t proj_row_stride, int64_t proj_col_stride, float* out, int32_t* info,
                                   dmr_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0) { set_error("negative batch size"); return DMR_EINVAL; }
    if (B == 0) return DMR_OK;
    if (!mv_mats || !proj_mats || !out || !info) { set_error("null pointer"); return DMR_EINVAL; }
    if (((uintptr_t)out & 15) != 0) { set_error("out must be 16-byte aligned"); return DMR_EINVAL; }
    MatStrides a = { mv_batch_stride, mv_row_stride, mv_col_stride }, b = { proj_batch_stride, proj_row_stride, proj_col_stride };
    count_launch(1);
    camera_inverses_kernel<<<(2 * B + 63) / 64, 64, 0, stream>>>(B, mv_mats, a, proj_mats, b, out, info);
    DMR_LAUNCH_CHECK("camera_inverses_kernel");
    return DMR_OK;
}
