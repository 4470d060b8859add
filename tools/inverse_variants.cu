// Dev tool: which fp32 operation order reproduces torch.inverse (LU with partial pivoting + two triangular solves
// against the permuted identity) bit for bit on 4x4 matrices?  One thread per matrix, `variant` selects the arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC tools/inverse_variants.cu -o tools/_bin/libinv.so
#include <cuda_runtime.h>
#include <cstdint>

// variant bits: 0 LU scales by reciprocal (else divides)   1 LU update fused (fma)   2 solve update fused
//               3 upper solve multiplies by reciprocal      4 upper solve row-oriented ascending j (else column-oriented)
//               5 lower solve row-oriented descending j     6 reciprocal via __frcp_rn vs 1.0f/x (same thing, sanity)
__global__ void inverse4_kernel(const float* __restrict__ in, float* __restrict__ out, int* __restrict__ info, int n,
                                int variant) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const bool lu_rcp = variant & 1, lu_fma = variant & 2, sv_fma = variant & 4, up_rcp = variant & 8,
               up_row = variant & 16, lo_row = variant & 32;
    float a[4][4], b[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { a[r][c] = in[16 * m + 4 * r + c]; b[r][c] = (r == c) ? 1.0f : 0.0f; }
    int inf = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int p = k; float best = fabsf(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) { float v = fabsf(a[i][k]); if (v > best) { best = v; p = i; } }
#pragma unroll
        for (int i = k + 1; i < 4; ++i) if (p == i) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { float t = a[k][c]; a[k][c] = a[i][c]; a[i][c] = t; t = b[k][c]; b[k][c] = b[i][c]; b[i][c] = t; }
        }
        if (a[k][k] == 0.0f) { if (!inf) inf = k + 1; continue; }
        float r = __frcp_rn(a[k][k]);
#pragma unroll
        for (int i = k + 1; i < 4; ++i) {
            float l = lu_rcp ? __fmul_rn(a[i][k], r) : __fdiv_rn(a[i][k], a[k][k]);
            a[i][k] = l;
#pragma unroll
            for (int j = k + 1; j < 4; ++j)
                a[i][j] = lu_fma ? __fmaf_rn(-l, a[k][j], a[i][j]) : __fsub_rn(a[i][j], __fmul_rn(l, a[k][j]));
        }
    }
    // L y = P I   (unit lower)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (!lo_row) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = j + 1; i < 4; ++i)
                    b[i][c] = sv_fma ? __fmaf_rn(-a[i][j], b[j][c], b[i][c]) : __fsub_rn(b[i][c], __fmul_rn(a[i][j], b[j][c]));
        } else {
#pragma unroll
            for (int i = 1; i < 4; ++i)
#pragma unroll
                for (int j = i - 1; j >= 0; --j)
                    b[i][c] = sv_fma ? __fmaf_rn(-a[i][j], b[j][c], b[i][c]) : __fsub_rn(b[i][c], __fmul_rn(a[i][j], b[j][c]));
        }
        // U x = y
        if (!up_row) {
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                b[j][c] = up_rcp ? __fmul_rn(b[j][c], __frcp_rn(a[j][j])) : __fdiv_rn(b[j][c], a[j][j]);
#pragma unroll
                for (int i = 0; i < j; ++i)
                    b[i][c] = sv_fma ? __fmaf_rn(-a[i][j], b[j][c], b[i][c]) : __fsub_rn(b[i][c], __fmul_rn(a[i][j], b[j][c]));
            }
        } else {
#pragma unroll
            for (int i = 3; i >= 0; --i) {
#pragma unroll
                for (int j = i + 1; j < 4; ++j)
                    b[i][c] = sv_fma ? __fmaf_rn(-a[i][j], b[j][c], b[i][c]) : __fsub_rn(b[i][c], __fmul_rn(a[i][j], b[j][c]));
                b[i][c] = up_rcp ? __fmul_rn(b[i][c], __frcp_rn(a[i][i])) : __fdiv_rn(b[i][c], a[i][i]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[16 * m + 4 * r + c] = b[r][c];
    info[m] = inf;
}

extern "C" int inv4_run(const float* in, float* out, int* info, int n, int variant, void* stream) {
    inverse4_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(in, out, info, n, variant);
    return (int)cudaGetLastError();
}
