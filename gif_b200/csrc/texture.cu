// Texture "stealing": project every texel of the FLAME UV atlas into the image through the posed mesh and sample the
// generated image there (reference: FlameTextureSpace.compute_texture_map, model/stg2_generator.py:376-421 -- three gathers
// of (B,N,3) vertex tensors, a scatter into a (B,256,256,2) grid, F.grid_sample, three more gathers for the normals and a
// scatter for the mask).  Here: ONE pass, one thread per (sample, texel); SURVEY 8f.2.
//   valid texel n: p = sum_k bary[n][k] * verts[b][vid[n][k]];  g = (s (p.x + tx), -s (p.y + ty));  tex = bilinear(src, g)
//                  mask = (sum_k bary[n][k] * normal_z[b][vid[n][k]]) < 0
//   other texels:  g = (0,0) (the reference leaves its grid at zero there, i.e. they sample the image centre), mask = 0
// grid_sample semantics: bilinear, zeros padding, align_corners = False (pixel = ((g + 1) * size - 1) / 2).
// HBM-bound and tiny: B*T*T*(C*4 + 1) bytes out, table 28 B / valid texel, gathers hit L2 (mesh 60 KB, image <= 786 KB / sample).
#include "common.cuh"

namespace gifb200 {

struct Bilinear {
    int x0, y0;
    float w00, w01, w10, w11;   // (y0,x0), (y0,x0+1), (y0+1,x0), (y0+1,x0+1)
};

__device__ __forceinline__ Bilinear bilinear_setup(float gx, float gy, int W, int H) {
    const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    Bilinear b;
    b.x0 = static_cast<int>(fx); b.y0 = static_cast<int>(fy);
    const float ex = fx + 1.f - ix, ey = fy + 1.f - iy, dx = ix - fx, dy = iy - fy;
    b.w00 = ex * ey; b.w01 = dx * ey; b.w10 = ex * dy; b.w11 = dx * dy;
    return b;
}

__device__ __forceinline__ bool texel_grid(int b, int t, const int* __restrict__ texel_to_valid, const int* __restrict__ vid,
                                           const float* __restrict__ bary, const float* __restrict__ verts,
                                           const float* __restrict__ cam, int V, float& gx, float& gy, int& n) {
    n = texel_to_valid[t];
    gx = 0.f; gy = 0.f;
    if (n < 0) return false;
    const float* vb = verts + static_cast<long long>(b) * V * 3;
    float px = 0.f, py = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int v = vid[n * 3 + k];
        const float w = bary[n * 3 + k];
        px = __fadd_rn(px, __fmul_rn(vb[v * 3 + 0], w));
        py = __fadd_rn(py, __fmul_rn(vb[v * 3 + 1], w));
    }
    const float s = cam[b * 3 + 0];
    gx = s * (px + cam[b * 3 + 1]);
    gy = -(s * (py + cam[b * 3 + 2]));
    return true;
}

__global__ void __launch_bounds__(256) texture_steal_fwd_kernel(const float* __restrict__ src, const float* __restrict__ verts,
                                                               const float* __restrict__ normals,
                                                               const float* __restrict__ cam,
                                                               const int* __restrict__ texel_to_valid,
                                                               const int* __restrict__ vid, const float* __restrict__ bary,
                                                               float* __restrict__ tex, unsigned char* __restrict__ mask,
                                                               int B, int H, int W, int C, int V, int TT) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long long>(B) * TT) return;
    const int b = static_cast<int>(e / TT), t = static_cast<int>(e % TT);
    float gx, gy;
    int n;
    const bool valid = texel_grid(b, t, texel_to_valid, vid, bary, verts, cam, V, gx, gy, n);
    const Bilinear bl = bilinear_setup(gx, gy, W, H);
    const float* sb = src + static_cast<long long>(b) * H * W * C;
    const bool x0 = bl.x0 >= 0 && bl.x0 < W, x1 = bl.x0 + 1 >= 0 && bl.x0 + 1 < W;
    const bool y0 = bl.y0 >= 0 && bl.y0 < H, y1 = bl.y0 + 1 >= 0 && bl.y0 + 1 < H;
    float* o = tex + e * C;
    for (int c = 0; c < C; ++c) {
        float a = 0.f;
        if (y0 && x0) a += sb[(static_cast<long long>(bl.y0) * W + bl.x0) * C + c] * bl.w00;
        if (y0 && x1) a += sb[(static_cast<long long>(bl.y0) * W + bl.x0 + 1) * C + c] * bl.w01;
        if (y1 && x0) a += sb[(static_cast<long long>(bl.y0 + 1) * W + bl.x0) * C + c] * bl.w10;
        if (y1 && x1) a += sb[(static_cast<long long>(bl.y0 + 1) * W + bl.x0 + 1) * C + c] * bl.w11;
        o[c] = a;
    }
    if (mask) {
        unsigned char m = 0;
        if (valid && normals) {
            const float* nb = normals + static_cast<long long>(b) * V * 3;
            float nz = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) nz = __fadd_rn(nz, __fmul_rn(nb[vid[n * 3 + k] * 3 + 2], bary[n * 3 + k]));
            m = nz < 0.f;
        }
        mask[e] = m;
    }
}

// adjoint w.r.t. the image: g_src[b, y, x, c] += w * g_tex[b, t, c]  (g_src zero-filled by the entry point)
__global__ void __launch_bounds__(256) texture_steal_bwd_kernel(const float* __restrict__ g_tex, const float* __restrict__ verts,
                                                               const float* __restrict__ cam,
                                                               const int* __restrict__ texel_to_valid,
                                                               const int* __restrict__ vid, const float* __restrict__ bary,
                                                               float* __restrict__ g_src, int B, int H, int W, int C, int V,
                                                               int TT) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long long>(B) * TT) return;
    const int b = static_cast<int>(e / TT), t = static_cast<int>(e % TT);
    float gx, gy;
    int n;
    texel_grid(b, t, texel_to_valid, vid, bary, verts, cam, V, gx, gy, n);
    const Bilinear bl = bilinear_setup(gx, gy, W, H);
    float* gb = g_src + static_cast<long long>(b) * H * W * C;
    const bool x0 = bl.x0 >= 0 && bl.x0 < W, x1 = bl.x0 + 1 >= 0 && bl.x0 + 1 < W;
    const bool y0 = bl.y0 >= 0 && bl.y0 < H, y1 = bl.y0 + 1 >= 0 && bl.y0 + 1 < H;
    const float* g = g_tex + e * C;
    for (int c = 0; c < C; ++c) {
        const float gv = g[c];
        if (gv == 0.f) continue;
        if (y0 && x0) atomicAdd(gb + (static_cast<long long>(bl.y0) * W + bl.x0) * C + c, gv * bl.w00);
        if (y0 && x1) atomicAdd(gb + (static_cast<long long>(bl.y0) * W + bl.x0 + 1) * C + c, gv * bl.w01);
        if (y1 && x0) atomicAdd(gb + (static_cast<long long>(bl.y0 + 1) * W + bl.x0) * C + c, gv * bl.w10);
        if (y1 && x1) atomicAdd(gb + (static_cast<long long>(bl.y0 + 1) * W + bl.x0 + 1) * C + c, gv * bl.w11);
    }
}

}  // namespace gifb200

using namespace gifb200;

extern "C" int gifb200_texture_steal_fwd(const float* src, const float* verts, const float* normals, const float* cam,
                                         const int32_t* texel_to_valid, const int32_t* vid, const float* bary, float* tex,
                                         unsigned char* mask, int B, int H, int W, int C, int V, int T,
                                         gifb200_stream_t stream) {
    GIFB200_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && V > 0 && T > 0 && T <= 8192, GIFB200_E_SHAPE, "texture_steal_fwd: bad shape");
    if (B == 0) return GIFB200_OK;
    const long long total = static_cast<long long>(B) * T * T;
    GIFB200_REQUIRE(total <= 2147483647LL * 256, GIFB200_E_SHAPE, "texture_steal_fwd: too many texels");
    texture_steal_fwd_kernel<<<cdiv(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, verts, normals, cam, texel_to_valid, vid, bary, tex, mask, B, H, W, C, V, T * T);
    GIFB200_LAUNCH_CHECK("texture_steal_fwd_kernel");
    return GIFB200_OK;
}

extern "C" int gifb200_texture_steal_bwd(const float* g_tex, const float* verts, const float* cam,
                                         const int32_t* texel_to_valid, const int32_t* vid, const float* bary, float* g_src,
                                         int B, int H, int W, int C, int V, int T, gifb200_stream_t stream) {
    GIFB200_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && V > 0 && T > 0 && T <= 8192, GIFB200_E_SHAPE, "texture_steal_bwd: bad shape");
    if (B == 0) return GIFB200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(g_src, 0, sizeof(float) * static_cast<size_t>(B) * H * W * C, st);
    if (e != cudaSuccess) return fail(GIFB200_E_CUDA, "texture_steal_bwd memset", cudaGetErrorString(e));
    const long long total = static_cast<long long>(B) * T * T;
    texture_steal_bwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(g_tex, verts, cam, texel_to_valid, vid, bary, g_src, B, H, W, C,
                                                               V, T * T);
    GIFB200_LAUNCH_CHECK("texture_steal_bwd_kernel");
    return GIFB200_OK;
}
