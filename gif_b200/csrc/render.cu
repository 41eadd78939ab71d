// Fused shading epilogue of the FLAME conditioning render: per pixel, from the rasteriser's (triangle, bary)
// buffers, interpolate the UV coordinates and the world-space vertex normals of the owning face, fetch the albedo with
// grid_sample semantics (bilinear, align_corners=False, zero padding), evaluate the 9-term spherical-harmonics shading,
// and write the textured image, the normal image and (optionally) the quantised 6-channel condition map that the
// generator consumes -- one pass, nothing but the final maps goes to HBM.
// Replaces the attribute gather + interpolation of Pytorch3dRasterizer.forward (renderer.py:69-84), Renderer.forward's
// albedo lookup / add_SHlight / image composition (renderer.py:152-221), Renderer.render_normal (renderer.py:291-305)
// and the quantisation of OverLayViz.get_rendered_mesh (visualize_flame_overlay.py:29-31) + the consumer's mapping to
// [-1,1] (loss_functions/losses.py:213-214).
#include "common.cuh"

namespace gifb200 {

struct ShadeParams {
    int B, F, h, w, T;
};

__device__ __forceinline__ float tex_fetch(const float* __restrict__ plane, int T, int x, int y) {
    return (x >= 0 && x < T && y >= 0 && y < T) ? __ldg(plane + y * T + x) : 0.f;
}

__global__ void __launch_bounds__(256) render_shade_kernel(const int* __restrict__ tri, const float* __restrict__ bary,
                                                           const float* __restrict__ face_uv,       // (F,3,2) shared
                                                           const float* __restrict__ face_normals,  // (B,F,3,3)
                                                           const float* __restrict__ albedo,        // (B,3,T,T)
                                                           const float* __restrict__ sh,            // (B,9,3)
                                                           float* __restrict__ tex, float* __restrict__ nrm,
                                                           float* __restrict__ cond, ShadeParams p) {
    const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long npix = static_cast<long long>(p.B) * p.h * p.w;
    if (pix >= npix) return;
    const int b = static_cast<int>(pix / (static_cast<long long>(p.h) * p.w));
    const int f = tri[pix];
    float t3[3] = {0.f, 0.f, 0.f}, n3[3] = {0.f, 0.f, 0.f};
    if (f >= 0) {
        const float w0 = bary[pix * 3], w1 = bary[pix * 3 + 1], w2 = bary[pix * 3 + 2];
        const float* uv = face_uv + static_cast<long long>(f) * 6;
        const float gu = w0 * uv[0] + w1 * uv[2] + w2 * uv[4];
        const float gv = w0 * uv[1] + w1 * uv[3] + w2 * uv[5];
        const float* fn = face_normals + (static_cast<long long>(b) * p.F + f) * 9;
#pragma unroll
        for (int k = 0; k < 3; ++k) n3[k] = w0 * fn[k] + w1 * fn[3 + k] + w2 * fn[6 + k];
        // F.grid_sample(albedo, grid, mode='bilinear', padding_mode='zeros', align_corners=False)
        const float ix = ((gu + 1.f) * p.T - 1.f) * 0.5f, iy = ((gv + 1.f) * p.T - 1.f) * 0.5f;
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
        const float ax = ix - fx, ay = iy - fy;
        // SH basis (renderer.py:207-221) with the constant factors of renderer.py:119-126
        const float pi = 3.14159265358979323846f;
        const float c0 = 1.f / sqrtf(4.f * pi), c1 = (2.f * pi / 3.f) * sqrtf(3.f / (4.f * pi));
        const float c2 = (pi / 4.f) * 3.f * sqrtf(5.f / (12.f * pi)), c3 = (pi / 4.f) * 1.5f * sqrtf(5.f / (12.f * pi));
        const float c4 = (pi / 4.f) * 0.5f * sqrtf(5.f / (4.f * pi));
        const float nx = n3[0], ny = n3[1], nz = n3[2];
        const float basis[9] = {c0, c1 * nx, c1 * ny, c1 * nz, c2 * nx * ny, c2 * nx * nz, c2 * ny * nz,
                                c3 * (nx * nx - ny * ny), c4 * (3.f * nz * nz - 1.f)};
        const float* shb = sh + static_cast<long long>(b) * 27;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* plane = albedo + (static_cast<long long>(b) * 3 + c) * p.T * p.T;
            const float a = (1.f - ay) * ((1.f - ax) * tex_fetch(plane, p.T, x0, y0) + ax * tex_fetch(plane, p.T, x0 + 1, y0)) +
                            ay * ((1.f - ax) * tex_fetch(plane, p.T, x0, y0 + 1) + ax * tex_fetch(plane, p.T, x0 + 1, y0 + 1));
            float shade = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) shade += shb[k * 3 + c] * basis[k];
            t3[c] = a * shade;            // * alpha (= 1 on covered pixels)
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (tex) tex[pix * 3 + c] = t3[c];
        if (nrm) nrm[pix * 3 + c] = n3[c];
    }
    if (cond) {
        // texture: floor(clamp(x,0,255))/255 ; normals: floor(clamp(n,0,1)*255)/255 ; then clamp(0,1)*2-1
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float tq = floorf(fminf(fmaxf(t3[c], 0.f), 255.f)) / 255.f;
            const float nq = floorf(fminf(fmaxf(n3[c], 0.f), 1.f) * 255.f) / 255.f;
            cond[pix * 6 + c] = fminf(fmaxf(tq, 0.f), 1.f) * 2.f - 1.f;
            cond[pix * 6 + 3 + c] = fminf(fmaxf(nq, 0.f), 1.f) * 2.f - 1.f;
        }
    }
}

}  // namespace gifb200

using namespace gifb200;

extern "C" int gifb200_render_shade(const int32_t* triangle, const float* bary, const float* face_uv,
                                    const float* face_normals, const float* albedo, const float* sh, float* tex,
                                    float* nrm, float* cond, int B, int F, int h, int w, int T, gifb200_stream_t stream) {
    GIFB200_REQUIRE(B >= 0 && F > 0 && h > 0 && w > 0 && T > 0, GIFB200_E_SHAPE, "render_shade: bad shape");
    if (B == 0) return GIFB200_OK;
    ShadeParams p{B, F, h, w, T};
    const long long npix = static_cast<long long>(B) * h * w;
    render_shade_kernel<<<cdiv(npix, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(triangle, bary, face_uv, face_normals,
                                                                                        albedo, sh, tex, nrm, cond, p);
    GIFB200_LAUNCH_CHECK("render_shade_kernel");
    return GIFB200_OK;
}
