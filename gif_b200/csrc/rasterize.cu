// Deterministic tile-binned z-buffer rasteriser (forward) + barycentric backward.
//
// The reference (standard_rasterize_cuda_kernel.cu:112-233) runs one thread per triangle, resolves depth with a
// global-memory CAS loop per covered pixel, writes the winner's payload through uncoalesced scattered stores, and
// launches the whole kernel twice to paper over the write-after-atomicMin race.  Here:
//   1. bin_count / bin_scan / bin_fill: every front-facing triangle with a non-empty clamped bbox is appended to the
//      list of each 8x8-pixel bin its bbox touches (counting pass, per-image exclusive scan, fill pass);
//   2. raster: one warp per 8x4 half-bin, one pixel per lane.  The warp walks its bin's list (all lanes read the same
//      triangle -> broadcast loads), every lane evaluates the reference's barycentric formula for its own pixel and
//      keeps the best (zp, face) pair in registers; depth / triangle / payload are then written once, coalesced.
// No atomics on the outputs, a single pass, and the result is a pure function of the inputs (ties: lowest face index).
// fp32 arithmetic uses the non-contracting intrinsics (__fmul_rn/__fadd_rn/...) so it is bit-identical to
// oracle/rasterize_oracle.c built with -ffp-contract=off.
#include "common.cuh"

namespace gifb200 {

constexpr int BIN = 8;

struct RasterGeom {
    int B, F, h, w, bins_x, bins_y, nbins;  // nbins per image
};

__device__ __forceinline__ bool tri_setup(const float* __restrict__ fc, int w, int h, int& xmin, int& xmax, int& ymin,
                                          int& ymax) {
    const float x0 = fc[0], y0 = fc[1], x1 = fc[3], y1 = fc[4], x2 = fc[6], y2 = fc[7];
    // check_face_frontside (:32-34), separately rounded products
    const bool front = __fmul_rn(__fsub_rn(y2, y0), __fsub_rn(x1, x0)) < __fmul_rn(__fsub_rn(y1, y0), __fsub_rn(x2, x0));
    xmin = max(static_cast<int>(ceilf(fminf(x0, fminf(x1, x2)))), 0);     // :133-136
    xmax = min(static_cast<int>(floorf(fmaxf(x0, fmaxf(x1, x2)))), w - 1);
    ymin = max(static_cast<int>(ceilf(fminf(y0, fminf(y1, y2)))), 0);
    ymax = min(static_cast<int>(floorf(fmaxf(y0, fmaxf(y1, y2)))), h - 1);
    return front && xmin <= xmax && ymin <= ymax;
}

// pass 1 (fill == 0): count[b][bin] += 1 ; pass 2 (fill == 1): list[offset[b][bin] + cursor++] = f
template <int FILL>
__global__ void __launch_bounds__(256) bin_kernel(const float* __restrict__ fv, RasterGeom g, int* __restrict__ count,
                                                  const int* __restrict__ offset, int* __restrict__ list,
                                                  int capacity, int* __restrict__ overflow) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(g.B) * g.F) return;
    const int b = static_cast<int>(i / g.F), f = static_cast<int>(i % g.F);
    int xmin, xmax, ymin, ymax;
    if (!tri_setup(fv + i * 9, g.w, g.h, xmin, xmax, ymin, ymax)) return;
    const int bx0 = xmin / BIN, bx1 = xmax / BIN, by0 = ymin / BIN, by1 = ymax / BIN;
    for (int by = by0; by <= by1; ++by)
        for (int bx = bx0; bx <= bx1; ++bx) {
            const int bin = b * g.nbins + by * g.bins_x + bx;
            if (FILL == 0) {
                atomicAdd(count + bin, 1);
            } else {
                const int pos = offset[bin] + atomicAdd(count + bin, 1);   // count was re-zeroed: acts as cursor
                if (pos < (b + 1) * capacity) list[pos] = f;               // capacity = list slots per image
                else overflow[b] = 1;
            }
        }
}

// exclusive scan of one image's bin counts (n = bins per image) -- one 1024-thread CTA per image; image b's lists
// live in list[b*capacity, (b+1)*capacity).  The pass also saves the totals in `total` and re-zeroes `count`, which
// becomes the fill cursor and, after the fill pass, the list length again.
__global__ void __launch_bounds__(1024) bin_scan_kernel(int* __restrict__ count, int* __restrict__ offset, int n,
                                                        int capacity, int* __restrict__ overflow) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    count += static_cast<long long>(blockIdx.x) * n;
    offset += static_cast<long long>(blockIdx.x) * n;
    if (threadIdx.x == 0) carry = blockIdx.x * capacity;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? count[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) >= o) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = warp_tot[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (threadIdx.x >= o) t += u;
            }
            warp_tot[threadIdx.x] = t;
        }
        __syncthreads();
        const int warp_prefix = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
        if (i < n) {
            offset[i] = carry + warp_prefix + s - v;
            count[i] = 0;  // becomes the fill cursor
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0 && carry > (static_cast<int>(blockIdx.x) + 1) * capacity) overflow[blockIdx.x] = 1;
}

struct Frag {
    float zp;
    int f;
    float w0, w1, w2;
};

// barycentric_weight (:79-109) + inside test (:144) + depth (:148) for pixel (px,py); no contraction.
__device__ __forceinline__ bool shade(const float* __restrict__ fc, float px, float py, float& w0, float& w1, float& w2,
                                      float& zp) {
    const float x0 = fc[0], y0 = fc[1], x1 = fc[3], y1 = fc[4], x2 = fc[6], y2 = fc[7];
    const float v0x = __fsub_rn(x2, x0), v0y = __fsub_rn(y2, y0);
    const float v1x = __fsub_rn(x1, x0), v1y = __fsub_rn(y1, y0);
    const float v2x = __fsub_rn(px, x0), v2y = __fsub_rn(py, y0);
    const float d00 = __fadd_rn(__fmul_rn(v0x, v0x), __fmul_rn(v0y, v0y));
    const float d01 = __fadd_rn(__fmul_rn(v0x, v1x), __fmul_rn(v0y, v1y));
    const float d02 = __fadd_rn(__fmul_rn(v0x, v2x), __fmul_rn(v0y, v2y));
    const float d11 = __fadd_rn(__fmul_rn(v1x, v1x), __fmul_rn(v1y, v1y));
    const float d12 = __fadd_rn(__fmul_rn(v1x, v2x), __fmul_rn(v1y, v2y));
    const float den = __fsub_rn(__fmul_rn(d00, d11), __fmul_rn(d01, d01));
    const float inv = (den == 0.f) ? 0.f : __fdiv_rn(1.f, den);
    const float u = __fmul_rn(__fsub_rn(__fmul_rn(d11, d02), __fmul_rn(d01, d12)), inv);
    const float v = __fmul_rn(__fsub_rn(__fmul_rn(d00, d12), __fmul_rn(d01, d02)), inv);
    w0 = __fsub_rn(__fsub_rn(1.f, u), v);
    w1 = v;
    w2 = u;
    if (!(w2 >= 0.f && w1 >= 0.f && w0 > 0.f)) return false;
    const float s = __fadd_rn(__fadd_rn(__fdiv_rn(w0, fc[2]), __fdiv_rn(w1, fc[5])), __fdiv_rn(w2, fc[8]));
    zp = __double2float_rn(__ddiv_rn(1.0, static_cast<double>(s)));   // '1.' is a double literal in the reference
    return true;
}

// One warp per 8x4 half-bin.  brute != 0: ignore the lists and walk all F triangles (overflow fallback).
__global__ void __launch_bounds__(256) raster_kernel(const float* __restrict__ fv, const float* __restrict__ colors,
                                                     float* __restrict__ depth, int* __restrict__ tri,
                                                     float* __restrict__ out3, RasterGeom g,
                                                     const int* __restrict__ count, const int* __restrict__ offset,
                                                     const int* __restrict__ list, const int* __restrict__ overflow) {
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nwarps = static_cast<long long>(g.B) * g.nbins * 2;
    if (warp >= nwarps) return;
    const int half = static_cast<int>(warp & 1);
    const long long gbin = warp >> 1;
    const int b = static_cast<int>(gbin / g.nbins), bin = static_cast<int>(gbin % g.nbins);
    const int by = bin / g.bins_x, bx = bin % g.bins_x;
    const int px = bx * BIN + (lane & 7), py = by * BIN + half * 4 + (lane >> 3);
    const bool in_img = px < g.w && py < g.h;
    const long long pix = (static_cast<long long>(b) * g.h + py) * g.w + px;
    float best_z = in_img ? depth[pix] : 0.f;
    int best_f = 0x7fffffff;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    const bool brute = overflow[b] != 0;     // this image's lists did not fit: walk all F triangles (still exact)
    const int n = brute ? g.F : count[gbin];
    const int* lst = list + (brute ? 0 : offset[gbin]);
    const float fpx = static_cast<float>(px), fpy = static_cast<float>(py);
    const float* fvb = fv + static_cast<long long>(b) * g.F * 9;
    for (int j = 0; j < n; ++j) {
        const int f = brute ? j : lst[j];
        const float* fc = fvb + static_cast<long long>(f) * 9;
        if (brute) {
            int a, c, d, e;
            if (!tri_setup(fc, g.w, g.h, a, c, d, e)) continue;
            if (px < a || px > c || py < d || py > e) continue;   // the reference only visits bbox pixels
        } else {
            // the bbox test is part of the reference's semantics (pixels outside the clamped bbox are never visited)
            const float x0 = fc[0], y0 = fc[1], x1 = fc[3], y1 = fc[4], x2 = fc[6], y2 = fc[7];
            const int xmin = static_cast<int>(ceilf(fminf(x0, fminf(x1, x2))));
            const int xmax = static_cast<int>(floorf(fmaxf(x0, fmaxf(x1, x2))));
            const int ymin = static_cast<int>(ceilf(fminf(y0, fminf(y1, y2))));
            const int ymax = static_cast<int>(floorf(fmaxf(y0, fmaxf(y1, y2))));
            if (px < xmin || px > xmax || py < ymin || py > ymax) continue;
        }
        float w0, w1, w2, zp;
        if (!shade(fc, fpx, fpy, w0, w1, w2, zp)) continue;
        if (zp < best_z || (zp == best_z && f < best_f)) {
            best_z = zp; best_f = f; b0 = w0; b1 = w1; b2 = w2;
        }
    }
    if (in_img && best_f != 0x7fffffff) {
        depth[pix] = best_z;
        tri[pix] = best_f;
        float* o = out3 + pix * 3;
        if (colors) {
            const float* c = colors + (static_cast<long long>(b) * g.F + best_f) * 9;
#pragma unroll
            for (int k = 0; k < 3; ++k)   // :225, left-to-right, no contraction
                o[k] = __fadd_rn(__fadd_rn(__fmul_rn(b0, c[k]), __fmul_rn(b1, c[3 + k])), __fmul_rn(b2, c[6 + k]));
        } else {
            o[0] = b0; o[1] = b1; o[2] = b2;
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
// One thread per pixel; analytic derivatives of bw = (1-u-v, v, u) (kernel.cu:79-109), of zp (:148) and of the
// colour interpolation (:225) w.r.t. the owning face's vertices / colours, scatter-added with fp32 atomics.
__global__ void __launch_bounds__(256) raster_bwd_kernel(const float* __restrict__ fv, const float* __restrict__ colors,
                                                         const int* __restrict__ tri, const float* __restrict__ g_bary,
                                                         const float* __restrict__ g_img, const float* __restrict__ g_depth,
                                                         float* __restrict__ g_fv, float* __restrict__ g_col, int B, int F,
                                                         int h, int w) {
    const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (pix >= static_cast<long long>(B) * h * w) return;
    const int f = tri[pix];
    if (f < 0) return;
    const int px = static_cast<int>(pix % w), py = static_cast<int>((pix / w) % h), b = static_cast<int>(pix / (static_cast<long long>(w) * h));
    const long long fo = (static_cast<long long>(b) * F + f) * 9;
    const float* fc = fv + fo;
    const float x0 = fc[0], y0 = fc[1], z0 = fc[2], x1 = fc[3], y1 = fc[4], z1 = fc[5], x2 = fc[6], y2 = fc[7], z2 = fc[8];
    const float v0x = x2 - x0, v0y = y2 - y0, v1x = x1 - x0, v1y = y1 - y0, v2x = px - x0, v2y = py - y0;
    const float d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d02 = v0x * v2x + v0y * v2y;
    const float d11 = v1x * v1x + v1y * v1y, d12 = v1x * v2x + v1y * v2y;
    const float den = d00 * d11 - d01 * d01;
    const float inv = den == 0.f ? 0.f : 1.f / den;
    const float nu = d11 * d02 - d01 * d12, nv = d00 * d12 - d01 * d02;
    const float u = nu * inv, v = nv * inv;
    const float w0 = 1.f - u - v, w1 = v, w2 = u;
    // upstream gradient on (w0,w1,w2)
    float gw0 = 0.f, gw1 = 0.f, gw2 = 0.f;
    if (g_bary) { gw0 += g_bary[pix * 3]; gw1 += g_bary[pix * 3 + 1]; gw2 += g_bary[pix * 3 + 2]; }
    if (g_img && colors) {
        const float* c = colors + fo;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float gk = g_img[pix * 3 + k];
            gw0 += gk * c[k]; gw1 += gk * c[3 + k]; gw2 += gk * c[6 + k];
            if (g_col) {
                atomicAdd(g_col + fo + k, gk * w0);
                atomicAdd(g_col + fo + 3 + k, gk * w1);
                atomicAdd(g_col + fo + 6 + k, gk * w2);
            }
        }
    }
    if (g_depth) {
        const float s = w0 / z0 + w1 / z1 + w2 / z2;
        const float zp = 1.f / s;
        const float gz = g_depth[pix] * (-zp * zp);   // d zp / d s
        gw0 += gz / z0; gw1 += gz / z1; gw2 += gz / z2;
        atomicAdd(g_fv + fo + 2, gz * (-w0 / (z0 * z0)));
        atomicAdd(g_fv + fo + 5, gz * (-w1 / (z1 * z1)));
        atomicAdd(g_fv + fo + 8, gz * (-w2 / (z2 * z2)));
    }
    // (w0,w1,w2) = (1-u-v, v, u)  ->  gu = gw2 - gw0, gv = gw1 - gw0
    const float gu = gw2 - gw0, gv = gw1 - gw0;
    // u = nu*inv, v = nv*inv, inv = 1/den
    const float g_nu = gu * inv, g_nv = gv * inv;
    const float g_den = den == 0.f ? 0.f : -(gu * nu + gv * nv) * inv * inv;
    // nu = d11*d02 - d01*d12 ; nv = d00*d12 - d01*d02 ; den = d00*d11 - d01^2
    const float g_d00 = g_nv * d12 + g_den * d11;
    const float g_d11 = g_nu * d02 + g_den * d00;
    const float g_d01 = -g_nu * d12 - g_nv * d02 - 2.f * g_den * d01;
    const float g_d02 = g_nu * d11 - g_nv * d01;
    const float g_d12 = -g_nu * d01 + g_nv * d00;
    // dots -> vectors
    const float g_v0x = 2.f * g_d00 * v0x + g_d01 * v1x + g_d02 * v2x;
    const float g_v0y = 2.f * g_d00 * v0y + g_d01 * v1y + g_d02 * v2y;
    const float g_v1x = 2.f * g_d11 * v1x + g_d01 * v0x + g_d12 * v2x;
    const float g_v1y = 2.f * g_d11 * v1y + g_d01 * v0y + g_d12 * v2y;
    const float g_v2x = g_d02 * v0x + g_d12 * v1x;
    const float g_v2y = g_d02 * v0y + g_d12 * v1y;
    // v0 = p2-p0, v1 = p1-p0, v2 = p-p0
    atomicAdd(g_fv + fo + 0, -(g_v0x + g_v1x + g_v2x));
    atomicAdd(g_fv + fo + 1, -(g_v0y + g_v1y + g_v2y));
    atomicAdd(g_fv + fo + 3, g_v1x);
    atomicAdd(g_fv + fo + 4, g_v1y);
    atomicAdd(g_fv + fo + 6, g_v0x);
    atomicAdd(g_fv + fo + 7, g_v0y);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace gifb200

using namespace gifb200;

// workspace layout: [count: B*nbins ints][offset: B*nbins ints][overflow flags: B ints (padded)][list: B*capacity ints]
static int list_capacity(int F) {   // list slots per image; a FLAME render at 256^2 needs ~1.2 F, allow 8 F
    const long long cap = static_cast<long long>(F) * 8 + 1024;
    return static_cast<int>(cap > 0x3fffffffLL ? 0x3fffffffLL : cap);
}

extern "C" size_t gifb200_rasterize_workspace_bytes(int B, int F, int h, int w) {
    if (B <= 0 || h <= 0 || w <= 0 || F < 0) return 0;
    const size_t nb = static_cast<size_t>(B) * ((h + BIN - 1) / BIN) * ((w + BIN - 1) / BIN);
    return align_up(nb * 4, 256) * 2 + align_up(static_cast<size_t>(B) * 4, 256) + static_cast<size_t>(list_capacity(F)) * 4 * B;
}

extern "C" int gifb200_rasterize_fwd(const float* face_vertices, const float* face_colors, float* depth,
                                     int32_t* triangle, float* out3, int B, int F, int h, int w, void* workspace,
                                     size_t workspace_bytes, gifb200_stream_t stream) {
    GIFB200_REQUIRE(B >= 0 && F >= 0 && h > 0 && w > 0, GIFB200_E_SHAPE, "rasterize: bad shape");
    if (B == 0 || F == 0) return GIFB200_OK;     // nothing can be written: buffers keep the caller's initial values
    GIFB200_REQUIRE(workspace && workspace_bytes >= gifb200_rasterize_workspace_bytes(B, F, h, w), GIFB200_E_WORKSPACE,
                    "rasterize: workspace too small (see gifb200_rasterize_workspace_bytes)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RasterGeom g;
    g.B = B; g.F = F; g.h = h; g.w = w;
    g.bins_x = (w + BIN - 1) / BIN; g.bins_y = (h + BIN - 1) / BIN; g.nbins = g.bins_x * g.bins_y;
    const size_t nb = static_cast<size_t>(B) * g.nbins;
    GIFB200_REQUIRE(nb < 0x7fffffffULL / 64, GIFB200_E_SHAPE, "rasterize: too many bins");
    char* ws = static_cast<char*>(workspace);
    int* count = reinterpret_cast<int*>(ws);
    int* offset = reinterpret_cast<int*>(ws + align_up(nb * 4, 256));
    int* overflow = reinterpret_cast<int*>(ws + align_up(nb * 4, 256) * 2);
    int* list = reinterpret_cast<int*>(reinterpret_cast<char*>(overflow) + align_up(static_cast<size_t>(B) * 4, 256));
    const int capacity = list_capacity(F);
    GIFB200_REQUIRE(static_cast<long long>(B) * capacity < 0x7fffffffLL, GIFB200_E_SHAPE, "rasterize: B*F too large");
    cudaError_t e = cudaMemsetAsync(ws, 0, align_up(nb * 4, 256) * 2 + align_up(static_cast<size_t>(B) * 4, 256), st);
    if (e != cudaSuccess) return fail(GIFB200_E_CUDA, "rasterize memset", cudaGetErrorString(e));
    const long long ntri = static_cast<long long>(B) * F;
    bin_kernel<0><<<cdiv(ntri, 256), 256, 0, st>>>(face_vertices, g, count, offset, list, capacity, overflow);
    GIFB200_LAUNCH_CHECK("bin_kernel<count>");
    bin_scan_kernel<<<B, 1024, 0, st>>>(count, offset, g.nbins, capacity, overflow);
    GIFB200_LAUNCH_CHECK("bin_scan_kernel");
    bin_kernel<1><<<cdiv(ntri, 256), 256, 0, st>>>(face_vertices, g, count, offset, list, capacity, overflow);
    GIFB200_LAUNCH_CHECK("bin_kernel<fill>");
    const long long nwarps = static_cast<long long>(nb) * 2;
    raster_kernel<<<cdiv(nwarps * 32, 256), 256, 0, st>>>(face_vertices, face_colors, depth, triangle, out3, g, count,
                                                          offset, list, overflow);
    GIFB200_LAUNCH_CHECK("raster_kernel");
    return GIFB200_OK;
}

extern "C" int gifb200_rasterize_bwd(const float* face_vertices, const float* face_colors, const int32_t* triangle,
                                     const float* g_bary, const float* g_img, const float* g_depth,
                                     float* g_face_vertices, float* g_face_colors, int B, int F, int h, int w,
                                     gifb200_stream_t stream) {
    GIFB200_REQUIRE(B >= 0 && F >= 0 && h > 0 && w > 0, GIFB200_E_SHAPE, "rasterize_bwd: bad shape");
    GIFB200_REQUIRE(g_face_vertices != nullptr, GIFB200_E_SHAPE, "rasterize_bwd: g_face_vertices is required");
    GIFB200_REQUIRE(!g_img || face_colors, GIFB200_E_SHAPE, "rasterize_bwd: g_img needs face_colors");
    if (B == 0 || F == 0) return GIFB200_OK;
    const long long npix = static_cast<long long>(B) * h * w;
    raster_bwd_kernel<<<cdiv(npix, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        face_vertices, face_colors, triangle, g_bary, g_img, g_depth, g_face_vertices, g_face_colors, B, F, h, w);
    GIFB200_LAUNCH_CHECK("raster_bwd_kernel");
    return GIFB200_OK;
}
