// Deterministic tile-binned z-buffer rasteriser (forward) + per-face gather backward, two conventions.
//
// The reference (standard_rasterize_cuda_kernel.cu:112-233) runs one thread per triangle, resolves depth with a
// global-memory CAS loop per covered pixel, writes the winner's payload through uncoalesced scattered stores, and
// launches the whole kernel twice to paper over the write-after-atomicMin race.  Here:
//   1. bin_kernel<count> / bin_scan / bin_kernel<fill>: every triangle that passes the convention's face test and has a
//      non-empty clamped bbox is appended to the list of each 16x16-pixel tile its bbox touches;
//   2. raster_tile_kernel: one CTA per (image, tile).  The tile's depth buffer lives in SHARED memory as 64-bit keys
//      (order-preserving depth bits << 32 | face index), initialised from the caller's depth buffer.  Each thread sets up
//      one triangle of the list (once), a CTA-wide prefix sum over the clipped bbox areas flattens all fragments of the
//      tile into one index space that the 256 threads walk evenly (the work is the number of bbox pixels, like the
//      reference's per-triangle threads, not tile pixels x list length -- and no lane idles behind a large triangle), and
//      depth is resolved with shared-memory atomicMin on the key: min depth
//      wins, exact ties are won by the lowest face index, a fragment must be closer than (or tie) the caller's initial
//      depth -- a pure function of the inputs.  After a barrier one thread per pixel re-evaluates the winner's
//      barycentrics with the same arithmetic and writes depth / triangle / payload once, coalesced; up to two attribute
//      sets (e.g. vertex colours AND vertex normals: "texture + normal render") are interpolated in the same pass.
//   3. raster_bwd_face_kernel: one thread per triangle GATHERS the gradient of the pixels it owns (triangle buffer ==
//      its index) over its bbox and writes its 9 (+9 per attribute set) outputs once: no atomics, no zero-initialised
//      accumulator, deterministic summation order.  (Round 1: 6-15 fp32 atomics per covered pixel.)
//
// Conventions (template parameter CONV):
//   0  the in-repo standard_rasterize semantics (kernel.cu:31-34 front faces only, :79-109 barycentric_weight, :133-136
//      clamped integer bbox, :144 asymmetric inside test, :148 perspective-interpolated depth): pixel-space input, pixel
//      centres at integer coordinates.  fp32 arithmetic uses the non-contracting intrinsics so it is bit-identical to
//      oracle/rasterize_oracle.c built with -ffp-contract=off (and to the reference kernels compiled for the host).
//   1  pytorch3d's rasterize_meshes as the reference calls it (photometric_optimization/renderer.py:35-67: blur_radius 0,
//      faces_per_pixel 1, perspective_correct False): NDC input (the caller has already negated x, y, renderer.py:55),
//      +X left / +Y up so pixel (yi, xi) samples NDC (PixToNdc(W-1-xi), PixToNdc(H-1-yi)) with PixToNdc(i, S) =
//      -1 + (2 i + 1) / S, edge-function barycentrics over (area + 1e-8), no back-face culling, faces with zmax < 0 or
//      |area| <= 1e-8 skipped, strictly-inside test (all three > 0), linear depth pz = sum w_i z_i >= 0, min depth wins.
//      The fork the reference pins is absent and unversioned (requirements.txt:36): PARITY UNPINNED; the checker is the
//      restatement of these published rules in oracle/rasterize_oracle.c.
#include "common.cuh"

namespace gifb200 {

constexpr int TILE = 16;
constexpr int kNoFace = 0x7fffffff;
constexpr int kBwdLanes = 8;          // lanes per triangle in the backward gather
constexpr float kP3dEps = 1e-8f;

struct RasterGeom {
    int B, F, h, w, bins_x, bins_y, nbins;  // nbins per image
};

// ---------------------------------------------------------------------------------------------- convention 0
struct Setup0 {
    float x0, y0, v0x, v0y, v1x, v1y, d00, d01, d11, inv, z0, z1, z2;
};

__device__ __forceinline__ bool bbox0(const float* __restrict__ fc, int w, int h, int& xmin, int& xmax, int& ymin, int& ymax) {
    const float x0 = fc[0], y0 = fc[1], x1 = fc[3], y1 = fc[4], x2 = fc[6], y2 = fc[7];
    // check_face_frontside (:32-34), separately rounded products
    const bool front = __fmul_rn(__fsub_rn(y2, y0), __fsub_rn(x1, x0)) < __fmul_rn(__fsub_rn(y1, y0), __fsub_rn(x2, x0));
    xmin = max(static_cast<int>(ceilf(fminf(x0, fminf(x1, x2)))), 0);     // :133-136
    xmax = min(static_cast<int>(floorf(fmaxf(x0, fmaxf(x1, x2)))), w - 1);
    ymin = max(static_cast<int>(ceilf(fminf(y0, fminf(y1, y2)))), 0);
    ymax = min(static_cast<int>(floorf(fmaxf(y0, fmaxf(y1, y2)))), h - 1);
    return front && xmin <= xmax && ymin <= ymax;
}

__device__ __forceinline__ Setup0 setup0(const float* __restrict__ fc) {
    Setup0 s;
    const float x0 = fc[0], y0 = fc[1], x1 = fc[3], y1 = fc[4], x2 = fc[6], y2 = fc[7];
    s.x0 = x0; s.y0 = y0;
    s.v0x = __fsub_rn(x2, x0); s.v0y = __fsub_rn(y2, y0);
    s.v1x = __fsub_rn(x1, x0); s.v1y = __fsub_rn(y1, y0);
    s.d00 = __fadd_rn(__fmul_rn(s.v0x, s.v0x), __fmul_rn(s.v0y, s.v0y));
    s.d01 = __fadd_rn(__fmul_rn(s.v0x, s.v1x), __fmul_rn(s.v0y, s.v1y));
    s.d11 = __fadd_rn(__fmul_rn(s.v1x, s.v1x), __fmul_rn(s.v1y, s.v1y));
    const float den = __fsub_rn(__fmul_rn(s.d00, s.d11), __fmul_rn(s.d01, s.d01));
    s.inv = (den == 0.f) ? 0.f : __fdiv_rn(1.f, den);
    s.z0 = fc[2]; s.z1 = fc[5]; s.z2 = fc[8];
    return s;
}

// barycentric_weight (:79-109) + inside test (:144) + depth (:148) for pixel (px,py); no contraction.
__device__ __forceinline__ bool eval0(const Setup0& s, int px, int py, float& w0, float& w1, float& w2, float& zp) {
    const float v2x = __fsub_rn(static_cast<float>(px), s.x0), v2y = __fsub_rn(static_cast<float>(py), s.y0);
    const float d02 = __fadd_rn(__fmul_rn(s.v0x, v2x), __fmul_rn(s.v0y, v2y));
    const float d12 = __fadd_rn(__fmul_rn(s.v1x, v2x), __fmul_rn(s.v1y, v2y));
    const float u = __fmul_rn(__fsub_rn(__fmul_rn(s.d11, d02), __fmul_rn(s.d01, d12)), s.inv);
    const float v = __fmul_rn(__fsub_rn(__fmul_rn(s.d00, d12), __fmul_rn(s.d01, d02)), s.inv);
    w0 = __fsub_rn(__fsub_rn(1.f, u), v);
    w1 = v;
    w2 = u;
    if (!(w2 >= 0.f && w1 >= 0.f && w0 > 0.f)) return false;
    const float t = __fadd_rn(__fadd_rn(__fdiv_rn(w0, s.z0), __fdiv_rn(w1, s.z1)), __fdiv_rn(w2, s.z2));
    // The reference divides in double ('1.' is a double literal, :148) and rounds the quotient to float.  For a quotient of
    // two binary32 numbers, rounding first to binary64 (53 bits >= 2*24 + 2) and then to binary32 gives the correctly rounded
    // binary32 quotient (double rounding is innocuous for +,-,*,/ when the wide format has >= 2p+2 digits -- Figueroa), so
    // the IEEE single-precision division below returns the same bits without the fp64 division sequence; the bit-exactness
    // tests against the host build of the reference kernel cover it.
    zp = __fdiv_rn(1.f, t);
    return true;
}

// ---------------------------------------------------------------------------------------------- convention 1 (pytorch3d)
struct Setup1 {
    float x0, y0, z0, x1, y1, z1, x2, y2, z2, area, bx0, bx1, by0, by1;
};

__device__ __forceinline__ float pix_to_ndc(int i, int S) {   // -1 + (2 * i + 1.0f) / S
    return __fadd_rn(-1.f, __fdiv_rn(__fadd_rn(static_cast<float>(2 * i), 1.0f), static_cast<float>(S)));
}
// EdgeFunctionForward(p, v0, v1) = (p.x - v0.x) * (v1.y - v0.y) - (p.y - v0.y) * (v1.x - v0.x)
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)), __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}

// conservative pixel bbox of the NDC bbox (the exact float bbox test is part of eval1)
__device__ __forceinline__ bool bbox1(const float* __restrict__ fc, int w, int h, int& xmin, int& xmax, int& ymin, int& ymax) {
    const float x0 = fc[0], y0 = fc[1], z0 = fc[2], x1 = fc[3], y1 = fc[4], z1 = fc[5], x2 = fc[6], y2 = fc[7], z2 = fc[8];
    if (fmaxf(z0, fmaxf(z1, z2)) < 0.f) return false;                       // face behind the camera
    const float area = edge_fn(x0, y0, x1, y1, x2, y2);                      // EdgeFunctionForward(v0, v1, v2)
    if (area <= kP3dEps && area >= -kP3dEps) return false;                   // zero_face_area
    const float fx0 = fminf(x0, fminf(x1, x2)), fx1 = fmaxf(x0, fmaxf(x1, x2));
    const float fy0 = fminf(y0, fminf(y1, y2)), fy1 = fmaxf(y0, fmaxf(y1, y2));
    if (!(fx0 <= fx1 && fy0 <= fy1)) return false;                           // NaN vertices
    // column j = W-1-xi has ndc = -1 + (2j+1)/W  <=>  j = ((ndc+1) W - 1) / 2
    const float jx0 = floorf((fx0 + 1.f) * w * 0.5f - 0.5f) - 1.f, jx1 = ceilf((fx1 + 1.f) * w * 0.5f - 0.5f) + 1.f;
    const float jy0 = floorf((fy0 + 1.f) * h * 0.5f - 0.5f) - 1.f, jy1 = ceilf((fy1 + 1.f) * h * 0.5f - 0.5f) + 1.f;
    const int ja = static_cast<int>(fmaxf(jx0, 0.f)), jb = static_cast<int>(fminf(jx1, static_cast<float>(w - 1)));
    const int ka = static_cast<int>(fmaxf(jy0, 0.f)), kb = static_cast<int>(fminf(jy1, static_cast<float>(h - 1)));
    if (jx1 < 0.f || jy1 < 0.f || jx0 > static_cast<float>(w - 1) || jy0 > static_cast<float>(h - 1)) return false;
    xmin = w - 1 - jb; xmax = w - 1 - ja; ymin = h - 1 - kb; ymax = h - 1 - ka;
    return xmin <= xmax && ymin <= ymax;
}

__device__ __forceinline__ Setup1 setup1(const float* __restrict__ fc) {
    Setup1 s;
    s.x0 = fc[0]; s.y0 = fc[1]; s.z0 = fc[2]; s.x1 = fc[3]; s.y1 = fc[4]; s.z1 = fc[5]; s.x2 = fc[6]; s.y2 = fc[7]; s.z2 = fc[8];
    s.area = __fadd_rn(edge_fn(s.x2, s.y2, s.x0, s.y0, s.x1, s.y1), kP3dEps);     // BarycentricCoordsForward: area + kEpsilon
    s.bx0 = fminf(s.x0, fminf(s.x1, s.x2)); s.bx1 = fmaxf(s.x0, fmaxf(s.x1, s.x2));
    s.by0 = fminf(s.y0, fminf(s.y1, s.y2)); s.by1 = fmaxf(s.y0, fmaxf(s.y1, s.y2));
    return s;
}

__device__ __forceinline__ bool eval1(const Setup1& s, int px, int py, int w, int h, float& w0, float& w1, float& w2, float& zp) {
    const float xf = pix_to_ndc(w - 1 - px, w), yf = pix_to_ndc(h - 1 - py, h);
    if (xf > s.bx1 || xf < s.bx0 || yf > s.by1 || yf < s.by0) return false;     // CheckPointOutsideBoundingBox, blur 0
    w0 = __fdiv_rn(edge_fn(xf, yf, s.x1, s.y1, s.x2, s.y2), s.area);
    w1 = __fdiv_rn(edge_fn(xf, yf, s.x2, s.y2, s.x0, s.y0), s.area);
    w2 = __fdiv_rn(edge_fn(xf, yf, s.x0, s.y0, s.x1, s.y1), s.area);
    const float pz = __fadd_rn(__fadd_rn(__fmul_rn(w0, s.z0), __fmul_rn(w1, s.z1)), __fmul_rn(w2, s.z2));
    if (pz < 0.f) return false;                                                  // behind the image plane
    if (!(w0 > 0.f && w1 > 0.f && w2 > 0.f)) return false;                       // blur_radius 0: strictly inside
    zp = pz;
    return true;
}

template <int CONV>
__device__ __forceinline__ bool tri_bbox(const float* __restrict__ fc, int w, int h, int& xmin, int& xmax, int& ymin, int& ymax) {
    if (CONV == 0) return bbox0(fc, w, h, xmin, xmax, ymin, ymax);
    return bbox1(fc, w, h, xmin, xmax, ymin, ymax);
}

// ---------------------------------------------------------------------------------------------- binning
// pass 1 (fill == 0): count[b][bin] += 1 ; pass 2 (fill == 1): list[offset[b][bin] + cursor++] = f
template <int FILL, int CONV>
__global__ void __launch_bounds__(256) bin_kernel(const float* __restrict__ fv, RasterGeom g, int* __restrict__ count,
                                                  const int* __restrict__ offset, int* __restrict__ list,
                                                  int capacity, int* __restrict__ overflow) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(g.B) * g.F) return;
    const int b = static_cast<int>(i / g.F), f = static_cast<int>(i % g.F);
    int xmin, xmax, ymin, ymax;
    if (!tri_bbox<CONV>(fv + i * 9, g.w, g.h, xmin, xmax, ymin, ymax)) return;
    const int bx0 = xmin / TILE, bx1 = xmax / TILE, by0 = ymin / TILE, by1 = ymax / TILE;
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
// live in list[b*capacity, (b+1)*capacity).  The pass also re-zeroes `count`, which becomes the fill cursor and, after the
// fill pass, the list length again.
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

// ---------------------------------------------------------------------------------------------- tile kernel
// order-preserving float -> uint (IEEE total order on non-NaN values; -0 is canonicalised to +0 by the caller)
__device__ __forceinline__ unsigned int depth_bits(float z) {
    const unsigned int u = __float_as_uint(z);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long make_key(float z, int f) {
    return (static_cast<unsigned long long>(depth_bits(__fadd_rn(z, 0.f))) << 32) | static_cast<unsigned int>(f);
}

template <int CONV>
__device__ __forceinline__ void interpolate3(const float* __restrict__ c, float b0, float b1, float b2, float* __restrict__ o) {
#pragma unroll
    for (int k = 0; k < 3; ++k)   // :225, left-to-right, no contraction
        o[k] = __fadd_rn(__fadd_rn(__fmul_rn(b0, c[k]), __fmul_rn(b1, c[3 + k])), __fmul_rn(b2, c[6 + k]));
}

// per-face record staged in shared memory for the flattened fragment loop (16 floats: Setup0 has 13, Setup1 has 14)
template <int CONV> struct SetupOf;
template <> struct SetupOf<0> { using type = Setup0; };
template <> struct SetupOf<1> { using type = Setup1; };

template <int CONV>
__device__ __forceinline__ bool eval_any(const typename SetupOf<CONV>::type& s, int x, int y, int w, int h, float& w0, float& w1,
                                         float& w2, float& zp);
template <>
__device__ __forceinline__ bool eval_any<0>(const Setup0& s, int x, int y, int, int, float& w0, float& w1, float& w2, float& zp) {
    return eval0(s, x, y, w0, w1, w2, zp);
}
template <>
__device__ __forceinline__ bool eval_any<1>(const Setup1& s, int x, int y, int w, int h, float& w0, float& w1, float& w2, float& zp) {
    return eval1(s, x, y, w, h, w0, w1, w2, zp);
}
template <int CONV>
__device__ __forceinline__ typename SetupOf<CONV>::type setup_any(const float* __restrict__ fc);
template <> __device__ __forceinline__ Setup0 setup_any<0>(const float* __restrict__ fc) { return setup0(fc); }
template <> __device__ __forceinline__ Setup1 setup_any<1>(const float* __restrict__ fc) { return setup1(fc); }

// One CTA (256 threads) per 16x16 tile.  overflow[b] != 0: this image's lists did not fit -> every CTA of the image walks
// all F triangles (still exact).
//
// Phase 1 is FLATTENED: the tile's list is taken 256 triangles at a time; thread t sets up triangle t once (face test,
// bbox clipped to the tile, edge constants) into shared memory together with its fragment count (clipped bbox area), a
// CTA-wide prefix sum turns the counts into one index space, and every thread then takes fragments w = t, t+256, ...:
// a binary search in the prefix array gives the triangle, the remainder the pixel.  All lanes stay busy whatever the
// triangle sizes are (the first version gave each triangle 4 lanes that walked its bbox: ncu showed 3-4 of 32 lanes
// active in the fragment loop and 45% of all warp samples waiting at the barrier behind the few busy warps).
template <int CONV>
__global__ void __launch_bounds__(256, 8) raster_tile_kernel(const float* __restrict__ fv, const float* __restrict__ colors,
                                                          const float* __restrict__ colors2, float* __restrict__ depth,
                                                          int* __restrict__ tri, float* __restrict__ out3,
                                                          float* __restrict__ out3b, RasterGeom g,
                                                          const int* __restrict__ count, const int* __restrict__ offset,
                                                          const int* __restrict__ list, const int* __restrict__ overflow) {
    using SetupT = typename SetupOf<CONV>::type;
    __shared__ unsigned long long key[TILE * TILE];
    __shared__ SetupT s_setup[256];
    __shared__ int s_incl[256];                 // inclusive prefix of the fragment counts
    __shared__ int s_face[256];
    __shared__ unsigned int s_box[256];         // xa | ya << 4 | (width-1) << 8 | ceil(4096/width) << 12  (tile-relative)
    __shared__ unsigned char s_coarse[2048];    // record holding fragment 32*k: entry point of the per-fragment search
    __shared__ int s_warp[8];
    const int bx = blockIdx.x, by = blockIdx.y, b = blockIdx.z;      // 3-D grid: no per-thread integer divisions
    const int gbin = (b * g.bins_y + by) * g.bins_x + bx;
    const bool brute = overflow[b] != 0;
    const int n = brute ? g.F : count[gbin];
    if (n == 0) return;                                              // empty tile (uniform per CTA): buffers keep their values
    const int tx0 = bx * TILE, ty0 = by * TILE;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int px = tx0 + (t & (TILE - 1)), py = ty0 + (t >> 4);
    const bool in_img = px < g.w && py < g.h;
    const long long pix = (static_cast<long long>(b) * g.h + py) * g.w + px;
    // phase 0: the tile's depth buffer, initialised from the caller's (a fragment must beat or tie it)
    {
        unsigned long long k0 = 0ull;                    // out-of-image / NaN depth: nothing can win
        if (in_img) {
            const float d0 = depth[pix];
            if (d0 == d0) k0 = make_key(d0, kNoFace);
        }
        key[t] = k0;
    }
    // phase 1: triangles -> fragments -> shared-memory depth test
    const int* lst = list + (brute ? 0 : offset[gbin]);
    const float* fvb = fv + static_cast<long long>(b) * g.F * 9;
    for (int base = 0; base < n; base += 256) {
        __syncthreads();                                 // key[] initialised / previous chunk's records consumed
        int cnt = 0;
        const int i = base + t;
        if (i < n) {
            const int f = brute ? i : lst[i];
            const float* fc = fvb + static_cast<long long>(f) * 9;
            int xmin, xmax, ymin, ymax;
            if (tri_bbox<CONV>(fc, g.w, g.h, xmin, xmax, ymin, ymax)) {
                const int xa = max(xmin, tx0), xb = min(xmax, tx0 + TILE - 1), ya = max(ymin, ty0), yb = min(ymax, ty0 + TILE - 1);
                if (xa <= xb && ya <= yb) {
                    s_setup[t] = setup_any<CONV>(fc);
                    s_face[t] = f;
                    const unsigned int wd = static_cast<unsigned int>(xb - xa + 1);          // 1..16
                    s_box[t] = static_cast<unsigned int>(xa - tx0) | (static_cast<unsigned int>(ya - ty0) << 4) | ((wd - 1u) << 8) |
                               (((4096u + wd - 1u) / wd) << 12);
                    cnt = static_cast<int>(wd) * (yb - ya + 1);
                }
            }
        }
        int incl = cnt;                                   // CTA-wide inclusive scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int v = s_warp[k];
            if (k < wid) woff += v;
            total += v;
        }
        incl += woff;
        s_incl[t] = incl;
        // coarse index: the record that holds fragment 32*k (records are at most 256 fragments: <= 8 entries each, usually 0-1)
        for (int k = (incl - cnt + 31) >> 5; (k << 5) < incl; ++k) s_coarse[k] = static_cast<unsigned char>(t);
        __syncthreads();
        for (int wk = t; wk < total; wk += 256) {
            int r = s_coarse[wk >> 5];                    // first record whose inclusive prefix exceeds wk: a short walk
            int prev = r ? s_incl[r - 1] : 0, cur = s_incl[r];
            while (cur <= wk) { prev = cur; cur = s_incl[++r]; }
            const unsigned int box = s_box[r];
            const int wd = static_cast<int>((box >> 8) & 15u) + 1;
            const int local = wk - prev;                  // < 256
            // local / wd without a division: (local * ceil(4096/wd)) >> 12 is exact for local < 256, wd <= 16
            const int ry = static_cast<int>((static_cast<unsigned int>(local) * (box >> 12)) >> 12), rx = local - ry * wd;
            const int lx = static_cast<int>(box & 15u) + rx, ly = static_cast<int>((box >> 4) & 15u) + ry;
            float w0, w1, w2, zp;
            if (!eval_any<CONV>(s_setup[r], tx0 + lx, ty0 + ly, g.w, g.h, w0, w1, w2, zp) || !(zp == zp)) continue;
            const unsigned long long k = make_key(zp, s_face[r]);
            unsigned long long* slot = &key[ly * TILE + lx];
            if (k < *reinterpret_cast<volatile unsigned long long*>(slot)) atomicMin(slot, k);
        }
    }
    __syncthreads();
    // phase 2: one pixel per thread -- re-evaluate the winner (same arithmetic, same values) and write once, coalesced
    if (!in_img) return;
    const int f = static_cast<int>(static_cast<unsigned int>(key[t] & 0xffffffffull));
    if (f == kNoFace || key[t] == 0ull) return;
    const float* fc = fvb + static_cast<long long>(f) * 9;
    float w0, w1, w2, zp;
    {
        const SetupT sw = setup_any<CONV>(fc);
        eval_any<CONV>(sw, px, py, g.w, g.h, w0, w1, w2, zp);
    }
    depth[pix] = zp;
    tri[pix] = f;
    if (out3) {
        float* o = out3 + pix * 3;
        if (colors) interpolate3<CONV>(colors + (static_cast<long long>(b) * g.F + f) * 9, w0, w1, w2, o);
        else { o[0] = w0; o[1] = w1; o[2] = w2; }
    }
    if (out3b && colors2) interpolate3<CONV>(colors2 + (static_cast<long long>(b) * g.F + f) * 9, w0, w1, w2, out3b + pix * 3);
}

// ---------------------------------------------------------------------------------------------- backward
// One thread per triangle gathers the pixels it owns (triangle buffer == its index) over its bbox.  The barycentrics of a
// triangle are AFFINE in the pixel position q, so every term of the chain rule is linear in the moments
//     M[c][m] = sum_{owned pixels} g_c(pixel) * (1, q.x, q.y)[m]
// of the upstream gradients g_c (3 barycentric / 3 + 3 interpolated-attribute channels, and the depth term).  The pixel loop
// therefore only accumulates moments (a few FMAs per owned pixel: lanes of a warp hit their owned pixels at different
// iterations, so whatever sits inside the loop runs at a few active lanes), and the per-triangle algebra -- reciprocals,
// the gradients of the dot products / edge functions -- runs ONCE per triangle after the loop at full lane occupancy.
// No atomics, outputs are written (not accumulated), summation order is fixed: deterministic.
struct Moments {
    float b[3][3];      // upstream gradient of the barycentric buffer (or zero)
    float c1[3][3];     // upstream gradient of attribute image 1, channel k
    float c2[3][3];     // upstream gradient of attribute image 2
    float z[3];         // depth term: sum gz * (1, q.x, q.y), gz = upstream depth gradient times d(depth)/d(s) (conv. 0) or 1
};

template <int CONV>
__global__ void __launch_bounds__(128, 8) raster_bwd_face_kernel(const float* __restrict__ fv, const float* __restrict__ colors,
                                                              const float* __restrict__ colors2, const int* __restrict__ tri,
                                                              const float* __restrict__ g_bary, const float* __restrict__ g_img,
                                                              const float* __restrict__ g_img2, const float* __restrict__ g_depth,
                                                              float* __restrict__ g_fv, float* __restrict__ g_col,
                                                              float* __restrict__ g_col2, int B, int F, int h, int w) {
    // kBwdLanes consecutive lanes share one triangle and interleave its bbox pixels: a warp's loop length is the LARGEST
    // bbox among its triangles (each iteration is a dependent triangle-buffer read), so eight lanes per triangle cut the
    // latency chain 8x; the moments are then summed over the lane group with shuffles (fixed order: still deterministic).
    const long long gtid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i = gtid / kBwdLanes;
    const int sub = static_cast<int>(gtid % kBwdLanes);
    const bool live = i < static_cast<long long>(B) * F;
    const long long ic = live ? i : 0;                  // dead lanes of the last warp shadow triangle 0 (no writes)
    const int b = static_cast<int>(ic / F), f = static_cast<int>(ic - static_cast<long long>(b) * F);
    const long long fo = ic * 9;
    const float* fc = fv + fo;
    const float x0 = fc[0], y0 = fc[1], z0 = fc[2], x1 = fc[3], y1 = fc[4], z1 = fc[5], x2 = fc[6], y2 = fc[7], z2 = fc[8];
    Moments M;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        M.z[a] = 0.f;
#pragma unroll
        for (int m = 0; m < 3; ++m) { M.b[a][m] = 0.f; M.c1[a][m] = 0.f; M.c2[a][m] = 0.f; }
    }
    const bool use1 = g_img && colors, use2 = g_img2 && colors2;
    // affine form of the barycentrics: convention 0: u = Ux qx + Uy qy, v = Vx qx + Vy qy with q = p - p0
    //                                  convention 1: w_i = (k_i0 + k_ix qx + k_iy qy) / A with q = NDC pixel centre
    float Ux = 0.f, Uy = 0.f, Vx = 0.f, Vy = 0.f, inv = 0.f, d00 = 0.f, d01 = 0.f, d11 = 0.f;
    const float v0x = x2 - x0, v0y = y2 - y0, v1x = x1 - x0, v1y = y1 - y0;
    float A = 1.f;
    if (CONV == 0) {
        d00 = v0x * v0x + v0y * v0y; d01 = v0x * v1x + v0y * v1y; d11 = v1x * v1x + v1y * v1y;
        const float den = d00 * d11 - d01 * d01;
        inv = den == 0.f ? 0.f : 1.f / den;
        Ux = inv * (d11 * v0x - d01 * v1x); Uy = inv * (d11 * v0y - d01 * v1y);
        Vx = inv * (d00 * v1x - d01 * v0x); Vy = inv * (d00 * v1y - d01 * v0y);
    } else {
        A = (x2 - x0) * (y1 - y0) - (y2 - y0) * (x1 - x0) + kP3dEps;
    }
    int xmin, xmax, ymin, ymax;
    if (live && tri_bbox<CONV>(fc, w, h, xmin, xmax, ymin, ymax)) {
        const long long img = static_cast<long long>(b) * h * w;
        const int bw = xmax - xmin + 1, area = bw * (ymax - ymin + 1);
        // four triangle-buffer reads in flight per lane before the first ownership test (the loop is a chain of dependent
        // global round trips: ownership read -> gradient reads), same accumulation order as a plain loop
        for (int p0 = sub; p0 < area; p0 += kBwdLanes * 4) {
            int own[4];
            unsigned int xy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int pi = p0 + u * kBwdLanes;
                own[u] = -2;                                  // never a face index (faces >= 0, empty pixels -1)
                xy[u] = 0u;
                if (pi < area) {
                    const int ry = pi / bw;
                    const int y = ymin + ry, x = xmin + (pi - ry * bw);
                    xy[u] = static_cast<unsigned int>(x) | (static_cast<unsigned int>(y) << 16);   // h, w <= 65535
                    own[u] = tri[img + static_cast<long long>(y) * w + x];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (own[u] != f) continue;
                const int x = static_cast<int>(xy[u] & 0xffffu), y = static_cast<int>(xy[u] >> 16);
                const long long pix = img + static_cast<long long>(y) * w + x;
                float qx, qy;
                if (CONV == 0) { qx = x - x0; qy = y - y0; }
                else { qx = -1.f + (2 * (w - 1 - x) + 1.0f) / w; qy = -1.f + (2 * (h - 1 - y) + 1.0f) / h; }
                if (g_bary) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float g = g_bary[pix * 3 + a];
                        M.b[a][0] += g; M.b[a][1] += g * qx; M.b[a][2] += g * qy;
                    }
                }
                if (use1) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float g = g_img[pix * 3 + k];
                        M.c1[k][0] += g; M.c1[k][1] += g * qx; M.c1[k][2] += g * qy;
                    }
                }
                if (use2) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float g = g_img2[pix * 3 + k];
                        M.c2[k][0] += g; M.c2[k][1] += g * qx; M.c2[k][2] += g * qy;
                    }
                }
                if (g_depth) {
                    float gz = g_depth[pix];
                    if (CONV == 0) {          // zp = 1 / s, s = sum w_i / z_i: the factor d zp / d s = -zp^2 needs this pixel's weights
                        const float u = Ux * qx + Uy * qy, v = Vx * qx + Vy * qy;
                        const float sden = (1.f - u - v) / z0 + v / z1 + u / z2;
                        const float zp = 1.f / sden;
                        gz *= -zp * zp;
                    }
                    M.z[0] += gz; M.z[1] += gz * qx; M.z[2] += gz * qy;
                }
            }
        }
    }
    // ---- sum the moments over the lane group (xor butterflies within aligned groups of kBwdLanes lanes)
#pragma unroll
    for (int o = 1; o < kBwdLanes; o <<= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            M.z[a] += __shfl_xor_sync(0xffffffffu, M.z[a], o);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                M.b[a][m] += __shfl_xor_sync(0xffffffffu, M.b[a][m], o);
                M.c1[a][m] += __shfl_xor_sync(0xffffffffu, M.c1[a][m], o);
                M.c2[a][m] += __shfl_xor_sync(0xffffffffu, M.c2[a][m], o);
            }
        }
    }
    if (!live || sub != 0) return;
    // ---- per-triangle algebra.  Gw[i][m]: moments of the total upstream gradient of barycentric i
    float Gw[3][3];
    const float* c = colors ? colors + fo : nullptr;
    const float* c2 = colors2 ? colors2 + fo : nullptr;
    const float zi[3] = {z0, z1, z2};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            float t = M.b[a][m];
            if (use1) t += M.c1[0][m] * c[a * 3 + 0] + M.c1[1][m] * c[a * 3 + 1] + M.c1[2][m] * c[a * 3 + 2];
            if (use2) t += M.c2[0][m] * c2[a * 3 + 0] + M.c2[1][m] * c2[a * 3 + 1] + M.c2[2][m] * c2[a * 3 + 2];
            if (g_depth) t += (CONV == 0 ? M.z[m] / zi[a] : M.z[m] * zi[a]);     // d s / d w_i = 1/z_i   |   d pz / d w_i = z_i
            Gw[a][m] = t;
        }
    float gf[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) gf[k] = 0.f;
    float Wz[3] = {0.f, 0.f, 0.f};          // sum gz * w_i  (depth gradient w.r.t. the vertex depths)
    float gc1[9], gc2[9];
    if (CONV == 0) {
        // (w0, w1, w2) = (1-u-v, v, u): Su = moments of (gw2 - gw0), Sv = moments of (gw1 - gw0)
        float Su[3], Sv[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) { Su[m] = Gw[2][m] - Gw[0][m]; Sv[m] = Gw[1][m] - Gw[0][m]; }
        const float Au02 = v0x * Su[1] + v0y * Su[2], Au12 = v1x * Su[1] + v1y * Su[2];
        const float Av02 = v0x * Sv[1] + v0y * Sv[2], Av12 = v1x * Sv[1] + v1y * Sv[2];
        const float Gden = -((d11 * Au02 - d01 * Au12) + (d00 * Av12 - d01 * Av02)) * inv * inv;   // inv == 0 for den == 0
        const float g_d00 = inv * Av12 + Gden * d11;
        const float g_d11 = inv * Au02 + Gden * d00;
        const float g_d01 = -inv * Au12 - inv * Av02 - 2.f * Gden * d01;
        float Sd02[3], Sd12[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) { Sd02[m] = inv * (d11 * Su[m] - d01 * Sv[m]); Sd12[m] = inv * (-d01 * Su[m] + d00 * Sv[m]); }
        const float g_v0x = 2.f * g_d00 * v0x + g_d01 * v1x + Sd02[1], g_v0y = 2.f * g_d00 * v0y + g_d01 * v1y + Sd02[2];
        const float g_v1x = 2.f * g_d11 * v1x + g_d01 * v0x + Sd12[1], g_v1y = 2.f * g_d11 * v1y + g_d01 * v0y + Sd12[2];
        const float g_v2x = Sd02[0] * v0x + Sd12[0] * v1x, g_v2y = Sd02[0] * v0y + Sd12[0] * v1y;
        gf[0] = -(g_v0x + g_v1x + g_v2x); gf[1] = -(g_v0y + g_v1y + g_v2y);   // v0 = p2-p0, v1 = p1-p0, v2 = p-p0
        gf[3] = g_v1x; gf[4] = g_v1y; gf[6] = g_v0x; gf[7] = g_v0y;
        // sums of g * w_i with w2 = u = Ux qx + Uy qy, w1 = v = Vx qx + Vy qy, w0 = 1 - u - v
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float a2 = Ux * M.c1[k][1] + Uy * M.c1[k][2], a1 = Vx * M.c1[k][1] + Vy * M.c1[k][2];
            gc1[6 + k] = a2; gc1[3 + k] = a1; gc1[k] = M.c1[k][0] - a1 - a2;
            const float b2 = Ux * M.c2[k][1] + Uy * M.c2[k][2], b1 = Vx * M.c2[k][1] + Vy * M.c2[k][2];
            gc2[6 + k] = b2; gc2[3 + k] = b1; gc2[k] = M.c2[k][0] - b1 - b2;
        }
        Wz[2] = Ux * M.z[1] + Uy * M.z[2]; Wz[1] = Vx * M.z[1] + Vy * M.z[2]; Wz[0] = M.z[0] - Wz[1] - Wz[2];
        if (g_depth) { gf[2] = -Wz[0] / (z0 * z0); gf[5] = -Wz[1] / (z1 * z1); gf[8] = -Wz[2] / (z2 * z2); }   // d s / d z_i = -w_i / z_i^2
    } else {
        // e_0 = E(q,v1,v2), e_1 = E(q,v2,v0), e_2 = E(q,v0,v1), E(q,a,b) = (qx-ax)(by-ay) - (qy-ay)(bx-ax); w_i = e_i / A
        const float iA = 1.f / A;
        const float ax[3] = {x1, x2, x0}, ay[3] = {y1, y2, y0}, bx[3] = {x2, x0, x1}, by[3] = {y2, y0, y1};
        const int ia[3] = {1, 2, 0}, ib[3] = {2, 0, 1};       // vertex index of a and b for e_i
        float sum_ge = 0.f;                                    // sum_i sum_pix gw_i e_i
        float ki[3][3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const float dyb = by[e] - ay[e], dxb = bx[e] - ax[e];
            ki[e][0] = -ax[e] * dyb + ay[e] * dxb; ki[e][1] = dyb; ki[e][2] = -dxb;        // e = k0 + k1 qx + k2 qy
            sum_ge += ki[e][0] * Gw[e][0] + ki[e][1] * Gw[e][1] + ki[e][2] * Gw[e][2];
            // dE/da = (q.y - b.y, b.x - q.x), dE/db = (-(q.y - a.y), q.x - a.x), summed over the pixels with weight gw_e / A
            const float g0 = Gw[e][0] * iA, gx_ = Gw[e][1] * iA, gy_ = Gw[e][2] * iA;
            gf[ia[e] * 3 + 0] += gy_ - by[e] * g0; gf[ia[e] * 3 + 1] += bx[e] * g0 - gx_;
            gf[ib[e] * 3 + 0] += -(gy_ - ay[e] * g0); gf[ib[e] * 3 + 1] += gx_ - ax[e] * g0;
        }
        const float gA = -sum_ge * iA * iA;
        // A = E(v2, v0, v1) + eps: "q" = v2, a = v0, b = v1
        gf[6] += gA * (y1 - y0); gf[7] += gA * -(x1 - x0);
        gf[0] += gA * (y2 - y1); gf[1] += gA * (x1 - x2);
        gf[3] += gA * -(y2 - y0); gf[4] += gA * (x2 - x0);
#pragma unroll
        for (int e = 0; e < 3; ++e) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                gc1[e * 3 + k] = (ki[e][0] * M.c1[k][0] + ki[e][1] * M.c1[k][1] + ki[e][2] * M.c1[k][2]) * iA;
                gc2[e * 3 + k] = (ki[e][0] * M.c2[k][0] + ki[e][1] * M.c2[k][1] + ki[e][2] * M.c2[k][2]) * iA;
            }
            Wz[e] = (ki[e][0] * M.z[0] + ki[e][1] * M.z[1] + ki[e][2] * M.z[2]) * iA;
        }
        if (g_depth) { gf[2] = Wz[0]; gf[5] = Wz[1]; gf[8] = Wz[2]; }          // pz = sum w_i z_i
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) g_fv[fo + k] = gf[k];
    if (g_col)
#pragma unroll
        for (int k = 0; k < 9; ++k) g_col[fo + k] = use1 ? gc1[k] : 0.f;
    if (g_col2)
#pragma unroll
        for (int k = 0; k < 9; ++k) g_col2[fo + k] = use2 ? gc2[k] : 0.f;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace gifb200

using namespace gifb200;

// workspace layout: [count: B*nbins ints][offset: B*nbins ints][overflow flags: B ints (padded)][list: B*capacity ints]
static int list_capacity(int F) {   // list slots per image; a FLAME render at 256^2 needs ~1 F at 16x16 tiles, allow 8 F
    const long long cap = static_cast<long long>(F) * 8 + 1024;
    return static_cast<int>(cap > 0x3fffffffLL ? 0x3fffffffLL : cap);
}

extern "C" size_t gifb200_rasterize_workspace_bytes(int B, int F, int h, int w) {
    if (B <= 0 || h <= 0 || w <= 0 || F < 0) return 0;
    const size_t nb = static_cast<size_t>(B) * ((h + TILE - 1) / TILE) * ((w + TILE - 1) / TILE);
    return align_up(nb * 4, 256) * 2 + align_up(static_cast<size_t>(B) * 4, 256) + static_cast<size_t>(list_capacity(F)) * 4 * B;
}

template <int CONV>
static int rasterize_fwd_impl(const float* face_vertices, const float* face_colors, const float* face_colors2, float* depth,
                              int32_t* triangle, float* out3, float* out3b, int B, int F, int h, int w, void* workspace,
                              size_t workspace_bytes, cudaStream_t st) {
    RasterGeom g;
    g.B = B; g.F = F; g.h = h; g.w = w;
    g.bins_x = (w + TILE - 1) / TILE; g.bins_y = (h + TILE - 1) / TILE; g.nbins = g.bins_x * g.bins_y;
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
    bin_kernel<0, CONV><<<cdiv(ntri, 256), 256, 0, st>>>(face_vertices, g, count, offset, list, capacity, overflow);
    GIFB200_LAUNCH_CHECK("bin_kernel<count>");
    bin_scan_kernel<<<B, 1024, 0, st>>>(count, offset, g.nbins, capacity, overflow);
    GIFB200_LAUNCH_CHECK("bin_scan_kernel");
    bin_kernel<1, CONV><<<cdiv(ntri, 256), 256, 0, st>>>(face_vertices, g, count, offset, list, capacity, overflow);
    GIFB200_LAUNCH_CHECK("bin_kernel<fill>");
    GIFB200_REQUIRE(g.bins_y <= 65535 && B <= 65535, GIFB200_E_SHAPE, "rasterize: grid too large");
    // The tile kernel is latency-bound (a CTA's life is a chain of ~6 dependent global round trips: list -> vertices -> ... ->
    // winner's attributes): 8 CTAs of 32 registers per SM instead of 5, with the shared-memory carveout that lets them fit
    static bool carveout_set = false;
    if (!carveout_set) {
        cudaFuncSetAttribute(raster_tile_kernel<CONV>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carveout_set = true;
    }
    raster_tile_kernel<CONV><<<dim3(g.bins_x, g.bins_y, B), 256, 0, st>>>(face_vertices, face_colors, face_colors2, depth,
                                                                          triangle, out3, out3b, g, count, offset, list, overflow);
    GIFB200_LAUNCH_CHECK("raster_tile_kernel");
    return GIFB200_OK;
}

extern "C" int gifb200_rasterize_fwd_ex(const float* face_vertices, const float* face_colors, const float* face_colors2,
                                        float* depth, int32_t* triangle, float* out3, float* out3b, int B, int F, int h, int w,
                                        int convention, void* workspace, size_t workspace_bytes, gifb200_stream_t stream) {
    GIFB200_REQUIRE(B >= 0 && F >= 0 && h > 0 && w > 0, GIFB200_E_SHAPE, "rasterize: bad shape");
    GIFB200_REQUIRE(convention == 0 || convention == 1, GIFB200_E_SHAPE, "rasterize: convention must be 0 (standard_rasterize) or 1 (pytorch3d)");
    GIFB200_REQUIRE(!face_colors2 || (face_colors && out3b), GIFB200_E_SHAPE, "rasterize: a second attribute set needs the first one and out3b");
    if (B == 0 || F == 0) return GIFB200_OK;     // nothing can be written: buffers keep the caller's initial values
    GIFB200_REQUIRE(workspace && workspace_bytes >= gifb200_rasterize_workspace_bytes(B, F, h, w), GIFB200_E_WORKSPACE,
                    "rasterize: workspace too small (see gifb200_rasterize_workspace_bytes)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (convention == 0)
        return rasterize_fwd_impl<0>(face_vertices, face_colors, face_colors2, depth, triangle, out3, out3b, B, F, h, w, workspace, workspace_bytes, st);
    return rasterize_fwd_impl<1>(face_vertices, face_colors, face_colors2, depth, triangle, out3, out3b, B, F, h, w, workspace, workspace_bytes, st);
}

extern "C" int gifb200_rasterize_fwd(const float* face_vertices, const float* face_colors, float* depth,
                                     int32_t* triangle, float* out3, int B, int F, int h, int w, void* workspace,
                                     size_t workspace_bytes, gifb200_stream_t stream) {
    return gifb200_rasterize_fwd_ex(face_vertices, face_colors, nullptr, depth, triangle, out3, nullptr, B, F, h, w, 0, workspace,
                                    workspace_bytes, stream);
}

extern "C" int gifb200_rasterize_bwd_ex(const float* face_vertices, const float* face_colors, const float* face_colors2,
                                        const int32_t* triangle, const float* g_bary, const float* g_img, const float* g_img2,
                                        const float* g_depth, float* g_face_vertices, float* g_face_colors,
                                        float* g_face_colors2, int B, int F, int h, int w, int convention,
                                        gifb200_stream_t stream) {
    GIFB200_REQUIRE(B >= 0 && F >= 0 && h > 0 && w > 0 && h <= 65535 && w <= 65535, GIFB200_E_SHAPE, "rasterize_bwd: bad shape");
    GIFB200_REQUIRE(convention == 0 || convention == 1, GIFB200_E_SHAPE, "rasterize_bwd: convention must be 0 or 1");
    GIFB200_REQUIRE(g_face_vertices != nullptr, GIFB200_E_SHAPE, "rasterize_bwd: g_face_vertices is required");
    GIFB200_REQUIRE(!g_img || face_colors, GIFB200_E_SHAPE, "rasterize_bwd: g_img needs face_colors");
    GIFB200_REQUIRE(!g_img2 || face_colors2, GIFB200_E_SHAPE, "rasterize_bwd: g_img2 needs face_colors2");
    if (B == 0 || F == 0) return GIFB200_OK;
    const long long ntri = static_cast<long long>(B) * F * kBwdLanes;       // threads: kBwdLanes lanes per triangle
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (convention == 0)
        raster_bwd_face_kernel<0><<<cdiv(ntri, 128), 128, 0, st>>>(face_vertices, face_colors, face_colors2, triangle, g_bary, g_img,
                                                                  g_img2, g_depth, g_face_vertices, g_face_colors, g_face_colors2, B, F, h, w);
    else
        raster_bwd_face_kernel<1><<<cdiv(ntri, 128), 128, 0, st>>>(face_vertices, face_colors, face_colors2, triangle, g_bary, g_img,
                                                                  g_img2, g_depth, g_face_vertices, g_face_colors, g_face_colors2, B, F, h, w);
    GIFB200_LAUNCH_CHECK("raster_bwd_face_kernel");
    return GIFB200_OK;
}

extern "C" int gifb200_rasterize_bwd(const float* face_vertices, const float* face_colors, const int32_t* triangle,
                                     const float* g_bary, const float* g_img, const float* g_depth,
                                     float* g_face_vertices, float* g_face_colors, int B, int F, int h, int w,
                                     gifb200_stream_t stream) {
    return gifb200_rasterize_bwd_ex(face_vertices, face_colors, nullptr, triangle, g_bary, g_img, nullptr, g_depth, g_face_vertices,
                                    g_face_colors, nullptr, B, F, h, w, 0, stream);
}
