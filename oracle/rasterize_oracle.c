/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, single thread) of the reference's z-buffer triangle
 * rasteriser, my_utils/standard_rasterize_cuda/standard_rasterize_cuda_kernel.cu.  It is the oracle for
 * gif_b200/csrc/rasterize.cu; only tests/, smoke() and bench.py's cpu_baseline may call it.
 *
 * Fixed here where the reference leaves behaviour to the compiler / the scheduler:
 *   - every fp32 product and sum is rounded separately (no FMA contraction): build with -ffp-contract=off;
 *   - exact ties in zp at one pixel are won by the LOWEST face index (the reference's winner is a race);
 *   - the depth test is the net effect of the reference's atomicMin + "if depth == zp" + second identical launch
 *     (:150-160,:252-269): a fragment replaces the pixel iff zp < depth, or zp == depth (incl. the caller's
 *     initial depth value) and it has the lowest index among such fragments.
 * Pinned against the unmodified reference kernels executed on the CPU (oracle/_ref, tests/test_raster_oracle.py)
 * and against the reference's checked-in golden mesh data/obj/body_vis.obj (tests/golden/body_visibility.npz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { float w0, w1, w2; } bary_t;

/* :79-109  barycentric_weight(); p = (x,y) integer pixel centre. (the unused s = p.dot(p) of :83 is omitted) */
static bary_t bary_weights(float px, float py, float x0, float y0, float x1, float y1, float x2, float y2) {
    float v0x = x2 - x0, v0y = y2 - y0;   /* v0 = p2 - p0 */
    float v1x = x1 - x0, v1y = y1 - y0;   /* v1 = p1 - p0 */
    float v2x = px - x0, v2y = py - y0;   /* v2 = p  - p0 */
    float d00 = v0x * v0x + v0y * v0y;
    float d01 = v0x * v1x + v0y * v1y;
    float d02 = v0x * v2x + v0y * v2y;
    float d11 = v1x * v1x + v1y * v1y;
    float d12 = v1x * v2x + v1y * v2y;
    float den = d00 * d11 - d01 * d01;
    float inv = (den == 0.0f) ? 0.0f : 1.0f / den;
    float u = (d11 * d02 - d01 * d12) * inv;
    float v = (d00 * d12 - d01 * d02) * inv;
    bary_t r;
    r.w0 = 1.0f - u - v;
    r.w1 = v;
    r.w2 = u;
    return r;
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* colors == NULL: standard_rasterize (:112-167) -> out3 = barycentric weights (B,h,w,3)
 * colors != NULL: standard_rasterize_colors (:171-233) -> out3 = interpolated colours (B,h,w,3)
 * owner: scratch (B*h*w) ints, internal tie-break bookkeeping. */
static void rasterize(const float* fv, const float* colors, float* depth, int32_t* tri, float* out3,
                      int batch, int ntri, int h, int w) {
    size_t npix = (size_t)batch * h * w;
    int32_t* owner = (int32_t*)malloc(npix * sizeof(int32_t));
    for (size_t i = 0; i < npix; ++i) owner[i] = INT32_MAX;      /* no fragment written by this call yet */
    for (int b = 0; b < batch; ++b)
        for (int f = 0; f < ntri; ++f) {
            const float* fc = fv + ((size_t)b * ntri + f) * 9;
            float x0 = fc[0], y0 = fc[1], z0 = fc[2];
            float x1 = fc[3], y1 = fc[4], z1 = fc[5];
            float x2 = fc[6], y2 = fc[7], z2 = fc[8];
            /* :32-34 check_face_frontside */
            if (!((y2 - y0) * (x1 - x0) < (y1 - y0) * (x2 - x0))) continue;
            /* :133-136 clamped bbox */
            int xmin = imax((int)ceilf(fminf(x0, fminf(x1, x2))), 0);
            int xmax = imin((int)floorf(fmaxf(x0, fmaxf(x1, x2))), w - 1);
            int ymin = imax((int)ceilf(fminf(y0, fminf(y1, y2))), 0);
            int ymax = imin((int)floorf(fmaxf(y0, fmaxf(y1, y2))), h - 1);
            for (int y = ymin; y <= ymax; ++y)
                for (int x = xmin; x <= xmax; ++x) {
                    bary_t bw = bary_weights((float)x, (float)y, x0, y0, x1, y1, x2, y2);
                    if (!(bw.w2 >= 0.0f && bw.w1 >= 0.0f && bw.w0 > 0.0f)) continue;        /* :144 */
                    float s = bw.w0 / z0 + bw.w1 / z1 + bw.w2 / z2;
                    float zp = (float)(1.0 / (double)s);                                      /* :148 ('1.' is double) */
                    size_t pix = ((size_t)b * h + y) * w + x;
                    int take = (zp < depth[pix]) || (zp == depth[pix] && f < owner[pix]);
                    if (!take) continue;
                    depth[pix] = zp;
                    owner[pix] = f;
                    tri[pix] = f;                                                             /* :151 */
                    if (colors) {
                        const float* c = colors + ((size_t)b * ntri + f) * 9;
                        for (int k = 0; k < 3; ++k)                                           /* :225 */
                            out3[pix * 3 + k] = bw.w0 * c[k] + bw.w1 * c[3 + k] + bw.w2 * c[6 + k];
                    } else {
                        out3[pix * 3 + 0] = bw.w0;                                            /* :152-154 */
                        out3[pix * 3 + 1] = bw.w1;
                        out3[pix * 3 + 2] = bw.w2;
                    }
                }
        }
    free(owner);
}

void oracle_standard_rasterize(const float* face_vertices, float* depth, int32_t* tri, float* bary,
                               int batch, int ntri, int h, int w) {
    rasterize(face_vertices, NULL, depth, tri, bary, batch, ntri, h, w);
}

void oracle_standard_rasterize_colors(const float* face_vertices, const float* face_colors, float* depth,
                                      int32_t* tri, float* images, int batch, int ntri, int h, int w) {
    rasterize(face_vertices, face_colors, depth, tri, images, batch, ntri, h, w);
}

/* ---------------------------------------------------------------------------------------------------------------
 * pytorch3d `rasterize_meshes` as the reference calls it -- the rasteriser its conditioning maps are made with
 * (my_utils/photometric_optimization/renderer.py:35-67: image_size S, blur_radius 0, faces_per_pixel 1, bin_size None,
 * perspective_correct False; the caller negates x and y first, renderer.py:55).
 *
 * PARITY UNPINNED: pytorch3d is a third-party dependency, the reference pins an unversioned fork
 * (requirements.txt:36: git+https://github.com/ParthaEth/pytorch3d.git, no commit/tag; submodule README: PyTorch3D 0.2) that
 * is absent from /root/reference, and the reference holds no test or golden output at this boundary.  What
 * follows restates the PUBLISHED algorithm of pytorch3d 0.2's naive rasteriser (csrc/rasterize_meshes/rasterize_meshes.cu:
 * RasterizeMeshesNaiveCudaKernel / CheckPixelInsideFace; csrc/rasterize_meshes/geometry_utils.cuh: EdgeFunctionForward,
 * BarycentricCoordsForward, CheckPointOutsideBoundingBox; rasterization_utils.cuh: PixToNdc); the coarse-to-fine path the
 * heuristic bin_size selects gives the same result as long as no bin overflows.  Rules:
 *   - NDC axes +X left, +Y up: output pixel (yi, xi) samples p = (PixToNdc(W-1-xi, W), PixToNdc(H-1-yi, H)),
 *     PixToNdc(i, S) = -1 + (2 i + 1) / S;
 *   - a face is skipped if max z < 0, if |EdgeFunction(v0,v1,v2)| <= kEpsilon (1e-8), or if p lies outside its xy bbox;
 *   - barycentrics w0 = E(p,v1,v2)/A, w1 = E(p,v2,v0)/A, w2 = E(p,v0,v1)/A with A = E(v2,v0,v1) + kEpsilon,
 *     E(p,a,b) = (p.x-a.x)(b.y-a.y) - (p.y-a.y)(b.x-a.x); no back-face culling;
 *   - pz = w0 z0 + w1 z1 + w2 z2 (perspective_correct False); skipped if pz < 0;
 *   - blur_radius 0: kept only if w0, w1, w2 > 0 (strictly inside);
 *   - faces_per_pixel 1: the smallest pz wins; on an exact tie the earlier (lower-index) face stays.
 * Outputs (in place, like the rest of this file): depth = zbuf (caller initialises it, e.g. +inf), tri = face index within
 * the mesh (pytorch3d's pix_to_face adds the packed offset b*F), out3 = barycentrics.  fp32, no FMA contraction.
 * (kEpsilon: 1e-8 as in pytorch3d's float_math.cuh of the 0.2 line; unpinned like the rest.) */
static float p3d_pix_to_ndc(int i, int S) { return -1.0f + (2 * i + 1.0f) / S; }

static float p3d_edge(float px, float py, float ax, float ay, float bx, float by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

void oracle_rasterize_pytorch3d(const float* fv, float* depth, int32_t* tri, float* bary, int batch, int ntri, int h, int w) {
    const float eps = 1e-8f;
    for (int b = 0; b < batch; ++b)
        for (int yi = 0; yi < h; ++yi)
            for (int xi = 0; xi < w; ++xi) {
                const float xf = p3d_pix_to_ndc(w - 1 - xi, w), yf = p3d_pix_to_ndc(h - 1 - yi, h);
                size_t pix = ((size_t)b * h + yi) * w + xi;
                int owner = INT32_MAX;
                for (int f = 0; f < ntri; ++f) {
                    const float* fc = fv + ((size_t)b * ntri + f) * 9;
                    float x0 = fc[0], y0 = fc[1], z0 = fc[2], x1 = fc[3], y1 = fc[4], z1 = fc[5], x2 = fc[6], y2 = fc[7], z2 = fc[8];
                    if (fmaxf(z0, fmaxf(z1, z2)) < 0.0f) continue;
                    float area = p3d_edge(x0, y0, x1, y1, x2, y2);
                    if (area <= eps && area >= -eps) continue;
                    if (xf > fmaxf(x0, fmaxf(x1, x2)) || xf < fminf(x0, fminf(x1, x2)) ||
                        yf > fmaxf(y0, fmaxf(y1, y2)) || yf < fminf(y0, fminf(y1, y2))) continue;
                    float A = p3d_edge(x2, y2, x0, y0, x1, y1) + eps;
                    float w0 = p3d_edge(xf, yf, x1, y1, x2, y2) / A;
                    float w1 = p3d_edge(xf, yf, x2, y2, x0, y0) / A;
                    float w2 = p3d_edge(xf, yf, x0, y0, x1, y1) / A;
                    float pz = w0 * z0 + w1 * z1 + w2 * z2;
                    if (pz < 0.0f) continue;
                    if (!(w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) continue;
                    int take = (pz < depth[pix]) || (pz == depth[pix] && f < owner);
                    if (!take) continue;
                    depth[pix] = pz; owner = f; tri[pix] = f;
                    bary[pix * 3 + 0] = w0; bary[pix * 3 + 1] = w1; bary[pix * 3 + 2] = w2;
                }
            }
}
