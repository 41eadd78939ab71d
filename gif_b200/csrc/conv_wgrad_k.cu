// tcgen05 weight gradient, K-major variant (stride-1 convolutions, W % 32 == 0):
//     R[t][cs][cb] = sum_{pixels} ST[cs][pixel] * BT[cb][pixel + tap offset]
// on PIXEL-CONTIGUOUS ("NCHW") tf32 copies ST (B,Cs,H,W) / BT (B,Cb,H,W) of the two operands.  ncu shows that 32-bit
// MN-major operands (conv_wgrad_tc.cu, channels-last inputs) occupy the tensor pipe twice as long per MMA; with pixel-
// contiguous operands both tiles are ordinary K-major SWIZZLE_128B tiles -- [128 channels][32 pixels] = one 16 KB TMA box
// -- and the MMA runs at the full kind::tf32 rate.  The kw tap shift is a shift of the TMA box start along W (OOB zero
// fill = padding), kh a shift along H.  Otherwise identical to conv_wgrad_tc.cu: a CTA owns one kernel row and all KW taps
// (KW TMEM accumulators), split-K over pixels with per-split partial buffers + wgrad_reduce.
#include "tc_common.cuh"

namespace gifb200 {
namespace {

constexpr int kStagesK = 3;
constexpr int kPixK = 32;

struct WgkParams {
    int B, H, W, Cs, Cb, k, pad, flip, splits;
    long long units;         // 32-pixel units: B * H * (W/32)
    long long stride_t, stride_cs, stride_cb, part_stride;
};

template <int KW, int BLOCK_N>
struct WgkSmem {
    static constexpr int kABytes = 128 * kPixK * 4;
    static constexpr int kBBytesPerTap = BLOCK_N * kPixK * 4;
    static constexpr int kStageBytes = kABytes + KW * kBBytesPerTap;
    static constexpr int kBarrierOffset = kStagesK * kStageBytes;
    static constexpr int kDynamic = kBarrierOffset + 128 + 1024;
    static constexpr int kTmemCols = (KW * BLOCK_N <= 32) ? 32 : (KW * BLOCK_N <= 64) ? 64 : (KW * BLOCK_N <= 128) ? 128
                                     : (KW * BLOCK_N <= 256) ? 256 : 512;
};

template <int KW, int BLOCK_N>
__global__ void __launch_bounds__(256, 1) wgrad_k_kernel(const __grid_constant__ CUtensorMap map_s,
                                                         const __grid_constant__ CUtensorMap map_b,
                                                         float* __restrict__ part, const WgkParams p) {
    using L = WgkSmem<KW, BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
    uint64_t* empty_bar = full_bar + kStagesK;
    uint64_t* tmem_full_bar = empty_bar + kStagesK;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ms = blockIdx.x, nb = blockIdx.y;
    const int kh = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
    const long long per = (p.units + p.splits - 1) / p.splits;
    const long long u0 = split * per;
    const long long u1 = u0 + per < p.units ? u0 + per : p.units;
    const int iters = u1 > u0 ? static_cast<int>(u1 - u0) : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_s) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStagesK; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(L::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (iters > 0) {
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer: 1 + KW box loads per stage =====================
            const int segs = p.W / kPixK;
            int stage = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                const long long u = u0 + it;
                const long long row = u / segs;
                const int x0 = static_cast<int>(u % segs) * kPixK;
                const int n = static_cast<int>(row / p.H), y = static_cast<int>(row % p.H);
                mbar_wait(&empty_bar[stage], ph ^ 1);
                uint8_t* a_dst = smem + stage * L::kStageBytes;
                mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                tma_load_4d(a_dst, &map_s, &full_bar[stage], x0, y, ms * 128, n);
#pragma unroll
                for (int kw = 0; kw < KW; ++kw)
                    tma_load_4d(a_dst + L::kABytes + kw * L::kBBytesPerTap, &map_b, &full_bar[stage], x0 + kw - p.pad,
                                y + kh - p.pad, nb * BLOCK_N, n);
                if (++stage == kStagesK) { stage = 0; ph ^= 1; }
            }
        } else if (warp == 1 && lane == 0) {
            // ===================== MMA issuer (both operands K-major) =====================
            constexpr uint32_t idesc = make_idesc_tf32(128, BLOCK_N);
            int stage = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(&full_bar[stage], ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
                const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
#pragma unroll
                for (int kw = 0; kw < KW; ++kw) {
                    const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + L::kABytes + kw * L::kBBytesPerTap);
#pragma unroll
                    for (int j = 0; j < kPixK / 8; ++j)
                        umma_tf32(tmem_base + kw * BLOCK_N, adesc + 2 * j, bdesc + 2 * j, idesc, (it | j) != 0);
                }
                umma_commit(&empty_bar[stage]);
                if (++stage == kStagesK) { stage = 0; ph ^= 1; }
            }
            umma_commit(tmem_full_bar);
        } else if (warp >= 4) {
            // ===================== epilogue: TMEM -> this split's slice of the partial buffer =====================
            mbar_wait(tmem_full_bar, 0);
            tcgen05_fence_after();
            const int q = warp - 4;
            const int cs = ms * 128 + q * 32 + lane;
            float* pbase = part + split * p.part_stride;
#pragma unroll 1
            for (int kw = 0; kw < KW; ++kw) {
                const int t = kh * p.k + kw;
                float* obase = pbase + (static_cast<long long>(t) * p.Cs + cs) * p.Cb + nb * BLOCK_N;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kw * BLOCK_N + c;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j += 8) st_global_v8(obase + c + j, v + j);
                }
            }
            tcgen05_fence_before();
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int cs = ms * 128 + q * 32 + lane;
        float* pbase = part + split * p.part_stride;
        for (int kw = 0; kw < KW; ++kw) {
            const int t = kh * p.k + kw;
            float* obase = pbase + (static_cast<long long>(t) * p.Cs + cs) * p.Cb + nb * BLOCK_N;
            for (int c = 0; c < BLOCK_N; c += 4) *reinterpret_cast<float4*>(obase + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::kTmemCols) : "memory");
    }
}

__global__ void __launch_bounds__(256) wgrad_k_reduce_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                             WgkParams p, int T) {
    const long long total = static_cast<long long>(T) * p.Cs * p.Cb;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < p.splits; ++s) acc += part[s * p.part_stride + e];
        const int cb = static_cast<int>(e % p.Cb);
        const int cs = static_cast<int>((e / p.Cb) % p.Cs);
        const int t = static_cast<int>(e / (static_cast<long long>(p.Cb) * p.Cs));
        const int tt = p.flip ? T - 1 - t : t;
        out[tt * p.stride_t + cs * p.stride_cs + cb * p.stride_cb] = acc;
    }
}

// (B,P,C) channels-last -> (B,C,P) pixel-contiguous, rounded to tf32.  32x32 tiles through shared memory: both the reads
// (channels) and the writes (pixels) are 128-byte coalesced.
__global__ void __launch_bounds__(256) nhwc_to_nchw_tf32_kernel(const float* __restrict__ x, float* __restrict__ y, int P, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xb = x + static_cast<long long>(b) * P * C;
    float* yb = y + static_cast<long long>(b) * P * C;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int pp = p0 + ty + j, c = c0 + tx;
        tile[ty + j][tx] = (pp < P && c < C) ? __ldcs(xb + static_cast<long long>(pp) * C + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int c = c0 + ty + j, pp = p0 + tx;
        if (c < C && pp < P) yb[static_cast<long long>(c) * P + pp] = round_tf32(tile[tx][ty + j]);
    }
}

int pick_bn_k(int Cb) { return Cb % 128 == 0 ? 128 : Cb % 64 == 0 ? 64 : Cb % 32 == 0 ? 32 : 0; }

long long wgk_splits(int B, int H, int W, int Cs, int Cb, int k) {
    const long long units = static_cast<long long>(B) * H * (W / kPixK);
    const long long base = static_cast<long long>(Cs / 128) * (Cb / pick_bn_k(Cb)) * k;
    long long splits = (kNumSMs + base - 1) / base;
    if (splits > units) splits = units;
    return splits < 1 ? 1 : splits;
}

template <int KW, int BLOCK_N>
int launch_wgk(const CUtensorMap& ms, const CUtensorMap& mb, float* out, float* part, const WgkParams& p, cudaStream_t st) {
    using L = WgkSmem<KW, BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_k_kernel<KW, BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
        if (e != cudaSuccess) return fail(GIFB200_E_CUDA, "cudaFuncSetAttribute(wgrad_k_kernel)", cudaGetErrorString(e));
        attr_set = true;
    }
    wgrad_k_kernel<KW, BLOCK_N><<<dim3(p.Cs / 128, p.Cb / BLOCK_N, p.k * p.splits), 256, L::kDynamic, st>>>(ms, mb, part, p);
    GIFB200_LAUNCH_CHECK("wgrad_k_kernel");
    const int T = p.k * p.k;
    int blocks = cdiv(static_cast<long long>(T) * p.Cs * p.Cb, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    wgrad_k_reduce_kernel<<<blocks, 256, 0, st>>>(part, out, p, T);
    GIFB200_LAUNCH_CHECK("wgrad_k_reduce_kernel");
    return GIFB200_OK;
}

}  // namespace

bool conv2d_wgrad_k_supported(int B, int H, int W, int Ci, int Co, int k) {
    return B > 0 && (k == 1 || k == 3) && W % kPixK == 0 && H >= 1 && Co % 128 == 0 && pick_bn_k(Ci) != 0;
}

// workspace: [ST: B*Co*H*W][BT: B*Ci*H*W] (only the copies that the caller does not supply) + split-K partials
size_t conv2d_wgrad_k_workspace_bytes(int B, int H, int W, int Ci, int Co, int k, int have_xT, int have_gyT) {
    if (!conv2d_wgrad_k_supported(B, H, W, Ci, Co, k)) return 0;
    size_t n = static_cast<size_t>(wgk_splits(B, H, W, Co, Ci, k)) * k * k * Co * Ci * sizeof(float) + 1024;
    if (!have_gyT) n += static_cast<size_t>(B) * Co * H * W * sizeof(float) + 256;
    if (!have_xT) n += static_cast<size_t>(B) * Ci * H * W * sizeof(float) + 256;
    return n;
}

int nhwc_to_nchw_tf32(const float* x, float* y, int B, int P, int C, cudaStream_t st) {
    if (B == 0) return GIFB200_OK;
    GIFB200_REQUIRE(cdiv(C, 32) <= 65535 && B <= 65535, GIFB200_E_SHAPE, "nhwc_to_nchw: grid too large");
    nhwc_to_nchw_tf32_kernel<<<dim3(cdiv(P, 32), cdiv(C, 32), B), 256, 0, st>>>(x, y, P, C);
    GIFB200_LAUNCH_CHECK("nhwc_to_nchw_tf32_kernel");
    return GIFB200_OK;
}

// S1 convolution weight gradient.  x (B,H,W,Ci), gy (B,H,W,Co) channels-last (used to build the pixel-contiguous copies when
// xT / gyT are NULL); xT (B,Ci,H,W), gyT (B,Co,H,W) optional pre-built tf32 copies.
int conv2d_wgrad_k(const float* x, const float* gy, const float* xT, const float* gyT, float* gw, int B, int H, int W, int Ci,
                   int Co, int k, int flip, int transposed, void* ws, size_t ws_bytes, cudaStream_t st) {
    GIFB200_REQUIRE(conv2d_wgrad_k_supported(B, H, W, Ci, Co, k), GIFB200_E_SHAPE, "conv2d_wgrad_k: unsupported shape");
    GIFB200_REQUIRE(ws && ws_bytes >= conv2d_wgrad_k_workspace_bytes(B, H, W, Ci, Co, k, xT != nullptr, gyT != nullptr),
                    GIFB200_E_WORKSPACE, "conv2d_wgrad_k: workspace too small");
    char* wp = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
    const long long P = static_cast<long long>(H) * W;
    if (!gyT) {
        float* buf = reinterpret_cast<float*>(wp);
        int rc = nhwc_to_nchw_tf32(gy, buf, B, static_cast<int>(P), Co, st);
        if (rc != GIFB200_OK) return rc;
        gyT = buf;
        wp += ((static_cast<size_t>(B) * Co * P * 4 + 255) / 256) * 256;
    }
    if (!xT) {
        float* buf = reinterpret_cast<float*>(wp);
        int rc = nhwc_to_nchw_tf32(x, buf, B, static_cast<int>(P), Ci, st);
        if (rc != GIFB200_OK) return rc;
        xT = buf;
        wp += ((static_cast<size_t>(B) * Ci * P * 4 + 255) / 256) * 256;
    }
    float* part = reinterpret_cast<float*>(wp);
    WgkParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W; p.Cs = Co; p.Cb = Ci; p.k = k; p.pad = k / 2; p.flip = flip;
    p.units = static_cast<long long>(B) * H * (W / kPixK);
    p.splits = static_cast<int>(wgk_splits(B, H, W, Co, Ci, k));
    p.part_stride = static_cast<long long>(k) * k * Co * Ci;
    p.stride_t = static_cast<long long>(Co) * Ci;
    p.stride_cs = transposed ? 1 : Ci;       // cs = o
    p.stride_cb = transposed ? Co : 1;       // cb = i
    const int bn = pick_bn_k(Ci);
    CUtensorMap ms, mb;
    {
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(Co), static_cast<cuuint64_t>(B)};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 4, static_cast<cuuint64_t>(P) * 4, static_cast<cuuint64_t>(Co) * P * 4};
        const cuuint32_t box[4] = {kPixK, 1, 128, 1};
        int rc = encode_map(&ms, gyT, 4, dims, strides, box);
        if (rc != GIFB200_OK) return rc;
    }
    {
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(Ci), static_cast<cuuint64_t>(B)};
        const cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 4, static_cast<cuuint64_t>(P) * 4, static_cast<cuuint64_t>(Ci) * P * 4};
        const cuuint32_t box[4] = {kPixK, 1, static_cast<cuuint32_t>(bn), 1};
        int rc = encode_map(&mb, xT, 4, dims, strides, box);
        if (rc != GIFB200_OK) return rc;
    }
#define GIFB200_WGK(KW, BN) launch_wgk<KW, BN>(ms, mb, gw, part, p, st)
    if (k == 3) return bn == 128 ? GIFB200_WGK(3, 128) : bn == 64 ? GIFB200_WGK(3, 64) : GIFB200_WGK(3, 32);
    return bn == 128 ? GIFB200_WGK(1, 128) : bn == 64 ? GIFB200_WGK(1, 64) : GIFB200_WGK(1, 32);
#undef GIFB200_WGK
}

}  // namespace gifb200
