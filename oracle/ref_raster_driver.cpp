// TEST INFRASTRUCTURE ONLY -- host driver that executes the UNMODIFIED reference kernels
// (my_utils/standard_rasterize_cuda/standard_rasterize_cuda_kernel.cu:112-233) on the CPU.
//
// oracle/Makefile extracts the kernel part of the reference .cu (everything before the host launchers at :237,
// which use <<<>>> syntax g++ cannot parse) into oracle/_ref/ref_kernels.inc *at build time* (never committed)
// and compiles it here against oracle/shim/.  The launch configuration of the reference launchers
// (:249-269 and :291-314: 512 threads per block, ceil(B*ntri/512) blocks, the SAME kernel launched TWICE) is
// reproduced by the loops below; threads run sequentially in ascending global index, so on exact zp ties the
// highest face index wins (on the GPU the winner of a tie is a race).
#include "cuda_on_cpu.h"
#include "ref_kernels.inc"

template <typename F>
static void launch_twice(int total, F&& body) {
    const int threads = 512;
    const int blocks = (total - 1) / threads + 1;
    blockDim.x = threads;
    for (int pass = 0; pass < 2; ++pass)
        for (int b = 0; b < blocks; ++b)
            for (int t = 0; t < threads; ++t) {
                blockIdx.x = b;
                threadIdx.x = t;
                body();
            }
}

extern "C" void ref_standard_rasterize(const float* face_vertices, float* depth, int* tri, float* bary,
                                       int batch, int ntri, int h, int w) {
    launch_twice(batch * ntri, [&] {
        forward_rasterize_cuda_kernel<float>(face_vertices, depth, tri, bary, batch, h, w, ntri);
    });
}

extern "C" void ref_standard_rasterize_colors(const float* face_vertices, const float* face_colors, float* depth,
                                              int* tri, float* images, int batch, int ntri, int h, int w) {
    launch_twice(batch * ntri, [&] {
        forward_rasterize_colors_cuda_kernel<float>(face_vertices, face_colors, depth, tri, images, batch, h, w, ntri);
    });
}
