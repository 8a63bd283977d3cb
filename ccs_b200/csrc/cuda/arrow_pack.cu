// arrow_pack_rowcodes: Recursor::EncodeRead on the device.  The batch's emission codes are
// already resident (the Draft Stage uploaded them once); this kernel lays out, per mapped read, the two row-indexed
// copies the recursion kernels index without bounds checks:
//   copy A[i] = 4 * code of DP row i   (sentinel at row 0 and at rows >= I: the last read base is consumed only by the
//                                        pinned final match),
//   copy B[i] = A[i + 1]               (the backward pass and the link step look one row ahead),
// both `stride` bytes long (a multiple of 16), codes pre-multiplied by 4 (= byte offset into a table row).
// One CTA per read, a thread writes 4 rows of each copy per step: ~1 byte read and 2 bytes written per read base --
// noise next to the 256 bytes per base of the band stores, but it takes the pack loop and 2/3 of the upload off the host.
#include <cuda_runtime.h>
#include <cstdint>
#include "arrow_device.h"
#include "arrow_launch.h"

namespace ccs { constexpr int kCodeSentinel = 12; }   // the zero-emission code (common/arrow_tables.h)

namespace ccs {

namespace {

__global__ void __launch_bounds__(256) arrow_pack_rowcodes_kernel(const PackJob* __restrict__ jobs, const int n_jobs,
                                                                  const uint8_t* __restrict__ codes,
                                                                  uint8_t* __restrict__ rowcode) {
    const int j = blockIdx.x;
    if (j >= n_jobs) return;
    const PackJob J = jobs[j];
    const uint8_t* __restrict__ src = codes + J.src_off;
    uint32_t* __restrict__ A = reinterpret_cast<uint32_t*>(rowcode + J.dst_off);
    uint32_t* __restrict__ B = reinterpret_cast<uint32_t*>(rowcode + J.dst_off + J.stride);
    const int I = J.I;
    const unsigned sent = 4u * kCodeSentinel;
    for (int w = threadIdx.x; w < (J.stride >> 2); w += blockDim.x) {
        const int i0 = 4 * w;
        unsigned c[5];          // 4 * code of read bases i0-1 .. i0+3 where they exist and are used, else the sentinel
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int b = i0 - 1 + k;                      // read base index; row i uses base i-1, rows 1..I-1 only
            c[k] = (b >= 0 && b <= I - 2) ? 4u * src[b] : sent;
        }
        A[w] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
        B[w] = c[1] | (c[2] << 8) | (c[3] << 16) | (c[4] << 24);
    }
}

}  // namespace

void launch_pack_rowcodes(const PackJob* jobs, int n_jobs, const uint8_t* codes, uint8_t* rowcode, cudaStream_t stream) {
    if (n_jobs <= 0) return;
    arrow_pack_rowcodes_kernel<<<n_jobs, 256, 0, stream>>>(jobs, n_jobs, codes, rowcode);
}

}  // namespace ccs
