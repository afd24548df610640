// trt_multi.cu — the one exchange step of the multi-GPU render, fused with the epilogue.
//
// Samples-per-pixel are partitioned across the GPUs of one box; afterwards GPU k holds the f64 per-pixel sums of ITS samples
// for the whole image. The image is cut into N contiguous slices and GPU k produces slice k:
//
//   k_reduce_scatter_finalize   reads slice k of every GPU's accumulation buffer — its own through HBM, the others' directly
//                               through NVLink / NVSwitch peer loads (the pointers are peer-mapped device memory; no staging
//                               copy, no NCCL) —, adds the partial sums IN RANK ORDER in f64 (so the result does not depend on
//                               timing, and equals the 1-GPU image to <= 1 ulp after rounding), and applies the imager's
//                               normalisation and tone map (imager.rs:35-60, utils/image.rs:92-111) in the same pass.
//
// i.e. reduce-scatter + epilogue in one kernel per GPU, all N running concurrently, each moving (N - 1)/N of 8 B x W*H*3 / N
// over the switch; the N result slices then go to the host over N PCIe links at once. HBM/NVLink-stream bound:
// per output value 8 B x N read, 5 B written.
#include "trt_device.cuh"
#include "trt_kernels.h"

using namespace trtd;

struct ReduceArgs {
    const double* src[TRT_MULTI_MAX_DEVICES];   // accumulation buffers of ranks 0..n_src-1 (whole image each), rank order
    int n_src;
    unsigned long long begin, end;              // this GPU's slice of the W*H*3 values; begin is even
    double inv_spp;
    float* lin;                                  // [end - begin] or null
    uint8_t* rgb8;                               // [end - begin] or null
};

__global__ void __launch_bounds__(256) k_reduce_scatter_finalize(const ReduceArgs A) {
    const unsigned long long n = A.end - A.begin, pairs = n >> 1;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < pairs; j += stride) {
        const unsigned long long i = A.begin + 2 * j;                   // even -> 16-byte aligned double2
        double2 s = *reinterpret_cast<const double2*>(A.src[0] + i);
        for (int r = 1; r < A.n_src; ++r) {
            const double2 v = *reinterpret_cast<const double2*>(A.src[r] + i);
            s.x += v.x; s.y += v.y;
        }
        const float l0 = (float)(s.x * A.inv_spp), l1 = (float)(s.y * A.inv_spp);   // imager.rs:35,50 (f64 here)
        if (A.lin) { A.lin[2 * j] = l0; A.lin[2 * j + 1] = l1; }
        if (A.rgb8) { A.rgb8[2 * j] = tonemap_u8(l0); A.rgb8[2 * j + 1] = tonemap_u8(l1); }
    }
    if ((n & 1ull) && blockIdx.x == 0 && threadIdx.x == 0) {             // odd tail value
        const unsigned long long i = A.end - 1;
        double s = A.src[0][i];
        for (int r = 1; r < A.n_src; ++r) s += A.src[r][i];
        const float l = (float)(s * A.inv_spp);
        if (A.lin) A.lin[n - 1] = l;
        if (A.rgb8) A.rgb8[n - 1] = tonemap_u8(l);
    }
}

namespace trtk {

cudaError_t launch_reduce_scatter_finalize(const double* const* d_src, int n_src, size_t begin, size_t end, uint32_t spp_total,
                                           float* d_lin, uint8_t* d_rgb8, int sm_count, cudaStream_t st) {
    if (n_src < 1 || n_src > TRT_MULTI_MAX_DEVICES || (begin & 1) || end < begin) return cudaErrorInvalidValue;
    if (end == begin) return cudaSuccess;
    ReduceArgs A;
    for (int r = 0; r < TRT_MULTI_MAX_DEVICES; ++r) A.src[r] = r < n_src ? d_src[r] : nullptr;
    A.n_src = n_src; A.begin = begin; A.end = end; A.inv_spp = 1.0 / (double)spp_total; A.lin = d_lin; A.rgb8 = d_rgb8;
    const size_t pairs = (end - begin + 1) / 2;
    size_t blocks = (pairs + 255) / 256;
    const size_t cap = (size_t)sm_count * 8;      // grid-stride: a few waves of 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_reduce_scatter_finalize<<<(unsigned)blocks, 256, 0, st>>>(A);
    return cudaGetLastError();
}

} // namespace trtk
