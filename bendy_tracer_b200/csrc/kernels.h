// kernels.h -- host-callable launchers of the CUDA kernels in kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"

namespace bt {

struct DeviceSegment {  // mirrors bt_segment (include/bendy_b200.h) with the object INDEX
    int32_t face;
    uint32_t steps;
    int32_t obj;
    float t;
    float position[3], normal[3], direction[3];
};

enum { INTEGRATE_INLINE_LENSES = 16 };
struct IntegrateParams {
    // tables of up to INTEGRATE_INLINE_LENSES masses travel BY VALUE in the kernel parameters (no upload, nothing
    // for an asynchronous caller to race with); larger ones are read from `lens`
    float4 inline_lens[INTEGRATE_INLINE_LENSES * LENS_STRIDE];
    const float4* lens;   // LENS_STRIDE records, device memory (n_lens > INTEGRATE_INLINE_LENSES)
    uint32_t n_lens;
    float kappa, h_min, h_max;
    float* xv;            // n * 6
    uint32_t n, n_steps;
    uint32_t exact;       // correctly rounded rsqrt (bit-identical to the CPU oracle)
};

// Each returns the cudaError_t of the launch; `launches` is incremented per kernel launched.
// kernels.cu is compiled twice: the `_fast` flavour (MUFU-based reciprocal / square roots, FMA rect
// tests: ~1 ulp from IEEE) and the `_exact` flavour (-DBT_EXACT_SCAN: every operation is the IEEE
// one the reference performs; bit-identical to the CPU oracle).  engine.cu picks per scene.
#define BT_DECLARE_LAUNCHERS(SFX)                                                                                        \
    cudaError_t launch_render##SFX(const RenderParams& p, cudaStream_t stream, uint64_t* launches);                      \
    cudaError_t launch_trace##SFX(const RenderParams& p, uint32_t n, const float* origins, const float* dirs,           \
                                  DeviceSegment* out, cudaStream_t stream, uint64_t* launches);                         \
    cudaError_t launch_camera_rays##SFX(const RenderParams& p, uint32_t n, const uint32_t* xs, const uint32_t* ys,      \
                                        const uint64_t* path_index, float* out, cudaStream_t stream, uint64_t* launches);
BT_DECLARE_LAUNCHERS(_fast)
BT_DECLARE_LAUNCHERS(_exact)
#undef BT_DECLARE_LAUNCHERS
cudaError_t launch_integrate(const IntegrateParams& p, cudaStream_t stream, uint64_t* launches);
cudaError_t launch_resolve(const float4* fb, uint32_t n_pixels, uint64_t samples, int color_space, uchar4* out,
                           cudaStream_t stream, uint64_t* launches);
// slices of a multi-device render, summed into the frame on the device that owns it (peer-mapped or local pointers)
enum { MAX_PEER_FRAMES = 15 };
struct PeerFrames {
    const float4* src[MAX_PEER_FRAMES];
    int n;
};
cudaError_t launch_accumulate_frames(float4* dst, const PeerFrames& pf, uint32_t n_pixels, int sm_count, cudaStream_t stream, uint64_t* launches);
cudaError_t launch_fp32_peak(float* out, uint32_t iters, int blocks, cudaStream_t stream, uint64_t* launches);

size_t render_smem_bytes(const RenderParams& p, unsigned threads = 256);
// device bytes the pooled render kernel wants for its path-state arena on a GPU of `sm_count` SMs (an upper bound of its grid)
size_t render_pool_arena_bytes(uint32_t pool_w, int sm_count, bool bvh);
enum { FP32_PEAK_THREADS = 256, FP32_PEAK_CHAINS = 8, FP32_PEAK_UNROLL = 16 };

}  // namespace bt
