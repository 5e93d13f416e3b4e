// kernels.cu -- the CUDA kernels of the per-sample render loop (sm_100a).
//
// render_kernel: one lane = one pixel; the lane walks its pixel's paths in the reference's order
// (pass-major, sub-pixel minor: src/tracer/mod.rs:277-278) and regenerates a new path the moment
// the previous one terminates, so path-length divergence never idles a lane ("persistent lane
// with in-register path regeneration").  Path state lives in registers, the scene blob in shared
// memory; HBM traffic is one float4 read-modify-write per pixel per call (buffer.rs:159-178).
// Pixel sums are formed in the reference's order, so the result is deterministic.
#include "device.cuh"
#include "kernels.h"

namespace bt {

namespace {

struct SceneView {
    const float4* prims;
    const float4* mats;
    const float4* lights;
    const float4* vols;
    const float4* lens;
    const float* grids;
};

BT_DEV SceneView stage_scene(const RenderParams& p, float4* smem) {
    for (uint32_t i = threadIdx.x; i < p.scene.blob_f4; i += blockDim.x) smem[i] = p.blob[i];
    __syncthreads();
    SceneView s;
    s.prims = smem + p.scene.prim_off;
    s.mats = smem + p.scene.mat_off;
    s.lights = smem + p.scene.light_off;
    s.vols = smem + p.scene.vol_off;
    s.lens = smem + p.scene.lens_off;
    s.grids = p.grids;
    return s;
}

struct Traced {
    Hit h;            // h.prim < 0: miss (or capture)
    V3 o, d;          // the straight piece the hit lies on (chord under lensing) / escape direction
    float t_total;    // accumulated length up to the hit
    uint32_t steps;
    bool captured;
};

// One "ray" of the render loop.  Flat field (or inside a volume march): exactly try_hit /
// try_hit_volume.  Lens field: RK4 chords, each intersected with the same scan.
template <bool LENS, bool EXACT, class L>
BT_DEV Traced trace_ray(const RenderParams& p, const SceneView& sc, const L& lens, V3 o, V3 d, float tmin, float tmax, int vol_obj) {
    Traced r;
    r.steps = 0;
    r.captured = false;
    const int n_prims = (int)p.scene.n_prims;
    if (!LENS || vol_obj >= 0) {
        r.h = scan_prims(sc.prims, n_prims, o, d, tmin, tmax, vol_obj);
        r.o = o;
        r.d = d;
        r.t_total = r.h.t;
        return r;
    }
    V3 x = o, v = d;
    float travelled = 0.0f;
    for (;;) {
        float rmin;
        bool captured, far;
        D0Cache<L> cache;
        V3 k1 = lens_accel<2, EXACT, 0>(lens, cache, x, 0.0f, v, v, rmin, captured, far);
        if (captured) {
            r.captured = true;
            r.h.prim = -1;
            r.h.t = 0.0f;
            r.h.face = 0;
            r.o = x;
            r.d = v;
            r.t_total = travelled;
            return r;
        }
        float remaining = tmax - travelled;
        float cmin = fmaxf(tmin - travelled, 0.0f);
        if (far) {
            V3 dir = normalize_fma(v, 0);
            r.h = scan_prims(sc.prims, n_prims, x, dir, cmin, remaining, -1);
            r.o = x;
            r.d = dir;
            r.t_total = travelled + r.h.t;
            return r;
        }
        float h = step_size(p.scene.kappa, p.scene.h_min, p.scene.h_max, rmin);
        V3 x1 = x, v1 = v;
        rk4_from_k1<EXACT>(lens, cache, x1, v1, k1, h);
        float len;
        V3 dir = normalize_fma(x1 - x, &len);
        r.h = scan_prims(sc.prims, n_prims, x, dir, cmin, fminf(len, remaining), -1);
        if (r.h.prim >= 0) {
            r.o = x;
            r.d = dir;
            r.t_total = travelled + r.h.t;
            return r;
        }
        travelled += len;
        x = x1;
        v = v1;
        r.steps++;
        if (travelled >= tmax || r.steps >= p.scene.max_steps) {
            r.o = x;
            r.d = normalize_fma(v, 0);
            r.t_total = travelled;
            return r;
        }
    }
}

// NL: 0 = lens table walked in shared memory, N > 0 = exactly N masses held in registers
template <int NL>
struct LensSel {
    typedef LensRegs<NL> type;
    static BT_DEV type make(const float4* p, int) { return type(p); }
};
template <>
struct LensSel<0> {
    typedef LensShared type;
    static BT_DEV type make(const float4* p, int n) {
        LensShared l;
        l.p = p;
        l.count = n;
        return l;
    }
};

template <bool LENS, bool EXACT, int NL>
__global__ void __launch_bounds__(256) render_kernel(const __grid_constant__ RenderParams p) {
    extern __shared__ float4 smem[];
    const SceneView sc = stage_scene(p, smem);
    const typename LensSel<NL>::type lens = LensSel<NL>::make(sc.lens, (int)p.scene.n_lens);
    Consts k;
    k.tau_scale = p.tau_scale;
    k.one_scale = p.one_scale;

    // block = 16x16 pixels, warp = 8x4 pixel tile (coherent first hits, 128 B framebuffer rows)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t px = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    const bool valid = px < p.width && py < p.height;
    const uint64_t pixel = (uint64_t)py * p.width + px;
    const float inf = __int_as_float(0x7f800000);

    V3 acc = v3(0.0f, 0.0f, 0.0f);
    uint32_t path = 0;
    bool alive = false, done = !valid;

    // per-path state
    Rng rng;
    V3 o, d, T;
    uint32_t bounce = 0, vb = 0;
    int vol_obj = -1;
    bool latched = false;
    V3 aov_albedo, aov_normal;
    float aov_depth = inf;

    for (;;) {
        if (!alive && !done) {
            if (path < p.paths_per_pixel) {
                uint64_t path_index = p.path_base + path;
                rng.seed_from_u64(path_seed(p.seed, pixel, path_index));
                camera_ray(p.cam, k, rng, px, py, path % p.sub_count, o, d);
                T = v3(1.0f, 1.0f, 1.0f);
                bounce = 0;
                vb = 0;
                vol_obj = -1;
                latched = false;
                aov_albedo = v3(0.0f, 0.0f, 0.0f);
                aov_normal = v3(0.0f, 0.0f, 0.0f);
                aov_depth = inf;
                alive = true;
                ++path;
            } else {
                done = true;
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (!alive) continue;

        // ---- trace one segment ------------------------------------------------------------
        const bool in_volume = vol_obj >= 0;
        Traced tr = trace_ray<LENS, EXACT>(p, sc, lens, o, d, in_volume ? 0.0f : p.clip_min, in_volume ? p.volume_step : p.clip_max, vol_obj);

        // terminal outcome of this event (if any): colour and the AOVs it would latch
        bool finish = false;
        V3 fin_color = v3(0.0f, 0.0f, 0.0f), fin_albedo = v3(0.0f, 0.0f, 0.0f), fin_normal = v3(0.0f, 0.0f, 0.0f);
        float fin_depth = inf;

        if (tr.h.prim < 0) {
            finish = true;
            if (!tr.captured) {  // sample_root, mod.rs:429-452
                fin_color = v3(p.scene.root_color[0], p.scene.root_color[1], p.scene.root_color[2]);
                fin_albedo = v3(p.scene.root_albedo[0], p.scene.root_albedo[1], p.scene.root_albedo[2]);
                if (p.scene.root_keeps_normal) {
                    fin_normal = -tr.d;
                    fin_depth = p.clip_max;
                }
            }
        } else {
            const Surface s = resolve_hit(sc.prims, tr.h, tr.o, tr.d);
            const V3 din = tr.d;
            if (s.face <= 1) {
                // ---- sample_surface, mod.rs:454-486 + Material::shade, material.rs:81-199 ----
                const float4 m0 = sc.mats[s.mat * MAT_STRIDE], m1 = sc.mats[s.mat * MAT_STRIDE + 1];
                const int kind = __float_as_int(m0.w);
                const V3 A = v3(m0);
                if (kind == MAT_FLAT) {
                    finish = true;  // ColorData::from_emitted(albedo)
                    fin_color = A;
                    fin_albedo = A;
                } else if (kind == MAT_EMISSIVE) {
                    finish = true;
                    fin_color = A * m1.z;
                    fin_albedo = fin_color;
                } else {
                    V3 nd;
                    float pdf = 1.0f, mpdf = 1.0f;
                    if (kind == MAT_DIFFUSE) {
                        const uint32_t li = uniform_index(rng, p.scene.n_lights);
                        const float4* light = sc.lights + li * LIGHT_STRIDE;
                        if (gen_bool(rng, 0.5f))  // Pdf::Mix: true selects the light (material.rs:269-275)
                            nd = normalize_a(light_point(rng, k, light) - s.position);
                        else
                            nd = normalize_a(cosine_dir(rng, k, s.normal));
                        mpdf = dot(s.normal, nd) * 0.318309886183790671538f;
                        const float pb = light_pdf(sc.prims, light, s.position, nd, p.clip_min, p.clip_max);
                        pdf = lerpf(mpdf, pb, 0.5f);
                    } else if (kind == MAT_METALLIC) {
                        const V3 dir = reflect(din, s.normal);
                        nd = normalize_a(dir + unit_hemisphere(rng, k, s.normal) * m1.x);
                    } else {  // MAT_GLASS, material.rs:231-261
                        float ior = m1.y;
                        if (s.face == 0) ior = 1.0f / ior;
                        const float cos_theta = fminf(dot(-din, s.normal), 1.0f);
                        const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
                        const float fr = fresnel(din, s.normal, ior);
                        V3 dir;
                        if (ior * sin_theta > 1.0f || gen_bool(rng, fr))
                            dir = reflect(din, s.normal);
                        else
                            dir = refract(din, s.normal, ior);
                        nd = normalize_a(dir + unit_hemisphere(rng, k, s.normal) * m1.x);
                    }
                    if (fabsf(pdf) <= 1e-5f) {
                        finish = true;  // no scatter: from_emitted(BLACK)
                    } else {
                        if (!latched) {
                            latched = true;
                            aov_albedo = A;
                            aov_normal = s.normal;
                            aov_depth = tr.t_total;
                        }
                        T = T * ((A * mpdf) * (1.0f / pdf));
                        o = s.position;
                        d = nd;
                        vol_obj = -1;
                        ++bounce;
                        if (bounce > p.max_bounces) finish = true;  // sample(): black, AOVs already latched
                    }
                }
            } else {
                // ---- sample_volume, mod.rs:488-523 + Volume::shade, volume.rs:26-60 ----
                if (!in_volume) vb = 0;  // entered from sample(): volume_bounce = 0
                const V3 bmin = v3(s.center.x - s.radius, s.center.y - s.radius, s.center.z - s.radius);
                const V3 bmax = v3(s.center.x + s.radius, s.center.y + s.radius, s.center.z + s.radius);
                const V3 coord = (s.position - bmin) / (bmax - bmin);
                const float density = p.volume_step * density_trilinear(sc.vols + s.vol * VOL_STRIDE, sc.grids, coord);
                if (density >= 1.0f || gen_bool(rng, density)) {
                    V3 origin = s.position;
                    if (s.face == 2) origin = origin - din * p.volume_step * standard_f32(rng);
                    d = normalize_a(unit_sphere(rng, k));
                    o = origin;
                    if (!latched) {
                        latched = true;
                        aov_albedo = v3(0.8f, 0.8f, 0.8f);
                        aov_normal = s.normal;
                        aov_depth = tr.t_total;
                    }
                    T = T * 0.8f;
                } else {
                    o = s.position;
                    d = normalize_a(din);
                }
                if (s.face == 4) {  // VolumeBack: leave the medium through sample(ray, bounce + 1)
                    vol_obj = -1;
                    ++bounce;
                    if (bounce > p.max_bounces) finish = true;
                } else {            // keep marching: sample_volumetric(.., volume_bounce + 1)
                    vol_obj = s.obj;
                    ++vb;
                    if (vb > p.max_volume_bounces) finish = true;
                }
            }
        }

        if (finish) {
            if (!latched) {
                aov_albedo = fin_albedo;
                aov_normal = fin_normal;
                aov_depth = fin_depth;
            }
            switch (p.output) {  // mod.rs:306-315
                case 0: acc = acc + T * fin_color; break;
                case 1: acc = acc + aov_albedo; break;
                case 2: acc = acc + aov_normal; break;
                default: {
                    float dn = (aov_depth - p.clip_min) / (p.clip_max - p.clip_min);
                    dn = fminf(fmaxf(dn, 0.0f), 1.0f);
                    acc = acc + v3(dn, dn, dn);
                }
            }
            alive = false;
        }
    }

    if (valid) {  // Buffer::write_color: rgb += value, alpha untouched (buffer.rs:159-178)
        float4 v = p.fb[pixel];
        v.x += acc.x;
        v.y += acc.y;
        v.z += acc.z;
        p.fb[pixel] = v;
    }
}

template <bool LENS, bool EXACT, int NL>
__global__ void __launch_bounds__(256) trace_kernel(const __grid_constant__ RenderParams p, uint32_t n, const float* __restrict__ origins,
                                                    const float* __restrict__ dirs, DeviceSegment* __restrict__ out) {
    extern __shared__ float4 smem[];
    const SceneView sc = stage_scene(p, smem);
    const typename LensSel<NL>::type lens = LensSel<NL>::make(sc.lens, (int)p.scene.n_lens);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 o = v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
    V3 d = v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    Traced tr = trace_ray<LENS, EXACT>(p, sc, lens, o, d, p.clip_min, p.clip_max, -1);
    DeviceSegment seg;
    seg.steps = tr.steps;
    seg.obj = -1;
    seg.t = 0.0f;
    for (int c = 0; c < 3; ++c) seg.position[c] = seg.normal[c] = seg.direction[c] = 0.0f;
    if (tr.h.prim >= 0) {
        Surface s = resolve_hit(sc.prims, tr.h, tr.o, tr.d);
        seg.face = s.face;
        seg.obj = s.obj;
        seg.t = tr.t_total;
        seg.position[0] = s.position.x; seg.position[1] = s.position.y; seg.position[2] = s.position.z;
        seg.normal[0] = s.normal.x; seg.normal[1] = s.normal.y; seg.normal[2] = s.normal.z;
        seg.direction[0] = tr.d.x; seg.direction[1] = tr.d.y; seg.direction[2] = tr.d.z;
    } else if (tr.captured) {
        seg.face = -2;
    } else {
        seg.face = -1;
        seg.position[0] = tr.o.x; seg.position[1] = tr.o.y; seg.position[2] = tr.o.z;
        seg.direction[0] = tr.d.x; seg.direction[1] = tr.d.y; seg.direction[2] = tr.d.z;
    }
    out[i] = seg;
}

__global__ void camera_rays_kernel(const __grid_constant__ RenderParams p, uint32_t n, const uint32_t* __restrict__ xs,
                                   const uint32_t* __restrict__ ys, const uint64_t* __restrict__ path_index, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Consts k;
    k.tau_scale = p.tau_scale;
    k.one_scale = p.one_scale;
    Rng rng;
    const uint64_t pixel = (uint64_t)ys[i] * p.width + xs[i];
    rng.seed_from_u64(path_seed(p.seed, pixel, p.path_base + path_index[i]));
    V3 o, d;
    camera_ray(p.cam, k, rng, xs[i], ys[i], (uint32_t)(path_index[i] % p.sub_count), o, d);
    out[6 * i] = o.x; out[6 * i + 1] = o.y; out[6 * i + 2] = o.z;
    out[6 * i + 3] = d.x; out[6 * i + 4] = d.y; out[6 * i + 5] = d.z;
}

// The geodesic stepper in isolation: n_steps RK4 steps per ray, state in registers, the lens
// table in shared memory.  No memory traffic in the loop: this is the FP32-roofline kernel.
template <bool EXACT, int NL>
__global__ void __launch_bounds__(256) integrate_kernel(const IntegrateParams p) {
    extern __shared__ float4 slens[];
    for (uint32_t i = threadIdx.x; i < p.n_lens * LENS_STRIDE; i += blockDim.x) slens[i] = p.lens[i];
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float* s = p.xv + 6 * (size_t)i;
    V3 x = v3(s[0], s[1], s[2]), v = v3(s[3], s[4], s[5]);
    const typename LensSel<NL>::type lens = LensSel<NL>::make(slens, (int)p.n_lens);
#pragma unroll 1
    for (uint32_t it = 0; it < p.n_steps; ++it) {
        float rmin;
        bool captured, far;
        D0Cache<typename LensSel<NL>::type> cache;
        V3 k1 = lens_accel<1, EXACT, 0>(lens, cache, x, 0.0f, v, v, rmin, captured, far);
        rk4_from_k1<EXACT>(lens, cache, x, v, k1, step_size(p.kappa, p.h_min, p.h_max, rmin));
    }
    s[0] = x.x; s[1] = x.y; s[2] = x.z; s[3] = v.x; s[4] = v.y; s[5] = v.z;
}

// Buffer::preview, buffer.rs:117-138
BT_DEV float linear_to_srgb(float x) {  // color.rs:14-20
    if (x <= 0.0031308f) return 12.92f * x;
    return 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
}
BT_DEV unsigned char f32_to_u8(float x) {  // color.rs:22-24, Rust saturating cast (NaN -> 0)
    float v = x * 255.0f;
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (unsigned char)v;
}
__global__ void resolve_kernel(const float4* __restrict__ fb, uint32_t n, float samples_recip, int color_space, uchar4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 px = fb[i];
    V3 c = v3(px.x, px.y, px.z) * samples_recip;
    if (color_space == 1) {
        c = (normalize_s(c) + v3(1.0f, 1.0f, 1.0f)) * 0.5f;
    } else if (color_space == 3) {
        c = v3(linear_to_srgb(c.x), linear_to_srgb(c.y), linear_to_srgb(c.z));
    }
    uchar4 o;
    o.x = f32_to_u8(c.x);
    o.y = f32_to_u8(c.y);
    o.z = f32_to_u8(c.z);
    o.w = f32_to_u8(px.w);
    out[i] = o;
}

// FP32 peak probe: FP32_PEAK_CHAINS independent FMA chains per thread.
__global__ void __launch_bounds__(FP32_PEAK_THREADS) fp32_peak_kernel(float* out, uint32_t iters) {
    float a[FP32_PEAK_CHAINS];
    const float m = 1.0f + 1e-7f * (float)threadIdx.x, c = 1e-9f * (float)blockIdx.x;
#pragma unroll
    for (int j = 0; j < FP32_PEAK_CHAINS; ++j) a[j] = (float)j;
#pragma unroll 1
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < FP32_PEAK_UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < FP32_PEAK_CHAINS; ++j) a[j] = fmaf(a[j], m, c);
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < FP32_PEAK_CHAINS; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
cudaError_t ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

}  // namespace

size_t render_smem_bytes(const RenderParams& p) { return (size_t)p.scene.blob_f4 * sizeof(float4); }

// picks <LENS, EXACT, NL> from the scene header
#define BT_DISPATCH_LENS(KERNEL, GRID, BLOCK, SMEM, STREAM, ...)                                   \
    do {                                                                                           \
        cudaError_t e_;                                                                            \
        const bool exact_ = p.scene.lens_exact != 0;                                               \
        if (p.scene.n_lens == 0) {                                                                 \
            if ((e_ = ensure_smem(KERNEL<false, false, 0>, SMEM)) != cudaSuccess) return e_;       \
            KERNEL<false, false, 0><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                   \
        } else if (p.scene.n_lens == 1 && !exact_) {                                               \
            if ((e_ = ensure_smem(KERNEL<true, false, 1>, SMEM)) != cudaSuccess) return e_;        \
            KERNEL<true, false, 1><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                    \
        } else if (!exact_) {                                                                      \
            if ((e_ = ensure_smem(KERNEL<true, false, 0>, SMEM)) != cudaSuccess) return e_;        \
            KERNEL<true, false, 0><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                    \
        } else {                                                                                   \
            if ((e_ = ensure_smem(KERNEL<true, true, 0>, SMEM)) != cudaSuccess) return e_;         \
            KERNEL<true, true, 0><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                     \
        }                                                                                          \
    } while (0)

cudaError_t launch_render(const RenderParams& p, cudaStream_t stream, uint64_t* launches) {
    dim3 grid((p.width + 15) / 16, (p.height + 15) / 16), block(256);
    size_t smem = render_smem_bytes(p);
    BT_DISPATCH_LENS(render_kernel, grid, block, smem, stream, p);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_trace(const RenderParams& p, uint32_t n, const float* origins, const float* dirs,
                         DeviceSegment* out, cudaStream_t stream, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    size_t smem = render_smem_bytes(p);
    BT_DISPATCH_LENS(trace_kernel, (n + 255) / 256, 256, smem, stream, p, n, origins, dirs, out);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_camera_rays(const RenderParams& p, uint32_t n, const uint32_t* xs, const uint32_t* ys,
                               const uint64_t* path_index, float* out, cudaStream_t stream, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    camera_rays_kernel<<<(n + 255) / 256, 256, 0, stream>>>(p, n, xs, ys, path_index, out);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_integrate(const IntegrateParams& p, cudaStream_t stream, uint64_t* launches) {
    if (p.n == 0) return cudaSuccess;
    size_t smem = (size_t)p.n_lens * LENS_STRIDE * sizeof(float4);
    cudaError_t e;
    const unsigned grid = (p.n + 255) / 256;
#define BT_INTEGRATE(EXACT, NL)                                                                \
    do {                                                                                       \
        if ((e = ensure_smem(integrate_kernel<EXACT, NL>, smem)) != cudaSuccess) return e;     \
        integrate_kernel<EXACT, NL><<<grid, 256, smem, stream>>>(p);                           \
    } while (0)
    if (p.exact) BT_INTEGRATE(true, 0);
    else if (p.n_lens == 1) BT_INTEGRATE(false, 1);
    else if (p.n_lens == 2) BT_INTEGRATE(false, 2);
    else if (p.n_lens == 4) BT_INTEGRATE(false, 4);
    else BT_INTEGRATE(false, 0);
#undef BT_INTEGRATE
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_resolve(const float4* fb, uint32_t n_pixels, uint64_t samples, int color_space, uchar4* out,
                           cudaStream_t stream, uint64_t* launches) {
    if (n_pixels == 0) return cudaSuccess;
    float samples_recip = 1.0f / (float)samples;  // buffer.rs:124
    resolve_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(fb, n_pixels, samples_recip, color_space, out);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_fp32_peak(float* out, uint32_t iters, int blocks, cudaStream_t stream, uint64_t* launches) {
    fp32_peak_kernel<<<blocks, FP32_PEAK_THREADS, 0, stream>>>(out, iters);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace bt
