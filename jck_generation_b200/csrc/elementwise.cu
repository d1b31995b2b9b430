// HBM-bound kernels of the train step: image edge conversion, weight packing, BatchNorm finalize /
// apply / backward, the DCGAN head (conv5 dot product + sigmoid + BCE), generator output edge,
// gradient-penalty norm, fused Adam, Philox random numbers.  All are streaming passes: 128-bit
// accesses on the NHWC tensors, fp32 math, grid sized in multiples of the SM count.
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace jck {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("JCK_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
    long long b = (work_items + threads - 1) / threads;
    long long cap = (long long)kNumSMs * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- 8-wide channel vectors ------------------------------------------------------------------
template <typename T> struct Vec8;
template <> struct Vec8<float> {
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw load_raw(const float* p) {
        Raw r;
        r.a = __ldg(reinterpret_cast<const float4*>(p));
        r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        return r;
    }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
        v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
    }
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Vec8<__nv_bfloat16> {
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
    static __device__ __forceinline__ void unpack(const Raw& u, float (&v)[8]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};

// ---- Philox4x32-10 ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.f / 16777216.f); }  // (0,1)

// Four N(0,1) / U[0,1) values of counter c of stream `stream_id` -- the ONE definition of the random streams: the
// stand-alone generator kernels and the kernels that draw their noise in registers both call these.
__device__ __forceinline__ void philox_normal4(unsigned long long seed, unsigned long long stream_id, unsigned long long c,
                                               float o[4]) {
    const uint4 r = philox(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)),
                           make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float r0 = sqrtf(-2.f * logf(u01(r.x))), r1 = sqrtf(-2.f * logf(u01(r.z)));
    float s0, c0, s1, c1;
    sincospif(2.f * u01(r.y), &s0, &c0);
    sincospif(2.f * u01(r.w), &s1, &c1);
    o[0] = r0 * c0; o[1] = r0 * s0; o[2] = r1 * c1; o[3] = r1 * s1;
}
__device__ __forceinline__ void philox_uniform4(unsigned long long seed, unsigned long long stream_id, unsigned long long c,
                                                float o[4]) {
    const uint4 r = philox(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)),
                           make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    o[0] = (float)(r.x >> 8) * (1.f / 16777216.f); o[1] = (float)(r.y >> 8) * (1.f / 16777216.f);
    o[2] = (float)(r.z >> 8) * (1.f / 16777216.f); o[3] = (float)(r.w >> 8) * (1.f / 16777216.f);
}

// Where a kernel's Gaussian noise comes from: a tensor in memory (parity mode: host-made draws are injected), or
// drawn in registers from stream (seed, stream_id) at *counter + flat_index / 4 (what jck_randn would have written).
struct NoiseSrc {
    const float* mem;
    unsigned long long seed, stream_id;
    const unsigned long long* counter;
    int rng;
};

// ---- image edge ------------------------------------------------------------------------------
// Activation-side image layouts.  JCK_IMG_NHWC: dense [B][H][W][C].  JCK_IMG_P4: [B][H+2][W+2][4] with a
// zero border and zero pad channels -- the layout whose 4x4 patches are TMA-addressable (conv_tc.cu).
struct ImgLayout {
    int H, W, cs, pad;
    __host__ __device__ ImgLayout(int H_, int W_, int C, int layout)
        : H(H_), W(W_), cs(layout == JCK_IMG_P4 ? 4 : C), pad(layout == JCK_IMG_P4 ? 1 : 0) {}
    __device__ __forceinline__ size_t off(int n, int hw) const {
        const int h = hw / W, w = hw - h * W;
        return (((size_t)n * (H + 2 * pad) + h + pad) * (W + 2 * pad) + w + pad) * cs;
    }
};

template <typename T>
__global__ void prep_image_kernel(const float* __restrict__ x1, const float* __restrict__ m1, float a1, float b1,
                                  const float* __restrict__ x2, const float* __restrict__ alpha,
                                  T* __restrict__ out_nhwc, float* __restrict__ out_nchw, int B, int C, int HW,
                                  const ImgLayout lay) {
    pdl_entry();
    const long long total = (long long)B * HW;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(pix / HW), hw = (int)(pix % HW);
        const float al = alpha ? alpha[n] : 1.f;
        const size_t o = lay.off(n, hw);
        for (int c = 0; c < C; ++c) {
            const size_t src = ((size_t)n * C + c) * HW + hw;
            float v = a1 * x1[src];
            if (m1) v += b1 * m1[src];
            if (alpha) v = al * v + (1.f - al) * x2[src];
            if (out_nhwc) st_act(out_nhwc + o + c, v);
            if (out_nchw) out_nchw[src] = v;
        }
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int B, int C, int HW,
                                    const ImgLayout lay) {
    pdl_entry();
    const long long total = (long long)B * HW;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(pix / HW), hw = (int)(pix % HW);
        const size_t o = lay.off(n, hw);
        for (int c = 0; c < C; ++c) out[((size_t)n * C + c) * HW + hw] = ld_act(in + o + c);
    }
}


// Image-edge kernels, quad form (W % 4 == 0, C <= 4): a thread owns 4 consecutive pixels of a row -- one 128-bit
// access per channel on the NCHW fp32 side, 8-byte pixel records (all 4 channels, pad = 0) on the P4 side, and one
// Philox call per channel when the noise is drawn in registers.
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
template <typename T>
__device__ __forceinline__ void store_quad(T* out, const ImgLayout& lay, int n, int hw0, int C, const float v[4][4]) {
    const size_t o = lay.off(n, hw0);
    if constexpr (sizeof(T) == 2) {
        if (lay.pad) {      // P4: 4 pixels x (c0 c1 c2 c3) bf16
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint2 u;
                u.x = pack_bf16x2_rn(v[0][j], v[1][j]);
                u.y = pack_bf16x2_rn(v[2][j], v[3][j]);
                *reinterpret_cast<uint2*>(out + o + 4 * j) = u;
            }
            return;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        for (int c = 0; c < C; ++c) st_act(out + o + (size_t)j * lay.cs + c, v[c][j]);
}
template <typename T>
__device__ __forceinline__ void load_quad(const T* in, const ImgLayout& lay, int n, int hw0, int C, float v[4][4]) {
    const size_t o = lay.off(n, hw0);
    if constexpr (sizeof(T) == 2) {
        if (lay.pad) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint2 u = *reinterpret_cast<const uint2*>(in + o + 4 * j);
                const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
                const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
                v[0][j] = __low2float(lo); v[1][j] = __high2float(lo);
                v[2][j] = __low2float(hi); v[3][j] = __high2float(hi);
            }
            return;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        for (int c = 0; c < C; ++c) v[c][j] = ld_act(in + o + (size_t)j * lay.cs + c);
}
__device__ __forceinline__ void noise_quad(const NoiseSrc& ns, unsigned long long base, size_t src, float m[4]) {
    if (ns.rng) {
        philox_normal4(ns.seed, ns.stream_id, base + (unsigned long long)(src >> 2), m);
    } else {
        const float4 t = *reinterpret_cast<const float4*>(ns.mem + src);
        m[0] = t.x; m[1] = t.y; m[2] = t.z; m[3] = t.w;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
prep_image_quad_kernel(const float* __restrict__ x1, const NoiseSrc ns, float a1, float b1, const float* __restrict__ x2,
                       const float* __restrict__ alpha, T* __restrict__ out_nhwc, float* __restrict__ out_nchw, int B, int C,
                       int HW, const ImgLayout lay) {
    pdl_entry();
    const int qpi = HW / 4;
    const long long total = (long long)B * qpi;
    const bool noisy = ns.rng || ns.mem;
    const unsigned long long base = (ns.rng && ns.counter) ? *ns.counter : 0ull;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(q / qpi), hw0 = (int)(q % qpi) * 4;
        const float al = alpha ? alpha[n] : 1.f;
        float v[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < C) {
                const size_t src = ((size_t)n * C + c) * HW + hw0;
                const float4 x = *reinterpret_cast<const float4*>(x1 + src);
                const float xs[4] = {x.x, x.y, x.z, x.w};
                float m[4] = {0.f, 0.f, 0.f, 0.f}, y[4] = {0.f, 0.f, 0.f, 0.f};
                if (noisy) noise_quad(ns, base, src, m);
                if (alpha) {
                    const float4 t = *reinterpret_cast<const float4*>(x2 + src);
                    y[0] = t.x; y[1] = t.y; y[2] = t.z; y[3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float r = a1 * xs[j];
                    if (noisy) r += b1 * m[j];
                    if (alpha) r = al * r + (1.f - al) * y[j];
                    v[c][j] = r;
                }
                if (out_nchw) *reinterpret_cast<float4*>(out_nchw + src) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[c][j] = 0.f;
            }
        }
        if (out_nhwc) store_quad(out_nhwc, lay, n, hw0, C, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
g_out_fwd_quad_kernel(const T* __restrict__ y5, const NoiseSrc ns, float a, float b, float* __restrict__ fake_raw,
                      float* __restrict__ fake_mix, T* __restrict__ mix_nhwc, int B, int C, int HW, const ImgLayout lay) {
    pdl_entry();
    const int qpi = HW / 4;
    const long long total = (long long)B * qpi;
    const bool noisy = ns.rng || ns.mem;
    const unsigned long long base = (ns.rng && ns.counter) ? *ns.counter : 0ull;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(q / qpi), hw0 = (int)(q % qpi) * 4;
        float v[4][4], mx[4][4];
        load_quad(y5, lay, n, hw0, C, v);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < C) {
                const size_t dst = ((size_t)n * C + c) * HW + hw0;
                float m[4] = {0.f, 0.f, 0.f, 0.f};
                if (noisy) noise_quad(ns, base, dst, m);
                float t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    t[j] = tanhf(v[c][j]);
                    mx[c][j] = noisy ? a * t[j] + b * m[j] : a * t[j];
                }
                if (fake_raw) *reinterpret_cast<float4*>(fake_raw + dst) = make_float4(t[0], t[1], t[2], t[3]);
                if (fake_mix) *reinterpret_cast<float4*>(fake_mix + dst) = make_float4(mx[c][0], mx[c][1], mx[c][2], mx[c][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) mx[c][j] = 0.f;
            }
        }
        if (mix_nhwc) store_quad(mix_nhwc, lay, n, hw0, C, mx);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
g_out_bwd_quad_kernel(const T* __restrict__ dmix, const float* __restrict__ fake_raw, float a, T* __restrict__ dy5, int B,
                      int C, int HW, const ImgLayout lay) {
    pdl_entry();
    const int qpi = HW / 4;
    const long long total = (long long)B * qpi;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(q / qpi), hw0 = (int)(q % qpi) * 4;
        float v[4][4];
        load_quad(dmix, lay, n, hw0, C, v);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < C) {
                const float4 t4 = *reinterpret_cast<const float4*>(fake_raw + ((size_t)n * C + c) * HW + hw0);
                const float t[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) v[c][j] = a * v[c][j] * (1.f - t[j] * t[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[c][j] = 0.f;
            }
        }
        store_quad(dy5, lay, n, hw0, C, v);
    }
}

// ---- weight packing --------------------------------------------------------------------------
template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ w4, T* __restrict__ w_down, T* __restrict__ w_up, int Ca,
                                    int Cb) {
    pdl_entry();
    const long long total = (long long)Ca * Cb * 16;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % 16);
        const long long ab = idx / 16;
        const int b = (int)(ab % Cb), a = (int)(ab / Cb);
        const float v = w4[idx];
        if (w_down) st_act(w_down + ((size_t)a * 16 + tap) * Cb + b, v);
        if (w_up) {
            const int ky = tap >> 2, kx = tap & 3;
            // which output parity / tap slot uses this kernel row: see up_k() in common.cuh
            const int py = (ky == 1 || ky == 3) ? 0 : 1, ty = (ky == 1 || ky == 2) ? 0 : 1;
            const int px = (kx == 1 || kx == 3) ? 0 : 1, tx = (kx == 1 || kx == 2) ? 0 : 1;
            const int phase = py * 2 + px, t = ty * 2 + tx;
            st_act(w_up + (((size_t)phase * Cb + b) * 4 + t) * Ca + a, v);
        }
    }
}

// The same packing for the trunk layers (Ca % 16 == 0, Cb % 32 == 0) through a shared-memory tile of 16 a x 32 b x
// 16 taps: w4 is read as contiguous 2 KB runs, w_down written as 64-byte runs over b, w_up as 32-byte runs over a
// (the element-wise version above scatters 2-byte stores Cb / Ca elements apart).
constexpr int kPackAT = 16, kPackBT = 32;
template <typename T>
__global__ void __launch_bounds__(256)
pack_weights_tiled_kernel(const float* __restrict__ w4, T* __restrict__ w_down, T* __restrict__ w_up, int Ca, int Cb) {
    pdl_entry();
    __shared__ float tile[kPackAT][16 * (kPackBT + 1) + 1];   // [a][tap * 33 + b] (+1: odd row stride), 33 KB
    const int btiles = Cb / kPackBT;
    for (int blk = blockIdx.x; blk < (Ca / kPackAT) * btiles; blk += gridDim.x) {
        const int a0 = (blk / btiles) * kPackAT, b0 = (blk % btiles) * kPackBT;
        for (int i = threadIdx.x; i < kPackAT * 512; i += 256) {   // per a: 32 b x 16 taps = 512 contiguous floats
            const int al = i >> 9, r = i & 511;
            tile[al][(r & 15) * (kPackBT + 1) + (r >> 4)] = w4[((size_t)(a0 + al) * Cb + b0) * 16 + r];
        }
        __syncthreads();
        if (w_down) {
            for (int i = threadIdx.x; i < kPackAT * 16 * kPackBT; i += 256) {
                const int bl = i & 31, tap = (i >> 5) & 15, al = i >> 9;
                st_act(w_down + ((size_t)(a0 + al) * 16 + tap) * Cb + b0 + bl, tile[al][tap * (kPackBT + 1) + bl]);
            }
        }
        if (w_up) {
            for (int i = threadIdx.x; i < kPackAT * 16 * kPackBT; i += 256) {
                const int al = i & 15, tap = (i >> 4) & 15, bl = i >> 8;
                const int ky = tap >> 2, kx = tap & 3;
                const int py = (ky == 1 || ky == 3) ? 0 : 1, ty = (ky == 1 || ky == 2) ? 0 : 1;
                const int px = (kx == 1 || kx == 3) ? 0 : 1, tx = (kx == 1 || kx == 2) ? 0 : 1;
                st_act(w_up + (((size_t)(py * 2 + px) * Cb + b0 + bl) * 4 + (ty * 2 + tx)) * Ca + a0 + al,
                       tile[al][tap * (kPackBT + 1) + bl]);
            }
        }
        __syncthreads();
    }
}

// Image-edge layers (Cb = nc <= 4 image channels) for the tcgen05 path:
//   w_down_e[a][ky*16 + kx*4 + c]                 one 64-wide K step = the whole 4x4x(4) patch
//   w_up9[(py*2+px)*4 + c][s*Ca + a], s = 3x3 input shift (di+1)*3+(dj+1); zero where output parity
//   (py,px) does not read that shift (see up_k / up_d in common.cuh)
__device__ __forceinline__ int edge_k_of(int parity, int d) {   // kernel index for (output parity, input shift) or -1
    if (parity == 0) return d == 0 ? 1 : (d == -1 ? 3 : -1);
    return d == 0 ? 2 : (d == 1 ? 0 : -1);
}
__global__ void pack_weights_edge_kernel(const float* __restrict__ w4, __nv_bfloat16* __restrict__ w_down_e,
                                         __nv_bfloat16* __restrict__ w_up9, int Ca, int nc) {
    pdl_entry();
    const int n_down = Ca * 64, n_up = 16 * 9 * Ca;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_down + n_up; idx += gridDim.x * blockDim.x) {
        if (idx < n_down) {
            if (!w_down_e) continue;
            const int k = idx % 64, a = idx / 64;
            const int ky = k >> 4, kx = (k >> 2) & 3, c = k & 3;
            const float v = c < nc ? w4[(((size_t)a * nc + c) * 4 + ky) * 4 + kx] : 0.f;
            w_down_e[idx] = __float2bfloat16_rn(v);
        } else {
            if (!w_up9) continue;
            const int j = idx - n_down;
            const int k = j % (9 * Ca), n = j / (9 * Ca);
            const int a = k % Ca, s = k / Ca;
            const int di = s / 3 - 1, dj = s % 3 - 1;
            const int c = n & 3, px = (n >> 2) & 1, py = n >> 3;
            const int ky = edge_k_of(py, di), kx = edge_k_of(px, dj);
            const float v = (c < nc && ky >= 0 && kx >= 0) ? w4[(((size_t)a * nc + c) * 4 + ky) * 4 + kx] : 0.f;
            w_up9[j] = __float2bfloat16_rn(v);
        }
    }
}

template <typename T>
__global__ void pack_fc_kernel(const float* __restrict__ w4, T* __restrict__ w_fc, int K, int C) {
    pdl_entry();
    const long long total = (long long)K * C * 16;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % 16);
        const long long kc = idx / 16;
        const int c = (int)(kc % C), k = (int)(kc / C);
        st_act(w_fc + ((size_t)tap * C + c) * K + k, w4[idx]);
    }
}
__global__ void unpack_fc_grad_kernel(const float* __restrict__ dw_fc, float* __restrict__ dw4, int K, int C, int accumulate) {
    pdl_entry();
    const long long total = (long long)K * C * 16;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % 16);
        const long long kc = idx / 16;
        const int c = (int)(kc % C), k = (int)(kc / C);
        const float v = dw_fc[((size_t)tap * C + c) * K + k];
        dw4[idx] = accumulate ? dw4[idx] + v : v;
    }
}
// G.conv1 as a tcgen05 GEMM: the weight as an MN-major operand w_t[k][n = tap*C + c] (bf16), its gradient back from
// dw_t[k][n] (fp32), and the fp32 input rows cast to bf16 with the row pitch padded to a multiple of 8 elements (TMA).
__global__ void pack_fc_t_kernel(const float* __restrict__ w4, __nv_bfloat16* __restrict__ w_t, int K, int C) {
    pdl_entry();
    const long long total = (long long)K * C * 16;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(o % C);
        const long long kt = o / C;
        const int tap = (int)(kt % 16), k = (int)(kt / 16);
        w_t[o] = __float2bfloat16_rn(w4[((size_t)k * C + c) * 16 + tap]);       // coalesced writes, 64-byte-strided reads (L2 hits)
    }
}
__global__ void unpack_fc_grad_t_kernel(const float* __restrict__ dw_t, float* __restrict__ dw4, int K, int C, int accumulate) {
    pdl_entry();
    const long long total = (long long)K * C * 16;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(idx % 16);
        const long long kc = idx / 16;
        const int c = (int)(kc % C), k = (int)(kc / C);
        const float v = dw_t[((size_t)k * 16 + tap) * C + c];
        dw4[idx] = accumulate ? dw4[idx] + v : v;
    }
}
// out[m][0:K1] = a[m][:], out[m][K1:K1+K2] = b[m][:] (b fp32, or int64 when b_is_i64), zero up to ldo; bf16 or fp32 output
template <typename TO>
__global__ void concat_rows_kernel(const float* __restrict__ a, const void* __restrict__ b, TO* __restrict__ out, int M, int K1,
                                   int K2, int ldo, int b_is_i64) {
    pdl_entry();
    const long long total = (long long)M * ldo;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(o % ldo);
        const long long m = o / ldo;
        float v = 0.f;
        if (k < K1) v = a[m * K1 + k];
        else if (k < K1 + K2)
            v = b_is_i64 ? (float)reinterpret_cast<const long long*>(b)[m * K2 + (k - K1)]
                         : reinterpret_cast<const float*>(b)[m * K2 + (k - K1)];
        st_act(out + o, v);
    }
}
__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
    pdl_entry();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void cast_rows_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int M, int K, int ldo) {
    pdl_entry();
    const long long total = (long long)M * ldo;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(o % ldo);
        const long long m = o / ldo;
        out[o] = __float2bfloat16_rn(k < K ? x[m * K + k] : 0.f);
    }
}
template <typename T>
__global__ void pack_head_kernel(const float* __restrict__ w4, T* __restrict__ w5, int C4) {
    pdl_entry();
    const int total = C4 * 16;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int tap = idx % 16, c = idx / 16;
        st_act(w5 + tap * C4 + c, w4[idx]);
    }
}
__global__ void unpack_head_grad_kernel(const float* __restrict__ dw5, float* __restrict__ dw4, int C4, int accumulate) {
    pdl_entry();
    const int total = C4 * 16;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int tap = idx % 16, c = idx / 16;
        const float v = dw5[tap * C4 + c];
        dw4[idx] = accumulate ? dw4[idx] + v : v;
    }
}

// ---- BatchNorm -------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ nbt,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_rstd, int C, int groups,
                                   float count, float eps, float momentum) {
    pdl_entry();
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
        for (int g = 0; g < groups; ++g) {   // in order: the reference updates running stats pass by pass
            const double s1 = stats[(size_t)g * 2 * C + c], s2 = stats[(size_t)g * 2 * C + C + c];
            const double mean = s1 / count;
            double var = s2 / count - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)eps));
            const float sc = gamma[c] * rstd;
            scale_shift[(size_t)g * 2 * C + c] = sc;
            scale_shift[(size_t)g * 2 * C + C + c] = beta[c] - (float)mean * sc;
            mean_rstd[(size_t)g * 2 * C + c] = (float)mean;
            mean_rstd[(size_t)g * 2 * C + C + c] = rstd;
            const float unbiased = (float)(count > 1.f ? var * (double)count / ((double)count - 1.0) : var);
            rm = (1.f - momentum) * rm + momentum * (float)mean;
            rv = (1.f - momentum) * rv + momentum * unbiased;
        }
        if (running_mean) running_mean[c] = rm;
        if (running_var) running_var[c] = rv;
    }
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += groups;
}

// The three streaming BatchNorm passes share one thread mapping: a thread owns ONE 8-channel vector
// (cvi = tid % (C/8)) and walks pixels of one statistics group (blockIdx.y), so every per-channel
// coefficient lives in registers for the whole kernel.  The passes are pure HBM streams, so what matters is
// memory-level parallelism: the main loop issues ALL 128-bit loads of kBnUnroll rows (both tensors, raw,
// unconverted) before touching any of them -- no bounds test, no conversion between the loads -- and a
// separate tail loop takes the last < kBnUnroll rows.  Rows of one thread are 256/(C/8) pixels apart, i.e.
// always 2048 elements, so the unrolled addresses are immediates off one running pointer.
constexpr int kBnUnroll = 4;
constexpr int kBnRowElems = 2048;    // (256 / (C/8)) * C

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const T* __restrict__ y, const float* __restrict__ scale_shift, T* __restrict__ a, int C,
                  unsigned pix_per_group, unsigned slab, float slope) {
    pdl_entry();
    typedef typename Vec8<T>::Raw Raw;
    const unsigned cv = C / 8, rows = 256 / cv;
    const unsigned myc = threadIdx.x % cv, myr = threadIdx.x / cv;
    const unsigned g = blockIdx.y, c0 = myc * 8;
    const unsigned p_beg = blockIdx.x * slab, p_end = min(pix_per_group, p_beg + slab);
    float sc[8], sh[8];
    Vec8<float>::load(scale_shift + (size_t)g * 2 * C + c0, sc);
    Vec8<float>::load(scale_shift + (size_t)g * 2 * C + C + c0, sh);
    unsigned pp = p_beg + myr;
    const size_t off = ((size_t)g * pix_per_group + pp) * C + c0;
    const T* py = y + off;
    T* pa = a + off;
    auto one = [&](const Raw& r, T* dst) {
        float v[8];
        Vec8<T>::unpack(r, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float pre = fmaf(v[j], sc[j], sh[j]);
            v[j] = pre > 0.f ? pre : pre * slope;
        }
        Vec8<T>::store(dst, v);
    };
    for (; pp + (kBnUnroll - 1) * rows < p_end; pp += kBnUnroll * rows, py += kBnUnroll * kBnRowElems, pa += kBnUnroll * kBnRowElems) {
        Raw r[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) r[u] = Vec8<T>::load_raw(py + u * kBnRowElems);
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) one(r[u], pa + u * kBnRowElems);
    }
    for (; pp < p_end; pp += rows, py += kBnRowElems, pa += kBnRowElems) one(Vec8<T>::load_raw(py), pa);
}

// sums[g][0:C] += sum g,  sums[g][C:2C] += sum g*xhat, g = da * act'(pre)
template <typename T>
__global__ void __launch_bounds__(256, 4)
bn_act_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ y, const float* __restrict__ scale_shift,
                         const float* __restrict__ mean_rstd, float* __restrict__ sums, int C, unsigned pix_per_group,
                         unsigned slab, float slope) {
    pdl_entry();
    typedef typename Vec8<T>::Raw Raw;
    __shared__ float red[256][17];
    const unsigned cv = C / 8, rows = 256 / cv;
    const unsigned myc = threadIdx.x % cv, myr = threadIdx.x / cv;
    const unsigned g = blockIdx.y, c0 = myc * 8;
    const unsigned p_beg = blockIdx.x * slab, p_end = min(pix_per_group, p_beg + slab);
    float sc[8], sh[8], mu[8];
    Vec8<float>::load(scale_shift + (size_t)g * 2 * C + c0, sc);
    Vec8<float>::load(scale_shift + (size_t)g * 2 * C + C + c0, sh);
    Vec8<float>::load(mean_rstd + (size_t)g * 2 * C + c0, mu);
    float s1[8], s2[8];                    // s2 accumulates g * (y - mean); rstd is applied once at the end
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    unsigned pp = p_beg + myr;
    const size_t off = ((size_t)g * pix_per_group + pp) * C + c0;
    const T* pd = da + off;
    const T* py = y + off;
    auto one = [&](const Raw& rd, const Raw& ry) {
        float d[8], v[8];
        Vec8<T>::unpack(rd, d);
        Vec8<T>::unpack(ry, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float pre = fmaf(v[j], sc[j], sh[j]);
            const float gg = pre > 0.f ? d[j] : d[j] * slope;
            s1[j] += gg;
            s2[j] = fmaf(gg, v[j] - mu[j], s2[j]);
        }
    };
    for (; pp + (kBnUnroll - 1) * rows < p_end; pp += kBnUnroll * rows, pd += kBnUnroll * kBnRowElems, py += kBnUnroll * kBnRowElems) {
        Raw rd[kBnUnroll], ry[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            rd[u] = Vec8<T>::load_raw(pd + u * kBnRowElems);
            ry[u] = Vec8<T>::load_raw(py + u * kBnRowElems);
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) one(rd[u], ry[u]);
    }
    for (; pp < p_end; pp += rows, pd += kBnRowElems, py += kBnRowElems) one(Vec8<T>::load_raw(pd), Vec8<T>::load_raw(py));
    {
        float rs[8];
        Vec8<float>::load(mean_rstd + (size_t)g * 2 * C + C + c0, rs);
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s1[j]; red[threadIdx.x][8 + j] = s2[j] * rs[j]; }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < 2 * C; col += 256) {
        const int which = col / C, ch = col % C;
        const int vc = ch / 8, j = ch % 8;
        float s = 0.f;
        for (unsigned r = 0; r < rows; ++r) s += red[r * cv + vc][which * 8 + j];
        atomicAdd(sums + (size_t)g * 2 * C + which * C + ch, s);
    }
}

// dy = gamma*rstd*(g - sum_g/N - xhat*sum_gx/N) = k1*g - k3*y + k4,  k1 = gamma*rstd, k3 = k1*rstd*sum_gx/N,
// k4 = k3*mean - k1*sum_g/N
template <typename T>
__global__ void __launch_bounds__(256, 4)
bn_act_bwd_apply_kernel(const T* __restrict__ da, const T* __restrict__ y, const float* __restrict__ scale_shift,
                        const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                        const float* __restrict__ sums, T* __restrict__ dy, int C, unsigned pix_per_group,
                        unsigned slab, float inv_count, float slope) {
    pdl_entry();
    typedef typename Vec8<T>::Raw Raw;
    const unsigned cv = C / 8, rows = 256 / cv;
    const unsigned myc = threadIdx.x % cv, myr = threadIdx.x / cv;
    const unsigned g = blockIdx.y, c0 = myc * 8;
    const unsigned p_beg = blockIdx.x * slab, p_end = min(pix_per_group, p_beg + slab);
    float sc[8], sh[8], k1[8], k3[8], k4[8];
    {
        float mu[8], rs[8], ga[8], sg[8], sgx[8];
        const size_t gofs = (size_t)g * 2 * C;
        Vec8<float>::load(scale_shift + gofs + c0, sc);
        Vec8<float>::load(scale_shift + gofs + C + c0, sh);
        Vec8<float>::load(mean_rstd + gofs + c0, mu);
        Vec8<float>::load(mean_rstd + gofs + C + c0, rs);
        Vec8<float>::load(gamma + c0, ga);
        Vec8<float>::load(sums + gofs + c0, sg);
        Vec8<float>::load(sums + gofs + C + c0, sgx);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            k1[j] = ga[j] * rs[j];
            k3[j] = k1[j] * rs[j] * sgx[j] * inv_count;
            k4[j] = k3[j] * mu[j] - k1[j] * sg[j] * inv_count;
        }
    }
    unsigned pp = p_beg + myr;
    const size_t off = ((size_t)g * pix_per_group + pp) * C + c0;
    const T* pd = da + off;
    const T* py = y + off;
    T* po = dy + off;
    auto one = [&](const Raw& rd, const Raw& ry, T* dst) {
        float d[8], v[8];
        Vec8<T>::unpack(rd, d);
        Vec8<T>::unpack(ry, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float pre = fmaf(v[j], sc[j], sh[j]);
            const float gg = pre > 0.f ? d[j] : d[j] * slope;
            d[j] = fmaf(k1[j], gg, fmaf(-k3[j], v[j], k4[j]));
        }
        Vec8<T>::store(dst, d);
    };
    for (; pp + (kBnUnroll - 1) * rows < p_end;
         pp += kBnUnroll * rows, pd += kBnUnroll * kBnRowElems, py += kBnUnroll * kBnRowElems, po += kBnUnroll * kBnRowElems) {
        Raw rd[kBnUnroll], ry[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            rd[u] = Vec8<T>::load_raw(pd + u * kBnRowElems);
            ry[u] = Vec8<T>::load_raw(py + u * kBnRowElems);
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) one(rd[u], ry[u], po + u * kBnRowElems);
    }
    for (; pp < p_end; pp += rows, pd += kBnRowElems, py += kBnRowElems, po += kBnRowElems)
        one(Vec8<T>::load_raw(pd), Vec8<T>::load_raw(py), po);
}

// ---- coherent-front variants for tensors larger than the L2 -------------------------------------------------------
// The slab kernels above give every block its own contiguous slab: 600-900 parallel sweeps, so what is left in the
// 126 MB L2 when the kernel ends is the tail of EVERY slab, and what the producer left there (the last ~40-50 MB a
// convolution wrote, measured as the gap between its output size and its DRAM write bytes) has been evicted by the time
// a slab reaches it.  Here all blocks walk the tensor together, one 16 KB chunk (kBnUnroll row steps) per block per
// iteration, grid-strided, ascending or DESCENDING: a kernel that starts where its producer stopped reads the
// producer's tail from L2, and leaves its own last output where a consumer walking the other way starts.  The train
// step alternates directions along each chain (convolutions ascend; bn_act_fwd and bn_act_bwd_reduce descend;
// bn_act_bwd_apply ascends after a reduce pass and descends after a fused convolution).  Loads of tensors that are dead
// after the pass (or not needed again before the L2 has turned over) carry L2::evict_first so they do not displace the
// output.  Groups are contiguous pixel ranges, so a block reloads its per-channel coefficients when its chunk crosses
// into another group (and the reduce pass flushes its partial sums there).  Results are identical to the slab kernels.
template <typename T> struct Vec8Stream;
template <> struct Vec8Stream<float> {
    static __device__ __forceinline__ typename Vec8<float>::Raw load(const float* p) { return Vec8<float>::load_raw(p); }
};
template <> struct Vec8Stream<__nv_bfloat16> {
    static __device__ __forceinline__ uint4 load(const __nv_bfloat16* p) {
        uint64_t pol;     // fractional L2 policy: the whole access evict_first (hoisted out of the loop by the compiler)
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        uint4 u;
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p), "l"(pol));
        return u;
    }
};
constexpr int kBnChunkElems = kBnUnroll * kBnRowElems;     // 8192 elements per block per iteration

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_fwd_front_kernel(const T* __restrict__ y, const float* __restrict__ scale_shift, T* __restrict__ a, int C,
                        unsigned chunks_per_group, unsigned total_chunks, int descending, float slope) {
    pdl_entry();
    typedef typename Vec8<T>::Raw Raw;
    const unsigned cv = C / 8;
    const unsigned myc = threadIdx.x % cv, myr = threadIdx.x / cv, c0 = myc * 8;
    float sc[8], sh[8];
    unsigned cur_g = 0xffffffffu;
    for (unsigned i = blockIdx.x; i < total_chunks; i += gridDim.x) {
        const unsigned c = descending ? total_chunks - 1 - i : i;
        const unsigned g = c / chunks_per_group;
        if (g != cur_g) {
            Vec8<float>::load(scale_shift + (size_t)g * 2 * C + c0, sc);
            Vec8<float>::load(scale_shift + (size_t)g * 2 * C + C + c0, sh);
            cur_g = g;
        }
        const size_t off = (size_t)c * kBnChunkElems + (size_t)myr * C + c0;
        Raw r[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) r[u] = Vec8Stream<T>::load(y + off + u * kBnRowElems);   // y: next read in backward
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            float v[8];
            Vec8<T>::unpack(r[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float pre = fmaf(v[j], sc[j], sh[j]);
                v[j] = pre > 0.f ? pre : pre * slope;
            }
            Vec8<T>::store(a + off + u * kBnRowElems, v);
        }
    }
}

template <typename T, int MINB>       // MINB: resident blocks per SM the register budget is cut for (2: 128, 3: 80 registers)
__global__ void __launch_bounds__(256, MINB)
bn_act_bwd_reduce_front_kernel(const T* __restrict__ da, const T* __restrict__ y, const float* __restrict__ scale_shift,
                               const float* __restrict__ mean_rstd, float* __restrict__ sums, int C,
                               unsigned chunks_per_group, unsigned total_chunks, int descending, float slope) {
    pdl_entry();
    typedef typename Vec8<T>::Raw Raw;
    __shared__ float red[256][17];
    const unsigned cv = C / 8, rows = 256 / cv;
    const unsigned myc = threadIdx.x % cv, myr = threadIdx.x / cv, c0 = myc * 8;
    float sc[8], sh[8], mu[8], s1[8], s2[8];
    unsigned cur_g = 0xffffffffu;
    auto flush = [&](unsigned g) {      // block-uniform: every thread of the block changes group in the same iteration
        float rs[8];
        Vec8<float>::load(mean_rstd + (size_t)g * 2 * C + C + c0, rs);
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = s1[j]; red[threadIdx.x][8 + j] = s2[j] * rs[j]; }
        __syncthreads();
        for (int col = threadIdx.x; col < 2 * C; col += 256) {
            const int which = col / C, ch = col % C;
            const int vc = ch / 8, j = ch % 8;
            float s = 0.f;
            for (unsigned r = 0; r < rows; ++r) s += red[r * cv + vc][which * 8 + j];
            atomicAdd(sums + (size_t)g * 2 * C + which * C + ch, s);
        }
        __syncthreads();
    };
    for (unsigned i = blockIdx.x; i < total_chunks; i += gridDim.x) {
        const unsigned c = descending ? total_chunks - 1 - i : i;
        const unsigned g = c / chunks_per_group;
        if (g != cur_g) {
            if (cur_g != 0xffffffffu) flush(cur_g);
            Vec8<float>::load(scale_shift + (size_t)g * 2 * C + c0, sc);
            Vec8<float>::load(scale_shift + (size_t)g * 2 * C + C + c0, sh);
            Vec8<float>::load(mean_rstd + (size_t)g * 2 * C + c0, mu);
#pragma unroll
            for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
            cur_g = g;
        }
        const size_t off = (size_t)c * kBnChunkElems + (size_t)myr * C + c0;
        Raw rd[kBnUnroll], ry[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {        // both are read again by the apply pass that follows: default policy
            rd[u] = Vec8<T>::load_raw(da + off + u * kBnRowElems);
            ry[u] = Vec8<T>::load_raw(y + off + u * kBnRowElems);
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            float d[8], v[8];
            Vec8<T>::unpack(rd[u], d);
            Vec8<T>::unpack(ry[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float pre = fmaf(v[j], sc[j], sh[j]);
                const float gg = pre > 0.f ? d[j] : d[j] * slope;
                s1[j] += gg;
                s2[j] = fmaf(gg, v[j] - mu[j], s2[j]);
            }
        }
    }
    if (cur_g != 0xffffffffu) flush(cur_g);
}

template <typename T, int MINB>
__global__ void __launch_bounds__(256, MINB)
bn_act_bwd_apply_front_kernel(const T* __restrict__ da, const T* __restrict__ y, const float* __restrict__ scale_shift,
                              const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                              const float* __restrict__ sums, T* __restrict__ dy, int C, unsigned chunks_per_group,
                              unsigned total_chunks, int descending, float inv_count, float slope) {
    pdl_entry();
    typedef typename Vec8<T>::Raw Raw;
    const unsigned cv = C / 8;
    const unsigned myc = threadIdx.x % cv, myr = threadIdx.x / cv, c0 = myc * 8;
    float sc[8], sh[8], k1[8], k3[8], k4[8];
    unsigned cur_g = 0xffffffffu;
    for (unsigned i = blockIdx.x; i < total_chunks; i += gridDim.x) {
        const unsigned c = descending ? total_chunks - 1 - i : i;
        const unsigned g = c / chunks_per_group;
        if (g != cur_g) {
            float mu[8], rs[8], ga[8], sg[8], sgx[8];
            const size_t gofs = (size_t)g * 2 * C;
            Vec8<float>::load(scale_shift + gofs + c0, sc);
            Vec8<float>::load(scale_shift + gofs + C + c0, sh);
            Vec8<float>::load(mean_rstd + gofs + c0, mu);
            Vec8<float>::load(mean_rstd + gofs + C + c0, rs);
            Vec8<float>::load(gamma + c0, ga);
            Vec8<float>::load(sums + gofs + c0, sg);
            Vec8<float>::load(sums + gofs + C + c0, sgx);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                k1[j] = ga[j] * rs[j];
                k3[j] = k1[j] * rs[j] * sgx[j] * inv_count;
                k4[j] = k3[j] * mu[j] - k1[j] * sg[j] * inv_count;
            }
            cur_g = g;
        }
        const size_t off = (size_t)c * kBnChunkElems + (size_t)myr * C + c0;
        Raw rd[kBnUnroll], ry[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {        // da and y are dead after this pass: do not let them displace dy
            rd[u] = Vec8Stream<T>::load(da + off + u * kBnRowElems);
            ry[u] = Vec8Stream<T>::load(y + off + u * kBnRowElems);
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            float d[8], v[8];
            Vec8<T>::unpack(rd[u], d);
            Vec8<T>::unpack(ry[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float pre = fmaf(v[j], sc[j], sh[j]);
                const float gg = pre > 0.f ? d[j] : d[j] * slope;
                d[j] = fmaf(k1[j], gg, fmaf(-k3[j], v[j], k4[j]));
            }
            Vec8<T>::store(dy + off + u * kBnRowElems, d);
        }
    }
}

__global__ void bn_param_grad_kernel(const float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int C, int groups, int accumulate) {
    pdl_entry();
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        float sb = 0.f, sg = 0.f;
        for (int g = 0; g < groups; ++g) { sb += sums[(size_t)g * 2 * C + c]; sg += sums[(size_t)g * 2 * C + C + c]; }
        dgamma[c] = accumulate ? dgamma[c] + sg : sg;
        dbeta[c] = accumulate ? dbeta[c] + sb : sb;
    }
}

// ---- DCGAN head ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const T* __restrict__ a4, const T* __restrict__ w5, float* __restrict__ prob, float target,
                float* __restrict__ scalars, int B, int K) {
    pdl_entry();
    __shared__ float wsum[8];
    const int b = blockIdx.x;
    float acc = 0.f;
    for (int k = threadIdx.x * 8; k < K; k += 256 * 8) {
        float x[8], w[8];
        Vec8<T>::load(a4 + (size_t)b * K + k, x);
        Vec8<T>::load(w5 + k, w);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(x[j], w[j], acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float z = 0.f;
        for (int i = 0; i < 8; ++i) z += wsum[i];
        const float p = 1.f / (1.f + expf(-z));
        prob[b] = p;
        if (scalars) {
            // torch BCELoss: -(t*max(log p,-100) + (1-t)*max(log1p(-p),-100)), mean over the batch
            const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);
            atomicAdd(scalars + 0, -(target * lp + (1.f - target) * lq) / (float)B);
            atomicAdd(scalars + 1, p / (float)B);
        }
    }
}

__device__ __forceinline__ float head_dlogit(float p, float dp, float target, int mode, float invB) {
    const float pq = p * (1.f - p);
    if (mode == 1) return pq;
    if (mode == 2) return dp * pq;
    return (p - target) / fmaxf(pq, 1e-12f) * pq * invB;
}

// grid (K/8/256, sample chunks): da4 = dlogit (x) w5 ; dw5 += dlogit^T a4
template <typename T>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ dprob, float target, const T* __restrict__ w5,
                const T* __restrict__ a4,
                T* __restrict__ da4, float* __restrict__ dw5, int B, int K, int mode, int chunk, float invB) {
    pdl_entry();
    const int k = (blockIdx.x * 256 + threadIdx.x) * 8;
    if (k >= K) return;
    float w[8], acc[8];
    Vec8<T>::load(w5 + k, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int b_beg = blockIdx.y * chunk, b_end = min(B, b_beg + chunk);
    for (int b = b_beg; b < b_end; ++b) {
        const float dl = head_dlogit(prob[b], dprob ? dprob[b] : 0.f, target, mode, invB);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = dl * w[j];
        Vec8<T>::store(da4 + (size_t)b * K + k, o);
        if (dw5) {
            float x[8];
            Vec8<T>::load(a4 + (size_t)b * K + k, x);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(dl, x[j], acc[j]);
        }
    }
    if (dw5) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw5 + k + j, acc[j]);
    }
}

// ---- generator output edge -------------------------------------------------------------------
template <typename T>
__global__ void g_out_fwd_kernel(const T* __restrict__ y5, const float* __restrict__ noise, float a, float b,
                                 float* __restrict__ fake_raw, float* __restrict__ fake_mix, T* __restrict__ mix_nhwc,
                                 int B, int C, int HW, const ImgLayout lay) {
    pdl_entry();
    const long long total = (long long)B * HW;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(pix / HW), hw = (int)(pix % HW);
        const size_t o = lay.off(n, hw);
        for (int c = 0; c < C; ++c) {
            const size_t dst = ((size_t)n * C + c) * HW + hw;
            const float t = tanhf(ld_act(y5 + o + c));
            if (fake_raw) fake_raw[dst] = t;
            const float m = noise ? a * t + b * noise[dst] : a * t;
            if (fake_mix) fake_mix[dst] = m;
            if (mix_nhwc) st_act(mix_nhwc + o + c, m);
        }
    }
}
template <typename T>
__global__ void g_out_bwd_kernel(const T* __restrict__ dmix, const float* __restrict__ fake_raw, float a,
                                 T* __restrict__ dy5, int B, int C, int HW, const ImgLayout lay) {
    pdl_entry();
    const long long total = (long long)B * HW;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(pix / HW), hw = (int)(pix % HW);
        const size_t o = lay.off(n, hw);
        for (int c = 0; c < C; ++c) {
            const float t = fake_raw[((size_t)n * C + c) * HW + hw];
            st_act(dy5 + o + c, a * ld_act(dmix + o + c) * (1.f - t * t));
        }
    }
}

// ---- gradient penalty ------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gp_penalty_kernel(const T* __restrict__ dx, float* __restrict__ scalars, int B, long long per_sample) {
    pdl_entry();
    __shared__ float wsum[8];
    const int n = blockIdx.x;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < per_sample; i += 256) {
        const float v = ld_act(dx + (size_t)n * per_sample + i);
        acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += wsum[i];
        const float d = sqrtf(s) - 1.f;
        atomicAdd(scalars, d * d / (float)B);
    }
}

// ---- Adam ------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, const int* __restrict__ step_count) {
    pdl_entry();
    const float t = (float)(*step_count + 1);
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);          // lerp, as torch's _single_tensor_adam
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] -= step_size * (mi / denom);
    }
}
__global__ void adam_advance_kernel(int* step_count) {
    pdl_entry(); *step_count += 1; }

template <bool kNormal>
__global__ void rng_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long stream_id,
                           const unsigned long long* __restrict__ counter_base) {
    pdl_entry();
    const unsigned long long base = counter_base ? *counter_base : 0ull;
    const long long nquads = (n + 3) / 4;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nquads; q += (long long)gridDim.x * blockDim.x) {
        float o[4];
        if (kNormal) philox_normal4(seed, stream_id, base + (unsigned long long)q, o);
        else philox_uniform4(seed, stream_id, base + (unsigned long long)q, o);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (q * 4 + j < n) out[q * 4 + j] = o[j];
    }
}
__global__ void rng_advance_kernel(unsigned long long* c, unsigned long long by) {
    pdl_entry(); *c += by; }

}  // namespace
}  // namespace jck

using namespace jck;

#define DISPATCH_DTYPE(dtype, name, ...)                                        \
    if ((dtype) == JCK_F32) { using T = float; __VA_ARGS__ }                    \
    else if ((dtype) == JCK_BF16) { using T = __nv_bfloat16; __VA_ARGS__ }      \
    else return set_error(JCK_E_BADARG, name ": dtype %d", (int)(dtype));

extern "C" int jck_version(void) { return 100; }
extern "C" const char* jck_last_error_string(void) { return g_err; }
extern "C" unsigned long long jck_launch_count(void) { return g_launches.load(); }

static int prep_image_launch(const float* x1, const NoiseSrc& ns, float a1, float b1, const float* x2, const float* alpha,
                             void* out_nhwc, float* out_nchw_f32, int B, int C, int H, int W, int layout, int dtype,
                             void* stream) {
    JCK_REQUIRE(x1 && (out_nhwc || out_nchw_f32) && B > 0 && C > 0 && H > 0 && W > 0, "prep_image: bad argument");
    JCK_REQUIRE(!alpha || x2, "prep_image: alpha needs x2");
    JCK_REQUIRE(layout == JCK_IMG_NHWC || (layout == JCK_IMG_P4 && C <= 4), "prep_image: bad layout");
    const long long total = (long long)B * H * W;
    const ImgLayout lay(H, W, C, layout);
    if (W % 4 == 0 && C <= 4) {
        DISPATCH_DTYPE(dtype, "prep_image",
            launch_pdl(prep_image_quad_kernel<T>, dim3(grid_for(total / 4, 256)), dim3(256), 0, as_stream(stream), 
                x1, ns, a1, b1, x2, alpha, (T*)out_nhwc, out_nchw_f32, B, C, H * W, lay);)
    } else {
        JCK_REQUIRE(!ns.rng, "prep_image: in-register noise needs W % 4 == 0 and C <= 4");
        DISPATCH_DTYPE(dtype, "prep_image",
            launch_pdl(prep_image_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), x1, ns.mem, a1, b1, x2, alpha, (T*)out_nhwc,
                                                                                      out_nchw_f32, B, C, H * W, lay);)
    }
    JCK_LAUNCH_CHECK("prep_image");
    return JCK_OK;
}

extern "C" int jck_prep_image(const float* x1, const float* m1, float a1, float b1, const float* x2, const float* alpha,
                              void* out_nhwc, float* out_nchw_f32, int B, int C, int H, int W, int layout, int dtype,
                              void* stream) {
    const NoiseSrc ns{m1, 0ull, 0ull, nullptr, 0};
    return prep_image_launch(x1, ns, a1, b1, x2, alpha, out_nhwc, out_nchw_f32, B, C, H, W, layout, dtype, stream);
}

extern "C" int jck_prep_image_rng(const float* x1, unsigned long long seed, unsigned long long stream_id,
                                  const unsigned long long* counter_base, float a1, float b1, void* out_nhwc,
                                  float* out_nchw_f32, int B, int C, int H, int W, int layout, int dtype, void* stream) {
    const NoiseSrc ns{nullptr, seed, stream_id, counter_base, 1};
    return prep_image_launch(x1, ns, a1, b1, nullptr, nullptr, out_nhwc, out_nchw_f32, B, C, H, W, layout, dtype, stream);
}

extern "C" int jck_nhwc_to_nchw_f32(const void* in_nhwc, float* out_nchw, int B, int C, int H, int W, int layout, int dtype,
                                    void* stream) {
    JCK_REQUIRE(in_nhwc && out_nchw && B > 0 && C > 0, "nhwc_to_nchw: bad argument");
    JCK_REQUIRE(layout == JCK_IMG_NHWC || (layout == JCK_IMG_P4 && C <= 4), "nhwc_to_nchw: bad layout");
    const long long total = (long long)B * H * W;
    const ImgLayout lay(H, W, C, layout);
    DISPATCH_DTYPE(dtype, "nhwc_to_nchw",
        launch_pdl(nhwc_to_nchw_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), (const T*)in_nhwc, out_nchw, B, C, H * W, lay);)
    JCK_LAUNCH_CHECK("nhwc_to_nchw");
    return JCK_OK;
}

extern "C" int jck_pack_weights(const float* w4, void* w_down, void* w_up, int Ca, int Cb, int dtype, void* stream) {
    JCK_REQUIRE(w4 && (w_down || w_up) && Ca > 0 && Cb > 0, "pack_weights: bad argument");
    const long long total = (long long)Ca * Cb * 16;
    if (Ca % kPackAT == 0 && Cb % kPackBT == 0) {
        int blocks = (Ca / kPackAT) * (Cb / kPackBT);
        if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
        DISPATCH_DTYPE(dtype, "pack_weights",
            launch_pdl(pack_weights_tiled_kernel<T>, dim3(blocks), dim3(256), 0, as_stream(stream), w4, (T*)w_down, (T*)w_up, Ca, Cb);)
        JCK_LAUNCH_CHECK("pack_weights");
        return JCK_OK;
    }
    DISPATCH_DTYPE(dtype, "pack_weights",
        launch_pdl(pack_weights_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), w4, (T*)w_down, (T*)w_up, Ca, Cb);)
    JCK_LAUNCH_CHECK("pack_weights");
    return JCK_OK;
}

extern "C" int jck_pack_weights_edge(const float* w4, void* w_down_e, void* w_up9, int Ca, int nc, void* stream) {
    JCK_REQUIRE(w4 && (w_down_e || w_up9) && Ca > 0 && nc > 0 && nc <= 4, "pack_weights_edge: bad argument");
    launch_pdl(pack_weights_edge_kernel, dim3(grid_for((long long)Ca * 64 + 144LL * Ca, 256)), dim3(256), 0, as_stream(stream), 
        w4, (__nv_bfloat16*)w_down_e, (__nv_bfloat16*)w_up9, Ca, nc);
    JCK_LAUNCH_CHECK("pack_weights_edge");
    return JCK_OK;
}

extern "C" int jck_pack_fc(const float* w4, void* w_fc, int K, int C, int dtype, void* stream) {
    JCK_REQUIRE(w4 && w_fc && K > 0 && C > 0, "pack_fc: bad argument");
    DISPATCH_DTYPE(dtype, "pack_fc",
        launch_pdl(pack_fc_kernel<T>, dim3(grid_for((long long)K * C * 16, 256)), dim3(256), 0, as_stream(stream), w4, (T*)w_fc, K, C);)
    JCK_LAUNCH_CHECK("pack_fc");
    return JCK_OK;
}
extern "C" int jck_unpack_fc_grad(const float* dw_fc, float* dw4, int K, int C, int accumulate, void* stream) {
    JCK_REQUIRE(dw_fc && dw4 && K > 0 && C > 0, "unpack_fc_grad: bad argument");
    launch_pdl(unpack_fc_grad_kernel, dim3(grid_for((long long)K * C * 16, 256)), dim3(256), 0, as_stream(stream), dw_fc, dw4, K, C, accumulate);
    JCK_LAUNCH_CHECK("unpack_fc_grad");
    return JCK_OK;
}
extern "C" int jck_pack_fc_t(const float* w4, void* w_t_bf16, int K, int C, void* stream) {
    JCK_REQUIRE(w4 && w_t_bf16 && K > 0 && C > 0, "pack_fc_t: bad argument");
    launch_pdl(pack_fc_t_kernel, dim3(grid_for((long long)K * C * 16, 256)), dim3(256), 0, as_stream(stream), w4,
               (__nv_bfloat16*)w_t_bf16, K, C);
    JCK_LAUNCH_CHECK("pack_fc_t");
    return JCK_OK;
}
extern "C" int jck_unpack_fc_grad_t(const float* dw_t, float* dw4, int K, int C, int accumulate, void* stream) {
    JCK_REQUIRE(dw_t && dw4 && K > 0 && C > 0, "unpack_fc_grad_t: bad argument");
    launch_pdl(unpack_fc_grad_t_kernel, dim3(grid_for((long long)K * C * 16, 256)), dim3(256), 0, as_stream(stream), dw_t, dw4, K, C,
               accumulate);
    JCK_LAUNCH_CHECK("unpack_fc_grad_t");
    return JCK_OK;
}
extern "C" int jck_cast_rows_bf16(const float* x, void* out_bf16, int M, int K, int ldo, void* stream) {
    JCK_REQUIRE(x && out_bf16 && M > 0 && K > 0 && ldo >= K, "cast_rows_bf16: bad argument");
    launch_pdl(cast_rows_bf16_kernel, dim3(grid_for((long long)M * ldo, 256)), dim3(256), 0, as_stream(stream), x,
               (__nv_bfloat16*)out_bf16, M, K, ldo);
    JCK_LAUNCH_CHECK("cast_rows_bf16");
    return JCK_OK;
}
extern "C" int jck_concat_rows(const float* a, const void* b, int b_is_i64, void* out, int M, int K1, int K2, int ldo, int dtype,
                               void* stream) {
    JCK_REQUIRE(a && b && out && M > 0 && K1 > 0 && K2 > 0 && ldo >= K1 + K2, "concat_rows: bad argument");
    DISPATCH_DTYPE(dtype, "concat_rows",
        launch_pdl(concat_rows_kernel<T>, dim3(grid_for((long long)M * ldo, 256)), dim3(256), 0, as_stream(stream), a, b, (T*)out, M, K1,
                   K2, ldo, b_is_i64);)
    JCK_LAUNCH_CHECK("concat_rows");
    return JCK_OK;
}
extern "C" int jck_zero(void* p, size_t bytes, void* stream) {
    JCK_REQUIRE(p && bytes > 0, "zero: bad argument");
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, as_stream(stream));
    if (e != cudaSuccess) return set_error(JCK_E_CUDA, "zero: %s", cudaGetErrorString(e));
    return JCK_OK;
}
extern "C" int jck_copy_f32(const float* src, float* dst, long long n, void* stream) {
    JCK_REQUIRE(src && dst && n > 0, "copy_f32: bad argument");
    launch_pdl(copy_f32_kernel, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), src, dst, n);
    JCK_LAUNCH_CHECK("copy_f32");
    return JCK_OK;
}
extern "C" int jck_pack_head(const float* w4, void* w5, int C4, int dtype, void* stream) {
    JCK_REQUIRE(w4 && w5 && C4 > 0, "pack_head: bad argument");
    DISPATCH_DTYPE(dtype, "pack_head",
        launch_pdl(pack_head_kernel<T>, dim3(grid_for(C4 * 16, 256)), dim3(256), 0, as_stream(stream), w4, (T*)w5, C4);)
    JCK_LAUNCH_CHECK("pack_head");
    return JCK_OK;
}
extern "C" int jck_unpack_head_grad(const float* dw5, float* dw4, int C4, int accumulate, void* stream) {
    JCK_REQUIRE(dw5 && dw4 && C4 > 0, "unpack_head_grad: bad argument");
    launch_pdl(unpack_head_grad_kernel, dim3(grid_for(C4 * 16, 256)), dim3(256), 0, as_stream(stream), dw5, dw4, C4, accumulate);
    JCK_LAUNCH_CHECK("unpack_head_grad");
    return JCK_OK;
}

extern "C" int jck_bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, long long* num_batches_tracked, float* scale_shift, float* mean_rstd,
                               int C, int groups, float count, float eps, float momentum, void* stream) {
    JCK_REQUIRE(stats && gamma && beta && scale_shift && mean_rstd && C > 0 && groups > 0 && count > 0, "bn_finalize: bad argument");
    launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, as_stream(stream), stats, gamma, beta, running_mean, running_var,
                                                                       num_batches_tracked, scale_shift, mean_rstd, C,
                                                                       groups, count, eps, momentum);
    JCK_LAUNCH_CHECK("bn_finalize");
    return JCK_OK;
}

static bool bn_c_ok(int C) { return C >= 8 && C % 8 == 0 && (256 % (C / 8)) == 0; }

// grid (blocks per group, groups): `waves` x 148 blocks of 256 threads in total, each thread row gets >= 4 pixels.
// `waves` is chosen per kernel as a multiple of the blocks of that kernel that fit on one SM (registers:
// fwd 40 -> 6; reduce and apply are capped at 64 by __launch_bounds__(256, 4) -> 4), so the grid is a whole number of
// resident waves (no ragged tail)
static dim3 bn_grid(long long pix_per_group, int groups, int C, int waves, unsigned* slab) {
    const int rows = 256 / (C / 8);
    long long bpg = ((long long)waves * kNumSMs + groups - 1) / groups;
    const long long max_b = (pix_per_group + rows * 4 - 1) / (rows * 4);
    if (bpg > max_b) bpg = max_b;
    if (bpg < 1) bpg = 1;
    *slab = (unsigned)((pix_per_group + bpg - 1) / bpg);
    return dim3((unsigned)bpg, (unsigned)groups);
}

// Coherent-front kernels (above): for tensors of >= kBnFrontMinBytes whose groups are whole numbers of 8192-element
// chunks; JCK_ORDER_SLAB (or JCK_BN_FRONT=0 in the environment, for A/B timing) keeps the slabs.  Measured on the 512-image
// step (profiles/r02/run16-18.sh): 2.60 -> 2.51 ms (40-step runs at 1965 MHz), 2.72 -> 2.61 ms (200-step runs under the
// power cap); DRAM traffic per step 6.51 -> 5.65 GB (ncu --cache-control none, profiles/r02_ncu_launches_warmL2_summary.md).
// Threshold 8 / 24 / 48 MB and 2 / 3 resident blocks per SM for the backward passes are within run-to-run noise of each other.
constexpr long long kBnFrontMinBytes = 8LL << 20;
static bool bn_front_enabled() {
    static const bool on = [] { const char* e = getenv("JCK_BN_FRONT"); return !(e && e[0] == '0'); }();
    return on;
}
static bool bn_front(long long npix, int C, long long pix_per_group, int dtype, int order, unsigned* cpg, unsigned* total) {
    if (order == JCK_ORDER_SLAB || !bn_front_enabled()) return false;
    const long long elems = npix * C, per_group = pix_per_group * C;
    const long long bytes = elems * (dtype == JCK_BF16 ? 2 : 4);
    static const long long min_bytes = [] {
        const char* e = getenv("JCK_BN_FRONT_MIN_MB");
        return e ? (long long)atoi(e) << 20 : kBnFrontMinBytes;
    }();
    if (bytes < min_bytes || per_group % kBnChunkElems != 0 || elems / kBnChunkElems >= (1LL << 31)) return false;
    *cpg = (unsigned)(per_group / kBnChunkElems);
    *total = (unsigned)(elems / kBnChunkElems);
    return true;
}
// resident blocks per SM of the two backward passes (JCK_BN_OCC=2|3, default 2): 40 coefficient + 32 data registers per
// thread do not fit the 64 a 4-block budget allows -- ptxas then sinks each load pair to its use and only 2 x 16 B per thread
// are in flight; at 128 registers (2 blocks, 512 threads x 8 x 16 B = 64 KB per SM) every load of a chunk is issued up front
static int bn_front_occ() {
    static const int occ = [] { const char* e = getenv("JCK_BN_OCC"); return (e && e[0] == '3') ? 3 : 2; }();
    return occ;
}
static unsigned bn_front_grid(unsigned total_chunks, int waves) {
    const unsigned g = (unsigned)(waves * kNumSMs);
    return total_chunks < g ? total_chunks : g;
}

#define BN_COMMON_CHECKS(name)                                                                                   \
    JCK_REQUIRE(npix > 0 && pix_per_group > 0 && npix % pix_per_group == 0 && npix < (1LL << 31), name ": bad size"); \
    if (!bn_c_ok(C)) return set_error(JCK_E_UNSUPPORTED_SHAPE, name ": C=%d (need C/8 dividing 256)", C);

extern "C" int jck_bn_act_fwd(const void* y, const float* scale_shift, void* a, long long npix, int C, long long pix_per_group,
                              float slope, int dtype, int order, void* stream) {
    JCK_REQUIRE(y && scale_shift && a, "bn_act_fwd: bad argument");
    BN_COMMON_CHECKS("bn_act_fwd")
    unsigned cpg, total;
    if (bn_front(npix, C, pix_per_group, dtype, order, &cpg, &total)) {
        const unsigned grid = bn_front_grid(total, 6);
        DISPATCH_DTYPE(dtype, "bn_act_fwd",
            launch_pdl(bn_act_fwd_front_kernel<T>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)y, scale_shift, (T*)a, C, cpg,
                       total, (int)(order == JCK_ORDER_DESC), slope);)
        JCK_LAUNCH_CHECK("bn_act_fwd");
        return JCK_OK;
    }
    unsigned slab;
    const dim3 grid = bn_grid(pix_per_group, (int)(npix / pix_per_group), C, 6, &slab);
    DISPATCH_DTYPE(dtype, "bn_act_fwd",
        launch_pdl(bn_act_fwd_kernel<T>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)y, scale_shift, (T*)a, C, (unsigned)pix_per_group,
                                                                slab, slope);)
    JCK_LAUNCH_CHECK("bn_act_fwd");
    return JCK_OK;
}

extern "C" int jck_bn_act_bwd_reduce(const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                                     float* sums, long long npix, int C, long long pix_per_group, float slope, int dtype,
                                     int order, void* stream) {
    JCK_REQUIRE(da && y && scale_shift && mean_rstd && sums, "bn_act_bwd_reduce: bad argument");
    BN_COMMON_CHECKS("bn_act_bwd_reduce")
    unsigned cpg, total;
    if (bn_front(npix, C, pix_per_group, dtype, order, &cpg, &total)) {
        const int occ = bn_front_occ();
        const unsigned grid = bn_front_grid(total, occ);
        if (occ == 2) {
            DISPATCH_DTYPE(dtype, "bn_act_bwd_reduce",
                launch_pdl(bn_act_bwd_reduce_front_kernel<T, 2>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)da, (const T*)y,
                           scale_shift, mean_rstd, sums, C, cpg, total, (int)(order == JCK_ORDER_DESC), slope);)
        } else {
            DISPATCH_DTYPE(dtype, "bn_act_bwd_reduce",
                launch_pdl(bn_act_bwd_reduce_front_kernel<T, 3>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)da, (const T*)y,
                           scale_shift, mean_rstd, sums, C, cpg, total, (int)(order == JCK_ORDER_DESC), slope);)
        }
        JCK_LAUNCH_CHECK("bn_act_bwd_reduce");
        return JCK_OK;
    }
    unsigned slab;
    const dim3 grid = bn_grid(pix_per_group, (int)(npix / pix_per_group), C, 4, &slab);
    DISPATCH_DTYPE(dtype, "bn_act_bwd_reduce",
        launch_pdl(bn_act_bwd_reduce_kernel<T>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)da, (const T*)y, scale_shift, mean_rstd,
                                                                       sums, C, (unsigned)pix_per_group, slab, slope);)
    JCK_LAUNCH_CHECK("bn_act_bwd_reduce");
    return JCK_OK;
}

extern "C" int jck_bn_act_bwd_apply(const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                                    const float* gamma, const float* sums, void* dy, long long npix, int C,
                                    long long pix_per_group, float count, float slope, int dtype, int order, void* stream) {
    JCK_REQUIRE(da && y && scale_shift && mean_rstd && gamma && sums && dy && count > 0, "bn_act_bwd_apply: bad argument");
    BN_COMMON_CHECKS("bn_act_bwd_apply")
    unsigned cpg, total;
    if (bn_front(npix, C, pix_per_group, dtype, order, &cpg, &total)) {
        const int occ = bn_front_occ();
        const unsigned grid = bn_front_grid(total, occ);
        if (occ == 2) {
            DISPATCH_DTYPE(dtype, "bn_act_bwd_apply",
                launch_pdl(bn_act_bwd_apply_front_kernel<T, 2>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)da, (const T*)y,
                           scale_shift, mean_rstd, gamma, sums, (T*)dy, C, cpg, total, (int)(order == JCK_ORDER_DESC), 1.f / count,
                           slope);)
        } else {
            DISPATCH_DTYPE(dtype, "bn_act_bwd_apply",
                launch_pdl(bn_act_bwd_apply_front_kernel<T, 3>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)da, (const T*)y,
                           scale_shift, mean_rstd, gamma, sums, (T*)dy, C, cpg, total, (int)(order == JCK_ORDER_DESC), 1.f / count,
                           slope);)
        }
        JCK_LAUNCH_CHECK("bn_act_bwd_apply");
        return JCK_OK;
    }
    unsigned slab;
    const dim3 grid = bn_grid(pix_per_group, (int)(npix / pix_per_group), C, 4, &slab);
    DISPATCH_DTYPE(dtype, "bn_act_bwd_apply",
        launch_pdl(bn_act_bwd_apply_kernel<T>, dim3(grid), dim3(256), 0, as_stream(stream), (const T*)da, (const T*)y, scale_shift, mean_rstd,
                                                                      gamma, sums, (T*)dy, C, (unsigned)pix_per_group, slab,
                                                                      1.f / count, slope);)
    JCK_LAUNCH_CHECK("bn_act_bwd_apply");
    return JCK_OK;
}

extern "C" int jck_bn_param_grad(const float* sums, float* dgamma, float* dbeta, int C, int groups, int accumulate, void* stream) {
    JCK_REQUIRE(sums && dgamma && dbeta && C > 0 && groups > 0, "bn_param_grad: bad argument");
    launch_pdl(bn_param_grad_kernel, dim3((C + 127) / 128), dim3(128), 0, as_stream(stream), sums, dgamma, dbeta, C, groups, accumulate);
    JCK_LAUNCH_CHECK("bn_param_grad");
    return JCK_OK;
}

extern "C" int jck_head_fwd(const void* a4, const void* w5, float* prob, float target, float* scalars, int B, int K,
                            int dtype, void* stream) {
    JCK_REQUIRE(a4 && w5 && prob && B > 0 && K > 0 && K % 8 == 0, "head_fwd: bad argument");
    DISPATCH_DTYPE(dtype, "head_fwd",
        launch_pdl(head_fwd_kernel<T>, dim3(B), dim3(256), 0, as_stream(stream), (const T*)a4, (const T*)w5, prob, target, scalars, B, K);)
    JCK_LAUNCH_CHECK("head_fwd");
    return JCK_OK;
}

extern "C" int jck_head_bwd(const float* prob, const float* dprob, float target, const void* w5, const void* a4, void* da4,
                            float* dw5, int B, int mean_count, int K, int mode, int accumulate, int dtype, void* stream) {
    JCK_REQUIRE(prob && w5 && da4 && B > 0 && mean_count >= 0 && K > 0 && K % 8 == 0 && (!dw5 || a4) && (mode != 2 || dprob),
                "head_bwd: bad argument");
    const float invB = 1.f / (float)(mean_count > 0 ? mean_count : B);
    cudaStream_t st = as_stream(stream);
    if (dw5 && !accumulate) {
        cudaError_t e = cudaMemsetAsync(dw5, 0, sizeof(float) * K, st);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "head_bwd memset: %s", cudaGetErrorString(e));
    }
    int chunks = B < 64 ? B : 64;
    const int chunk = (B + chunks - 1) / chunks;
    chunks = (B + chunk - 1) / chunk;
    dim3 grid((K / 8 + 255) / 256, chunks);
    DISPATCH_DTYPE(dtype, "head_bwd",
        launch_pdl(head_bwd_kernel<T>, dim3(grid), dim3(256), 0, st, prob, dprob, target, (const T*)w5, (const T*)a4, (T*)da4, dw5, B, K, mode, chunk, invB);)
    JCK_LAUNCH_CHECK("head_bwd");
    return JCK_OK;
}

static int g_out_fwd_launch(const void* y5_nhwc, const NoiseSrc& ns, float a, float b, float* fake_raw_nchw,
                            float* fake_mix_nchw, void* fake_mix_nhwc, int B, int C, int H, int W, int layout, int dtype,
                            void* stream) {
    JCK_REQUIRE(y5_nhwc && B > 0 && C > 0 && H > 0 && W > 0, "g_out_fwd: bad argument");
    JCK_REQUIRE(layout == JCK_IMG_NHWC || (layout == JCK_IMG_P4 && C <= 4), "g_out_fwd: bad layout");
    const long long total = (long long)B * H * W;
    const ImgLayout lay(H, W, C, layout);
    if (W % 4 == 0 && C <= 4) {
        DISPATCH_DTYPE(dtype, "g_out_fwd",
            launch_pdl(g_out_fwd_quad_kernel<T>, dim3(grid_for(total / 4, 256)), dim3(256), 0, as_stream(stream), 
                (const T*)y5_nhwc, ns, a, b, fake_raw_nchw, fake_mix_nchw, (T*)fake_mix_nhwc, B, C, H * W, lay);)
    } else {
        JCK_REQUIRE(!ns.rng, "g_out_fwd: in-register noise needs W % 4 == 0 and C <= 4");
        DISPATCH_DTYPE(dtype, "g_out_fwd",
            launch_pdl(g_out_fwd_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), (const T*)y5_nhwc, ns.mem, a, b, fake_raw_nchw,
                                                                                   fake_mix_nchw, (T*)fake_mix_nhwc, B, C, H * W, lay);)
    }
    JCK_LAUNCH_CHECK("g_out_fwd");
    return JCK_OK;
}

extern "C" int jck_g_out_fwd(const void* y5_nhwc, const float* noise, float a, float b, float* fake_raw_nchw,
                             float* fake_mix_nchw, void* fake_mix_nhwc, int B, int C, int H, int W, int layout, int dtype,
                             void* stream) {
    const NoiseSrc ns{noise, 0ull, 0ull, nullptr, 0};
    return g_out_fwd_launch(y5_nhwc, ns, a, b, fake_raw_nchw, fake_mix_nchw, fake_mix_nhwc, B, C, H, W, layout, dtype, stream);
}

extern "C" int jck_g_out_fwd_rng(const void* y5_nhwc, unsigned long long seed, unsigned long long stream_id,
                                 const unsigned long long* counter_base, float a, float b, float* fake_raw_nchw,
                                 float* fake_mix_nchw, void* fake_mix_nhwc, int B, int C, int H, int W, int layout, int dtype,
                                 void* stream) {
    const NoiseSrc ns{nullptr, seed, stream_id, counter_base, 1};
    return g_out_fwd_launch(y5_nhwc, ns, a, b, fake_raw_nchw, fake_mix_nchw, fake_mix_nhwc, B, C, H, W, layout, dtype, stream);
}

extern "C" int jck_g_out_bwd(const void* dmix_nhwc, const float* fake_raw_nchw, float a, void* dy5_nhwc, int B, int C, int H,
                             int W, int layout, int dtype, void* stream) {
    JCK_REQUIRE(dmix_nhwc && fake_raw_nchw && dy5_nhwc && B > 0 && C > 0, "g_out_bwd: bad argument");
    JCK_REQUIRE(layout == JCK_IMG_NHWC || (layout == JCK_IMG_P4 && C <= 4), "g_out_bwd: bad layout");
    const long long total = (long long)B * H * W;
    const ImgLayout lay(H, W, C, layout);
    if (W % 4 == 0 && C <= 4) {
        DISPATCH_DTYPE(dtype, "g_out_bwd",
            launch_pdl(g_out_bwd_quad_kernel<T>, dim3(grid_for(total / 4, 256)), dim3(256), 0, as_stream(stream), 
                (const T*)dmix_nhwc, fake_raw_nchw, a, (T*)dy5_nhwc, B, C, H * W, lay);)
    } else {
        DISPATCH_DTYPE(dtype, "g_out_bwd",
            launch_pdl(g_out_bwd_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), (const T*)dmix_nhwc, fake_raw_nchw, a,
                                                                                   (T*)dy5_nhwc, B, C, H * W, lay);)
    }
    JCK_LAUNCH_CHECK("g_out_bwd");
    return JCK_OK;
}

extern "C" int jck_gp_penalty(const void* dx, float* scalars, int B, long long per_sample, int dtype, void* stream) {
    JCK_REQUIRE(dx && scalars && B > 0 && per_sample > 0, "gp_penalty: bad argument");
    DISPATCH_DTYPE(dtype, "gp_penalty",
        launch_pdl(gp_penalty_kernel<T>, dim3(B), dim3(256), 0, as_stream(stream), (const T*)dx, scalars, B, per_sample);)
    JCK_LAUNCH_CHECK("gp_penalty");
    return JCK_OK;
}

extern "C" int jck_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                        float beta1, float beta2, float eps, const int* step_count, void* stream) {
    JCK_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step_count, "adam: bad argument");
    launch_pdl(adam_kernel, dim3(grid_for(n, 256)), dim3(256), 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                 step_count);
    JCK_LAUNCH_CHECK("adam");
    return JCK_OK;
}
extern "C" int jck_adam_advance(int* step_count, void* stream) {
    JCK_REQUIRE(step_count, "adam_advance: bad argument");
    launch_pdl(adam_advance_kernel, dim3(1), dim3(1), 0, as_stream(stream), step_count);
    JCK_LAUNCH_CHECK("adam_advance");
    return JCK_OK;
}

extern "C" int jck_randn(float* out, long long n, unsigned long long seed, unsigned long long stream_id,
                         const unsigned long long* counter_base, void* stream) {
    JCK_REQUIRE(out && n > 0, "randn: bad argument");
    launch_pdl(rng_kernel<true>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), out, n, seed, stream_id, counter_base);
    JCK_LAUNCH_CHECK("randn");
    return JCK_OK;
}
extern "C" int jck_rand(float* out, long long n, unsigned long long seed, unsigned long long stream_id,
                        const unsigned long long* counter_base, void* stream) {
    JCK_REQUIRE(out && n > 0, "rand: bad argument");
    launch_pdl(rng_kernel<false>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), out, n, seed, stream_id, counter_base);
    JCK_LAUNCH_CHECK("rand");
    return JCK_OK;
}
extern "C" int jck_rng_advance(unsigned long long* counter_base, unsigned long long by, void* stream) {
    JCK_REQUIRE(counter_base, "rng_advance: bad argument");
    launch_pdl(rng_advance_kernel, dim3(1), dim3(1), 0, as_stream(stream), counter_base, by);
    JCK_LAUNCH_CHECK("rng_advance");
    return JCK_OK;
}
