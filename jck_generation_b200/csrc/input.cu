// Device-side input pipeline: the reference's per-sample CPU transform
//     tt.Resize(size) -> tt.ToTensor() -> tt.Normalize(mean, std)        preprocess/dcgan_data_preprocessor.py:38-49
//     OneHotEncoder                                                      preprocess/cgan_data_preprocessor.py:11-16
// on a uint8 dataset that lives in HBM.  Bit-exact with Pillow's bilinear resample (two separable passes, 22-bit fixed-point
// coefficients from the host, rounding to uint8 after each pass) and torchvision's fp32 ToTensor / Normalize arithmetic.
#include "common.cuh"

namespace jck {
namespace {

constexpr int kPrecisionBits = 22;   // Pillow Resample.c PRECISION_BITS = 32 - 8 - 2

struct Norm4 { float mean[4]; float std[4]; };

__device__ __forceinline__ uint8_t clip8(int acc) {
    const int v = acc >> kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// one block per output image: gather image index[b] (or b), horizontal pass -> shared memory, vertical pass + ToTensor +
// Normalize -> NCHW fp32, written x-fastest (coalesced)
__global__ void __launch_bounds__(256)
u8_resize_norm_kernel(const uint8_t* __restrict__ data, const long long* __restrict__ index, float* __restrict__ out, int Hi, int Wi,
                      int C, int Ho, int Wo, const int* __restrict__ hb, const int* __restrict__ hk, int hks,
                      const int* __restrict__ vb, const int* __restrict__ vk, int vks, Norm4 nm) {
    pdl_entry();
    extern __shared__ uint8_t sm[];
    uint8_t* src = sm;                                   // [Hi][Wi][C]
    uint8_t* mid = sm + ((Hi * Wi * C + 15) & ~15);      // [Hi][Wo][C]
    const int b = blockIdx.x;
    const long long img = index ? index[b] : (long long)b;
    const uint8_t* g = data + img * (long long)Hi * Wi * C;
    const int nsrc = Hi * Wi * C;
    if ((nsrc & 15) == 0 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
        for (int i = threadIdx.x; i < nsrc / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(src)[i] = __ldg(reinterpret_cast<const uint4*>(g) + i);
    } else {
        for (int i = threadIdx.x; i < nsrc; i += blockDim.x) src[i] = g[i];
    }
    __syncthreads();
    const uint8_t* h = src;
    if (Wo != Wi) {
        for (int i = threadIdx.x; i < Hi * Wo * C; i += blockDim.x) {
            const int c = i % C, xx = (i / C) % Wo, y = i / (C * Wo);
            const int x0 = hb[2 * xx], n = hb[2 * xx + 1];
            int acc = 1 << (kPrecisionBits - 1);
            for (int k = 0; k < n; ++k) acc += (int)src[(y * Wi + x0 + k) * C + c] * hk[xx * hks + k];
            mid[i] = clip8(acc);
        }
        __syncthreads();
        h = mid;
    }
    float* o = out + (long long)b * C * Ho * Wo;
    for (int i = threadIdx.x; i < C * Ho * Wo; i += blockDim.x) {
        const int xx = i % Wo, yy = (i / Wo) % Ho, c = i / (Wo * Ho);
        uint8_t u;
        if (Ho != Hi) {
            const int y0 = vb[2 * yy], n = vb[2 * yy + 1];
            int acc = 1 << (kPrecisionBits - 1);
            for (int k = 0; k < n; ++k) acc += (int)h[((y0 + k) * Wo + xx) * C + c] * vk[yy * vks + k];
            u = clip8(acc);
        } else {
            u = h[(yy * Wo + xx) * C + c];
        }
        // ToTensor: uint8 -> float / 255;  Normalize: (x - mean) / std   (IEEE fp32 ops, no contraction)
        const float x = __fdiv_rn((float)u, 255.f);
        o[i] = __fdiv_rn(__fsub_rn(x, nm.mean[c]), nm.std[c]);
    }
}

__global__ void __launch_bounds__(256)
one_hot_kernel(const long long* __restrict__ labels, const long long* __restrict__ index, long long* __restrict__ out, int B, int n) {
    pdl_entry();
    const long long total = (long long)B * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / n), j = (int)(i % n);
        const long long l = labels[index ? index[b] : (long long)b];
        out[i] = (l == j) ? 1 : 0;
    }
}

}  // namespace
}  // namespace jck

using namespace jck;

extern "C" int jck_u8_resize_norm(const void* data_u8, const long long* index, float* out_nchw, int B, int Hi, int Wi, int C, int Ho,
                                  int Wo, const int* h_bounds, const int* h_coef, int h_ksize, const int* v_bounds, const int* v_coef,
                                  int v_ksize, const float* mean, const float* std, void* stream) {
    JCK_REQUIRE(data_u8 && out_nchw && mean && std && B > 0 && Hi > 0 && Wi > 0 && C > 0 && C <= 4 && Ho > 0 && Wo > 0,
                "u8_resize_norm: bad argument (C <= 4; mean / std are host pointers to C floats)");
    JCK_REQUIRE((Wo == Wi || (h_bounds && h_coef && h_ksize > 0)) && (Ho == Hi || (v_bounds && v_coef && v_ksize > 0)),
                "u8_resize_norm: coefficient tables missing");
    const size_t smem = ((size_t)Hi * Wi * C + 15) / 16 * 16 + (Wo != Wi ? (size_t)Hi * Wo * C : 0);
    if (smem > 200 * 1024) return set_error(JCK_E_UNSUPPORTED_SHAPE, "u8_resize_norm: image too large for shared memory (%zu bytes)", smem);
    static size_t cfg = 48 * 1024;
    if (smem > cfg) {
        cudaError_t e = cudaFuncSetAttribute(u8_resize_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "u8_resize_norm smem attr: %s", cudaGetErrorString(e));
        cfg = 200 * 1024;
    }
    Norm4 nm;
    for (int c = 0; c < 4; ++c) { nm.mean[c] = c < C ? mean[c] : 0.f; nm.std[c] = c < C ? std[c] : 1.f; }
    launch_pdl(u8_resize_norm_kernel, dim3(B), dim3(256), smem, as_stream(stream), (const uint8_t*)data_u8, index, out_nchw, Hi, Wi, C, Ho,
               Wo, h_bounds, h_coef, h_ksize, v_bounds, v_coef, v_ksize, nm);
    JCK_LAUNCH_CHECK("u8_resize_norm");
    return JCK_OK;
}

extern "C" int jck_one_hot_i64(const long long* labels, const long long* index, long long* out, int B, int n_classes, void* stream) {
    JCK_REQUIRE(labels && out && B > 0 && n_classes > 0, "one_hot: bad argument");
    long long blocks = ((long long)B * n_classes + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    launch_pdl(one_hot_kernel, dim3((int)blocks), dim3(256), 0, as_stream(stream), labels, index, out, B, n_classes);
    JCK_LAUNCH_CHECK("one_hot");
    return JCK_OK;
}
