// CGAN-specific kernels: the small fp32 pieces of the discriminator head (label embedding,
// Linear 8392 -> 256 -> Dropout -> Linear 256 -> 1 -> Sigmoid; model/CGAN.py:83-84,103-107,112-122) and
// the second-order sweep the CGAN discriminator update needs because its gradient penalty IS
// back-propagated (train/cgan_trainer.py:200-204): the adjoint of the BatchNorm backward pass, the
// gradient-penalty seed, and the logit second derivative.  Products with the 8192-wide feature vector go
// through jck_dense (conv_simt.cu); everything here is a streaming pass.
#include "common.cuh"

namespace jck {
namespace {

inline int grid1d(long long n, int threads = 256) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

__global__ void rowop_kernel(int op, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                             int M, int N, int rows_y, float s) {
    pdl_entry();
    const long long total = (op == 4) ? (long long)rows_y * N : (long long)M * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N), m = (int)(i / N);
        float v;
        switch (op) {
            case 0: { v = x[i] + (y ? y[n] : 0.f); v = v > 0.f ? v : v * s; break; }
            case 1: v = x[i] * y[i] * s; break;
            case 2: v = x[i] * (y[i] > 0.f ? 1.f : s); break;
            case 3: v = x[i] + y[(size_t)(m % rows_y) * N + n]; break;
            case 4: { v = 0.f; for (int g = 0; g * rows_y < M; ++g) v += x[((size_t)g * rows_y + m) * N + n]; break; }
            case 5: v = x[m] * y[n]; break;
            case 6: v = x[i] * y[m]; break;
            default: v = x[i] >= s ? 1.f : 0.f; break;   // 7: keep-mask from uniforms
        }
        out[i] = v;
    }
}

__global__ void sigmoid_bce_kernel(const float* __restrict__ logit, float* __restrict__ prob, float target,
                                   float* __restrict__ scalars, int B) {
    pdl_entry();
    float l = 0.f, ps = 0.f;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float p = 1.f / (1.f + expf(-logit[b]));
        prob[b] = p;
        if (scalars) {
            l += -(target * fmaxf(logf(p), -100.f) + (1.f - target) * fmaxf(log1pf(-p), -100.f));
            ps += p;
        }
    }
    if (scalars) {
        l = warp_sum(l); ps = warp_sum(ps);
        if ((threadIdx.x & 31) == 0) { atomicAdd(scalars + 0, l / (float)B); atomicAdd(scalars + 1, ps / (float)B); }
    }
}

// d(loss)/d(logit).  mode 0: BCE mean through the sigmoid; 1: ones on the sigmoid output (GP sweep);
// 2: caller-supplied d/d(prob); 3: second order, up = adjoint of g_s = sigma'(s): d/ds sigma'(s) = pq(1-2p).
__global__ void logit_grad_kernel(const float* __restrict__ prob, const float* __restrict__ up, float target,
                                  float* __restrict__ out, int B, int mode, float scale) {
    pdl_entry();
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float p = prob[b], pq = p * (1.f - p);
        float v;
        if (mode == 0) v = (p - target) / fmaxf(pq, 1e-12f) * pq * scale;
        else if (mode == 1) v = pq * scale;
        else if (mode == 2) v = up[b] * pq * scale;
        else v = up[b] * pq * (1.f - 2.f * p) * scale;
        out[b] = v;
    }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
    pdl_entry();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}
// FID feature moments (metrics.py:118-124): centred features split into two bf16 terms, f - mean = hi + lo with
// |lo| <= 2^-9 |hi|, so that the Gram matrix hi^T hi + hi^T lo + lo^T hi on the bf16 tensor cores carries ~2^-17 relative
// error per product instead of 2^-9.  Output rows have pitch ldo >= d (zero padded: TMA wants 16-byte row pitches).
__global__ void center_split_kernel(const float* __restrict__ x, const float* __restrict__ mean, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, long long rows, int d, int ldo) {
    pdl_entry();
    const long long total = rows * ldo;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(o % ldo);
        const long long r = o / ldo;
        float v = 0.f;
        if (c < d) v = x[r * d + c] - mean[c];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[o] = h;
        lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}
__global__ void i64_to_f32_kernel(const long long* __restrict__ in, float* __restrict__ out, long long n) {
    pdl_entry();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

template <typename T>
__global__ void axpy_kernel(const T* __restrict__ x, T* __restrict__ y, float a, long long n) {
    pdl_entry();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        st_act(y + i, ld_act(y + i) + a * ld_act(x + i));
}

// one block per sample: norm = ||v_b||; scalars[0] += (norm-1)^2 / B; u_b = scale * (1 - 1/norm) * v_b
// (= d/dv_b of lambda * mean_b (||v_b|| - 1)^2 with scale = 2*lambda/B)
template <typename T>
__global__ void __launch_bounds__(256)
gp_seed_kernel(const T* __restrict__ v, T* __restrict__ u, float* __restrict__ scalars, int B, long long per_sample,
               float scale) {
    pdl_entry();
    __shared__ float wsum[8];
    __shared__ float coef;
    const size_t base = (size_t)blockIdx.x * per_sample;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < per_sample; i += 256) { const float t = ld_act(v + base + i); acc = fmaf(t, t, acc); }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += wsum[i];
        const float norm = sqrtf(s), d = norm - 1.f;
        if (scalars) atomicAdd(scalars, d * d / (float)B);
        coef = norm > 0.f ? scale * d / norm : 0.f;
    }
    __syncthreads();
    const float c = coef;
    if (u) for (long long i = threadIdx.x; i < per_sample; i += 256) st_act(u + base + i, c * ld_act(v + base + i));
}

// linear1.weight [O][C*HW + E] (feature columns in the reference's NCHW flatten order, model/CGAN.py:118)
// <-> w_a [O][HW*C] (NHWC order, activation dtype) and w_b [O][E] fp32
template <typename T>
__global__ void pack_linear_kernel(const float* __restrict__ w, T* __restrict__ wa, float* __restrict__ wb, int O, int C,
                                   int HW, int E) {
    pdl_entry();
    const int F = C * HW, ld = F + E;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)O * ld; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % ld), o = (int)(i / ld);
        if (col < F) { const int c = col / HW, hw = col % HW; st_act(wa + (size_t)o * F + hw * C + c, w[i]); }
        else wb[(size_t)o * E + (col - F)] = w[i];
    }
}
__global__ void unpack_linear_grad_kernel(const float* __restrict__ dwa, const float* __restrict__ dwb, float* __restrict__ dw,
                                          int O, int C, int HW, int E, int accumulate) {
    pdl_entry();
    const int F = C * HW, ld = F + E;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)O * ld; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % ld), o = (int)(i / ld);
        float v;
        if (col < F) { const int c = col / HW, hw = col % HW; v = dwa[(size_t)o * F + hw * C + c]; }
        else v = dwb[(size_t)o * E + (col - F)];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

// ---- adjoint of BatchNorm backward -------------------------------------------------------------
// Forward-of-backward being differentiated (per channel, N = count):
//     g = da * act'(pre),  d = (gamma/sigma) * (g - mean(g) - xhat * mean(g*xhat))          [bn_act_bwd_apply]
// Given the adjoint dbar of d:
//     S1 = sum dbar, S2 = sum dbar*xhat, S3 = sum dbar*r,   r = g - mean(g) - xhat*mean(g*xhat)
//     gbar  = c * (dbar - S1/N - xhat*S2/N),  c = gamma/sigma          (then * act'(pre) -> adjoint of da)
//     gamma_bar = S3 / sigma
//     xhat_bar = -c * (q*dbar + (S2/N)*g),  q = mean(g*xhat)
//     ybar  = (1/sigma) * (xhat_bar - T1/N - xhat*T2/N) - (gamma*S3/(N*sigma^2)) * xhat,
//             T1 = sum xhat_bar = -c*(q*S1 + (S2/N)*sum g),  T2 = sum xhat_bar*xhat = -2*c*q*S2
// ybar is the second-order gradient that re-enters the ordinary backward sweep at this layer's raw output.
template <typename T>
__global__ void __launch_bounds__(256)
bn_adj_reduce_kernel(const T* __restrict__ dbar, const T* __restrict__ da, const T* __restrict__ y,
                     const float* __restrict__ ss, const float* __restrict__ mr, const float* __restrict__ sums1,
                     float* __restrict__ asums, int C, long long npix, float inv_count, float slope) {
    pdl_entry();
    // one thread per (pixel-slab, channel): simple and exact; this pass only runs on the GP group
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = ss[c], sh = ss[C + c], mu = mr[c], rs = mr[C + c];
    const float mg = sums1[c] * inv_count, q = sums1[C + c] * inv_count;
    const long long slab = (npix + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * slab, p1 = min(npix, p0 + slab);
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (long long p = p0; p < p1; ++p) {
        const size_t i = (size_t)p * C + c;
        const float yv = ld_act(y + i), db = ld_act(dbar + i), xh = (yv - mu) * rs;
        const float g = fmaf(yv, sc, sh) > 0.f ? ld_act(da + i) : ld_act(da + i) * slope;
        s1 += db; s2 += db * xh; s3 += db * (g - mg - xh * q);
    }
    atomicAdd(asums + c, s1); atomicAdd(asums + C + c, s2); atomicAdd(asums + 2 * C + c, s3);
}

template <typename T>
__global__ void bn_adj_apply_kernel(const T* __restrict__ dbar, const T* __restrict__ da, const T* __restrict__ y,
                                    const float* __restrict__ ss, const float* __restrict__ mr,
                                    const float* __restrict__ gamma, const float* __restrict__ sums1,
                                    const float* __restrict__ asums, T* __restrict__ gbar_a, T* __restrict__ ybar,
                                    int C, long long npix, float inv_count, float slope) {
    pdl_entry();
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float sc = ss[c], sh = ss[C + c], mu = mr[c], rs = mr[C + c], ga = gamma[c];
        const float sg = sums1[c], q = sums1[C + c] * inv_count;
        const float S1 = asums[c], S2 = asums[C + c], S3 = asums[2 * C + c];
        const float cc = ga * rs;
        const float yv = ld_act(y + i), db = ld_act(dbar + i), xh = (yv - mu) * rs;
        const float mask = fmaf(yv, sc, sh) > 0.f ? 1.f : slope;
        const float g = ld_act(da + i) * mask;
        const float gb = cc * (db - S1 * inv_count - xh * S2 * inv_count);
        st_act(gbar_a + i, gb * mask);
        const float xb = -cc * (q * db + S2 * inv_count * g);
        const float T1 = -cc * (q * S1 + S2 * inv_count * sg), T2 = -2.f * cc * q * S2;
        st_act(ybar + i, rs * (xb - T1 * inv_count - xh * T2 * inv_count) - ga * S3 * rs * rs * inv_count * xh);
    }
}

__global__ void bn_adj_param_kernel(const float* __restrict__ asums, const float* __restrict__ mr, float* __restrict__ dgamma,
                                    int C, float scale) {
    pdl_entry();
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x)
        dgamma[c] += scale * asums[2 * C + c] * mr[C + c];
}

}  // namespace
}  // namespace jck

using namespace jck;

#define DISPATCH_DTYPE(dtype, name, ...)                                        \
    if ((dtype) == JCK_F32) { using T = float; __VA_ARGS__ }                    \
    else if ((dtype) == JCK_BF16) { using T = __nv_bfloat16; __VA_ARGS__ }      \
    else return set_error(JCK_E_BADARG, name ": dtype %d", (int)(dtype));

extern "C" int jck_rowop(int op, const float* x, const float* y, float* out, int M, int N, int rows_y, float s, void* stream) {
    JCK_REQUIRE(x && out && M > 0 && N > 0 && op >= 0 && op <= 7 && (y || op == 0 || op == 7), "rowop: bad argument");
    JCK_REQUIRE((op != 3 && op != 4) || (rows_y > 0 && M % rows_y == 0), "rowop: rows_y must divide M");
    const long long total = op == 4 ? (long long)rows_y * N : (long long)M * N;
    launch_pdl(rowop_kernel, dim3(grid1d(total)), dim3(256), 0, as_stream(stream), op, x, y, out, M, N, rows_y, s);
    JCK_LAUNCH_CHECK("rowop");
    return JCK_OK;
}

extern "C" int jck_sigmoid_bce(const float* logit, float* prob, float target, float* scalars, int B, void* stream) {
    JCK_REQUIRE(logit && prob && B > 0, "sigmoid_bce: bad argument");
    launch_pdl(sigmoid_bce_kernel, dim3(grid1d(B)), dim3(256), 0, as_stream(stream), logit, prob, target, scalars, B);
    JCK_LAUNCH_CHECK("sigmoid_bce");
    return JCK_OK;
}

extern "C" int jck_logit_grad(const float* prob, const float* up, float target, float* out, int B, int mode, float scale,
                              void* stream) {
    JCK_REQUIRE(prob && out && B > 0 && mode >= 0 && mode <= 3 && (mode < 2 || up), "logit_grad: bad argument");
    launch_pdl(logit_grad_kernel, dim3(grid1d(B)), dim3(256), 0, as_stream(stream), prob, up, target, out, B, mode, scale);
    JCK_LAUNCH_CHECK("logit_grad");
    return JCK_OK;
}

extern "C" int jck_f32_to_bf16(const float* in, void* out, long long n, void* stream) {
    JCK_REQUIRE(in && out && n > 0, "f32_to_bf16: bad argument");
    launch_pdl(f32_to_bf16_kernel, dim3(grid1d(n)), dim3(256), 0, as_stream(stream), in, (__nv_bfloat16*)out, n);
    JCK_LAUNCH_CHECK("f32_to_bf16");
    return JCK_OK;
}
extern "C" int jck_center_split_bf16(const float* x, const float* mean, void* hi, void* lo, long long rows, int d, int ldo,
                                     void* stream) {
    JCK_REQUIRE(x && mean && hi && lo && rows > 0 && d > 0 && ldo >= d, "center_split_bf16: bad argument");
    launch_pdl(center_split_kernel, dim3(grid1d(rows * ldo)), dim3(256), 0, as_stream(stream), x, mean, (__nv_bfloat16*)hi,
               (__nv_bfloat16*)lo, rows, d, ldo);
    JCK_LAUNCH_CHECK("center_split_bf16");
    return JCK_OK;
}
extern "C" int jck_i64_to_f32(const long long* in, float* out, long long n, void* stream) {
    JCK_REQUIRE(in && out && n > 0, "i64_to_f32: bad argument");
    launch_pdl(i64_to_f32_kernel, dim3(grid1d(n)), dim3(256), 0, as_stream(stream), in, out, n);
    JCK_LAUNCH_CHECK("i64_to_f32");
    return JCK_OK;
}

extern "C" int jck_axpy(const void* x, void* y, float a, long long n, int dtype, void* stream) {
    JCK_REQUIRE(x && y && n > 0, "axpy: bad argument");
    DISPATCH_DTYPE(dtype, "axpy", launch_pdl(axpy_kernel<T>, dim3(grid1d(n)), dim3(256), 0, as_stream(stream), (const T*)x, (T*)y, a, n);)
    JCK_LAUNCH_CHECK("axpy");
    return JCK_OK;
}

extern "C" int jck_gp_seed(const void* v, void* u, float* scalars, int B, long long per_sample, float scale, int dtype,
                           void* stream) {
    JCK_REQUIRE(v && B > 0 && per_sample > 0, "gp_seed: bad argument");
    DISPATCH_DTYPE(dtype, "gp_seed",
        launch_pdl(gp_seed_kernel<T>, dim3(B), dim3(256), 0, as_stream(stream), (const T*)v, (T*)u, scalars, B, per_sample, scale);)
    JCK_LAUNCH_CHECK("gp_seed");
    return JCK_OK;
}

extern "C" int jck_pack_linear(const float* w, void* w_a, float* w_b, int O, int C, int HW, int E, int dtype, void* stream) {
    JCK_REQUIRE(w && w_a && w_b && O > 0 && C > 0 && HW > 0 && E >= 0, "pack_linear: bad argument");
    DISPATCH_DTYPE(dtype, "pack_linear",
        launch_pdl(pack_linear_kernel<T>, dim3(grid1d((long long)O * (C * HW + E))), dim3(256), 0, as_stream(stream), w, (T*)w_a, w_b, O, C, HW, E);)
    JCK_LAUNCH_CHECK("pack_linear");
    return JCK_OK;
}
extern "C" int jck_unpack_linear_grad(const float* dwa, const float* dwb, float* dw, int O, int C, int HW, int E, int accumulate,
                                      void* stream) {
    JCK_REQUIRE(dwa && dwb && dw && O > 0, "unpack_linear_grad: bad argument");
    launch_pdl(unpack_linear_grad_kernel, dim3(grid1d((long long)O * (C * HW + E))), dim3(256), 0, as_stream(stream), dwa, dwb, dw, O, C, HW, E, accumulate);
    JCK_LAUNCH_CHECK("unpack_linear_grad");
    return JCK_OK;
}

extern "C" int jck_bn_adj_reduce(const void* dbar, const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                                 const float* sums1, float* asums, long long npix, int C, float count, float slope, int dtype,
                                 void* stream) {
    JCK_REQUIRE(dbar && da && y && scale_shift && mean_rstd && sums1 && asums && npix > 0 && C > 0 && count > 0,
                "bn_adj_reduce: bad argument");
    const int tx = C < 128 ? C : 128;
    long long slabs = (2LL * kNumSMs * 128) / C;
    if (slabs > npix / 4) slabs = npix / 4;
    if (slabs < 1) slabs = 1;
    dim3 grid((unsigned)slabs, (unsigned)((C + tx - 1) / tx));
    DISPATCH_DTYPE(dtype, "bn_adj_reduce",
        launch_pdl(bn_adj_reduce_kernel<T>, dim3(grid), dim3(tx), 0, as_stream(stream), (const T*)dbar, (const T*)da, (const T*)y, scale_shift,
                                                                   mean_rstd, sums1, asums, C, npix, 1.f / count, slope);)
    JCK_LAUNCH_CHECK("bn_adj_reduce");
    return JCK_OK;
}

extern "C" int jck_bn_adj_apply(const void* dbar, const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                                const float* gamma, const float* sums1, const float* asums, void* gbar_a, void* ybar,
                                long long npix, int C, float count, float slope, int dtype, void* stream) {
    JCK_REQUIRE(dbar && da && y && scale_shift && mean_rstd && gamma && sums1 && asums && gbar_a && ybar && npix > 0 && count > 0,
                "bn_adj_apply: bad argument");
    DISPATCH_DTYPE(dtype, "bn_adj_apply",
        launch_pdl(bn_adj_apply_kernel<T>, dim3(grid1d(npix * C)), dim3(256), 0, as_stream(stream), (const T*)dbar, (const T*)da, (const T*)y,
                                                                              scale_shift, mean_rstd, gamma, sums1, asums,
                                                                              (T*)gbar_a, (T*)ybar, C, npix, 1.f / count, slope);)
    JCK_LAUNCH_CHECK("bn_adj_apply");
    return JCK_OK;
}

extern "C" int jck_bn_adj_param(const float* asums, const float* mean_rstd, float* dgamma, int C, float scale, void* stream) {
    JCK_REQUIRE(asums && mean_rstd && dgamma && C > 0, "bn_adj_param: bad argument");
    launch_pdl(bn_adj_param_kernel, dim3((C + 127) / 128), dim3(128), 0, as_stream(stream), asums, mean_rstd, dgamma, C, scale);
    JCK_LAUNCH_CHECK("bn_adj_param");
    return JCK_OK;
}
