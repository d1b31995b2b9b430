// CUDA-core (fp32 FMA) implicit-GEMM kernels.
//
// Role: (1) the exact-fp32 parity mode of every 4x4 stride-2 layer (north_star: rel err <= 1e-4 in
// fp32 -- out of reach of single-pass TF32/BF16 tensor-core math), (2) the layers whose GEMM view is
// too skinny for a tcgen05 tile (nc = 3 image edge: D.conv1 / G.conv5, K or N = 3 channels) and the
// G.conv1 dense layer.  GEMM-shaped layers in bf16 mode run in conv_tc.cu instead.
//
// One kernel template: C[m][n] = sum_k A(m,k) * B(n,k), 64x64 tile, BK = 16, 256 threads, 4x4
// register micro-tile, operands gathered through a Problem functor (im2col addressing, zero padding),
// fp32 accumulation regardless of storage dtype.
#include "common.cuh"

namespace jck {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256, PAD = 4;

// ---------------------------------------------------------------------------------------------
// Problems
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DownProblem {  // large -> small.  M = B*Hs*Ws, N = Ca, K = 16*Cb
    static constexpr bool kKContigLoad = true;
    const T* in; const T* w; T* out; float* stats;
    int B, Hs, Ws, Ca, Cb, ipg;
    __device__ int M() const { return B * Hs * Ws; }
    __device__ int N() const { return Ca; }
    __device__ int K() const { return 16 * Cb; }
    struct Row { int n, oy, ox; bool ok; };
    __device__ Row rowA(int m, int) const {
        Row r; r.ok = m < M();
        r.ox = m % Ws; int t = m / Ws; r.oy = t % Hs; r.n = t / Hs; return r;
    }
    __device__ float loadA(const Row& r, int k, int) const {
        if (!r.ok || k >= K()) return 0.f;
        int tap = k / Cb, c = k - tap * Cb;
        int iy = 2 * r.oy + (tap >> 2) - 1, ix = 2 * r.ox + (tap & 3) - 1;
        if ((unsigned)iy >= (unsigned)(2 * Hs) || (unsigned)ix >= (unsigned)(2 * Ws)) return 0.f;
        return ld_act(in + ((size_t)(r.n * 2 * Hs + iy) * (2 * Ws) + ix) * Cb + c);
    }
    __device__ float loadB(int n, int k, int) const {
        if (n >= Ca || k >= K()) return 0.f;
        return ld_act(w + (size_t)n * 16 * Cb + k);
    }
    __device__ void store(int m, int n, float v, int) const {
        if (m < M() && n < Ca) st_act(out + (size_t)m * Ca + n, v);
    }
    __device__ float* stats_ptr(int m, int) const {  // group of row m
        return stats ? stats + (size_t)((m / (Hs * Ws)) / ipg) * 2 * Ca : nullptr;
    }
    __device__ int stats_channels() const { return Ca; }
    __device__ bool uniform_group() const { return ipg >= B || ((ipg * Hs * Ws) % BM) == 0; }
};

template <typename T>
struct UpProblem {  // small -> large, phase = blockIdx.z.  M = B*Hs*Ws, N = Cb, K = 4*Ca
    static constexpr bool kKContigLoad = true;
    const T* in; const T* w; T* out; float* stats;
    int B, Hs, Ws, Ca, Cb, ipg;
    __device__ int M() const { return B * Hs * Ws; }
    __device__ int N() const { return Cb; }
    __device__ int K() const { return 4 * Ca; }
    struct Row { int n, i, j; bool ok; };
    __device__ Row rowA(int m, int) const {
        Row r; r.ok = m < M();
        r.j = m % Ws; int t = m / Ws; r.i = t % Hs; r.n = t / Hs; return r;
    }
    __device__ float loadA(const Row& r, int k, int z) const {
        if (!r.ok || k >= K()) return 0.f;
        int t = k / Ca, a = k - t * Ca;
        int ii = r.i + up_d(z >> 1, t >> 1), jj = r.j + up_d(z & 1, t & 1);
        if ((unsigned)ii >= (unsigned)Hs || (unsigned)jj >= (unsigned)Ws) return 0.f;
        return ld_act(in + ((size_t)(r.n * Hs + ii) * Ws + jj) * Ca + a);
    }
    __device__ float loadB(int n, int k, int z) const {
        if (n >= Cb || k >= K()) return 0.f;
        return ld_act(w + ((size_t)z * Cb + n) * 4 * Ca + k);
    }
    __device__ void store(int m, int n, float v, int z) const {
        if (m >= M() || n >= Cb) return;
        int j = m % Ws; int t = m / Ws; int i = t % Hs; int nn = t / Hs;
        size_t pix = (size_t)(nn * 2 * Hs + 2 * i + (z >> 1)) * (2 * Ws) + 2 * j + (z & 1);
        st_act(out + pix * Cb + n, v);
    }
    __device__ float* stats_ptr(int m, int) const {
        return stats ? stats + (size_t)((m / (Hs * Ws)) / ipg) * 2 * Cb : nullptr;
    }
    __device__ int stats_channels() const { return Cb; }
    __device__ bool uniform_group() const { return ipg >= B || ((ipg * Hs * Ws) % BM) == 0; }
};

template <typename T>
struct WgradProblem {  // M = Ca, N = 16*Cb, K = B*Hs*Ws pixels split over blockIdx.z
    static constexpr bool kKContigLoad = false;
    const T* small; const T* large; float* part;
    int B, Hs, Ws, Ca, Cb, ksplit_len;
    __device__ int M() const { return Ca; }
    __device__ int N() const { return 16 * Cb; }
    __device__ int K() const { return B * Hs * Ws; }
    struct Row { int a; bool ok; };
    __device__ Row rowA(int m, int) const { return Row{m, m < Ca}; }
    __device__ float loadA(const Row& r, int k, int) const {
        if (!r.ok || k >= K()) return 0.f;
        return ld_act(small + (size_t)k * Ca + r.a);
    }
    __device__ float loadB(int n, int k, int) const {
        if (n >= N() || k >= K()) return 0.f;
        int tap = n / Cb, b = n - tap * Cb;
        int ox = k % Ws; int t = k / Ws; int oy = t % Hs; int img = t / Hs;
        int iy = 2 * oy + (tap >> 2) - 1, ix = 2 * ox + (tap & 3) - 1;
        if ((unsigned)iy >= (unsigned)(2 * Hs) || (unsigned)ix >= (unsigned)(2 * Ws)) return 0.f;
        return ld_act(large + ((size_t)(img * 2 * Hs + iy) * (2 * Ws) + ix) * Cb + b);
    }
    __device__ void store(int m, int n, float v, int z) const {
        if (m < Ca && n < N()) part[((size_t)z * Ca + m) * 16 * Cb + n] = v;
    }
    __device__ float* stats_ptr(int, int) const { return nullptr; }
    __device__ int stats_channels() const { return 0; }
    __device__ bool uniform_group() const { return true; }
};

template <typename T>
struct FcFwdProblem {  // out[m][n] = sum_k x[m][k] w[n][k]
    static constexpr bool kKContigLoad = true;
    const float* x; const T* w; T* out; float* stats;
    int Mm, Nn, Kk, C;
    __device__ int M() const { return Mm; }
    __device__ int N() const { return Nn; }
    __device__ int K() const { return Kk; }
    struct Row { int m; bool ok; };
    __device__ Row rowA(int m, int) const { return Row{m, m < Mm}; }
    __device__ float loadA(const Row& r, int k, int) const { return (r.ok && k < Kk) ? x[(size_t)r.m * Kk + k] : 0.f; }
    __device__ float loadB(int n, int k, int) const { return (n < Nn && k < Kk) ? ld_act(w + (size_t)n * Kk + k) : 0.f; }
    __device__ void store(int m, int n, float v, int) const { if (m < Mm && n < Nn) st_act(out + (size_t)m * Nn + n, v); }
    // channel = n % C: handled by the generic epilogue through a channel remap
    __device__ float* stats_ptr(int, int) const { return stats; }
    __device__ int stats_channels() const { return C; }
    __device__ bool uniform_group() const { return true; }
};

template <typename T>
struct FcWgradProblem {  // dw[n][k] = sum_m dy[m][n] x[m][k]:  M' = N, N' = K, K' = M
    static constexpr bool kKContigLoad = false;
    const T* dy; const float* x; float* dw;
    int Mm, Nn, Kk, accumulate;
    __device__ int M() const { return Nn; }
    __device__ int N() const { return Kk; }
    __device__ int K() const { return Mm; }
    struct Row { int n; bool ok; };
    __device__ Row rowA(int m, int) const { return Row{m, m < Nn}; }
    __device__ float loadA(const Row& r, int k, int) const { return (r.ok && k < Mm) ? ld_act(dy + (size_t)k * Nn + r.n) : 0.f; }
    __device__ float loadB(int n, int k, int) const { return (n < Kk && k < Mm) ? x[(size_t)k * Kk + n] : 0.f; }
    __device__ void store(int m, int n, float v, int) const {
        if (m < Nn && n < Kk) { float* p = dw + (size_t)m * Kk + n; *p = accumulate ? *p + v : v; }
    }
    __device__ float* stats_ptr(int, int) const { return nullptr; }
    __device__ int stats_channels() const { return 0; }
    __device__ bool uniform_group() const { return true; }
};

// Generic strided dense product C[m][n] (+)= sum_k A(m,k) * B(n,k), A(m,k) = A[m*sam + k*sak],
// B(n,k) = B[n*sbn + k*sbk]; operand / result dtypes chosen at run time (JCK_F32 / JCK_BF16).  The CGAN
// discriminator head (label embedding, Linear 8392 -> 256 -> 1 and every first / second order product
// around them) is built from this one kernel.
__device__ __forceinline__ float ld_any(const void* p, size_t i, int dt) {
    return dt == JCK_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                          : reinterpret_cast<const float*>(p)[i];
}
template <bool KCONTIG>
struct DenseProblem {
    static constexpr bool kKContigLoad = KCONTIG;
    const void* A; const void* B; void* C;
    int a_dt, b_dt, c_dt;
    long long sam, sak, sbn, sbk, ldc;
    int Mm, Nn, Kk, accumulate;
    int atomic;                // split-K launch: every split adds its partial with atomicAdd (C zeroed or accumulated into)
    __device__ int M() const { return Mm; }
    __device__ int N() const { return Nn; }
    __device__ int K() const { return Kk; }
    struct Row { int m; bool ok; };
    __device__ Row rowA(int m, int) const { return Row{m, m < Mm}; }
    __device__ float loadA(const Row& r, int k, int) const {
        return (r.ok && k < Kk) ? ld_any(A, (size_t)(r.m * sam + k * sak), a_dt) : 0.f;
    }
    __device__ float loadB(int n, int k, int) const {
        return (n < Nn && k < Kk) ? ld_any(B, (size_t)(n * sbn + k * sbk), b_dt) : 0.f;
    }
    __device__ void store(int m, int n, float v, int) const {
        if (m >= Mm || n >= Nn) return;
        const size_t i = (size_t)m * ldc + n;
        if (c_dt == JCK_BF16) {
            __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(C) + i;
            *p = __float2bfloat16_rn(accumulate ? __bfloat162float(*p) + v : v);
        } else {
            float* p = reinterpret_cast<float*>(C) + i;
            if (atomic) atomicAdd(p, v);
            else *p = accumulate ? *p + v : v;
        }
    }
    __device__ float* stats_ptr(int, int) const { return nullptr; }
    __device__ int stats_channels() const { return 0; }
    __device__ bool uniform_group() const { return true; }
};

// ---------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------
template <typename P>
__global__ void __launch_bounds__(NT) simt_gemm_kernel(const P p, int k_per_split) {
    pdl_entry();
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    __shared__ float red[2][16][BN];

    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int Kt = p.K();
    int kbeg = 0, kend = Kt;
    if (k_per_split > 0) { kbeg = z * k_per_split; kend = min(Kt, kbeg + k_per_split); }

    // load mapping: each thread fetches 4 consecutive k of one row
    int lrow, lkq;
    if (P::kKContigLoad) { lrow = tid >> 2; lkq = tid & 3; } else { lrow = tid & 63; lkq = tid >> 6; }
    const typename P::Row rowA = p.rowA(m0 + lrow, z);
    const int colB = n0 + lrow;

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        float av[4], bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + lkq * 4 + j;
            bool kin = k < kend;
            av[j] = kin ? p.loadA(rowA, k, z) : 0.f;
            bv[j] = kin ? p.loadB(colB, k, z) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[lkq * 4 + j][lrow] = av[j];
            Bs[lkq * 4 + j][lrow] = bv[j];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) p.store(m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j], z);

    // per-output-channel sum / sum of squares for BatchNorm (rows past M contribute exact zeros)
    const int SC = p.stats_channels();
    if (SC > 0 && p.stats_ptr(m0, z) != nullptr) {
        if (p.uniform_group()) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = 0.f, q = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) { s += acc[i][j]; q += acc[i][j] * acc[i][j]; }
                red[0][ty][tx * 4 + j] = s;
                red[1][ty][tx * 4 + j] = q;
            }
            __syncthreads();
            if (tid < 2 * BN) {
                const int which = tid / BN, col = tid % BN;
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < 16; ++r) s += red[which][r][col];
                const int n = n0 + col;
                if (n < p.N()) atomicAdd(p.stats_ptr(m0, z) + which * SC + (n % SC), s);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + ty * 4 + i;
                if (m >= p.M()) continue;
                float* sp = p.stats_ptr(m, z);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = n0 + tx * 4 + j;
                    if (n < p.N()) {
                        atomicAdd(sp + (n % SC), acc[i][j]);
                        atomicAdd(sp + SC + (n % SC), acc[i][j] * acc[i][j]);
                    }
                }
            }
        }
    }
}

template <typename P>
int launch(const P& p, int M, int N, int zdim, int k_per_split, cudaStream_t st, const char* name) {
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, zdim);
    launch_pdl(simt_gemm_kernel<P>, dim3(grid), dim3(NT), 0, st, p, k_per_split);
    JCK_LAUNCH_CHECK(name);
    return JCK_OK;
}

// reduce split-K partials [splits][Ca][16][Cb] and transpose into w4[Ca][Cb][16] (+= or =).
// One block per (a, 64-wide b tile): the 16 x 64 slab is read as 16 coalesced 256-byte rows per split, summed,
// transposed through shared memory and written as ONE contiguous 4 KB run of w4 (the old element-wise version
// scattered 4-byte read-modify-writes 64 bytes apart).
constexpr int kUnpackBT = 64;
// any Cb (the 3-channel image edge in fp32 parity mode): element-wise
__global__ void wgrad_unpack_small_kernel(const float* __restrict__ part, float* __restrict__ dw4, int Ca, int Cb,
                                          int splits, int accumulate) {
    pdl_entry();
    const size_t total = (size_t)Ca * 16 * Cb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int b = idx % Cb; const size_t t = idx / Cb; const int tap = t % 16; const int a = t / 16;
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[(size_t)z * total + idx];
        float* dst = dw4 + ((size_t)a * Cb + b) * 16 + tap;
        *dst = accumulate ? *dst + s : s;
    }
}

__global__ void __launch_bounds__(256)
wgrad_unpack_kernel(const float* __restrict__ part, float* __restrict__ dw4, int Ca, int Cb, int splits, int accumulate) {
    pdl_entry();
    __shared__ float tile[16][kUnpackBT + 1];
    const int btiles = Cb / kUnpackBT;
    const size_t total = (size_t)Ca * 16 * Cb;
    for (int blk = blockIdx.x; blk < Ca * btiles; blk += gridDim.x) {
        const int a = blk / btiles, b0 = (blk % btiles) * kUnpackBT;
        // thread t sums element (tap = t / 64 + 4 j, b = t % 64), j = 0..3
        const int bl = threadIdx.x % kUnpackBT, t0 = threadIdx.x / kUnpackBT;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        const float* src = part + ((size_t)a * 16 + t0) * Cb + b0 + bl;
        int z = 0;
        for (; z + 4 <= splits; z += 4) {                         // 16 independent loads in flight per thread
            float v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[u][j] = __ldg(src + (size_t)(z + u) * total + (size_t)(4 * j) * Cb);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[j] += v[u][j];
        }
        for (; z < splits; ++z) {
#pragma unroll
            for (int j = 0; j < 4; ++j) s[j] += __ldg(src + (size_t)z * total + (size_t)(4 * j) * Cb);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) tile[t0 + 4 * j][bl] = s[j];
        __syncthreads();
        float* dst = dw4 + ((size_t)a * Cb + b0) * 16;            // 64 * 16 contiguous floats
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = threadIdx.x + 256 * j;                     // o = b_local * 16 + tap
            const float v = tile[o & 15][o >> 4];
            dst[o] = accumulate ? dst[o] + v : v;
        }
        __syncthreads();
    }
}

}  // namespace

// exported to conv_tc.cu / api
int simt_wgrad_splits(int B, int Hs, int Ws, int Ca, int Cb) {
    const long long K = (long long)B * Hs * Ws;
    const int tiles = ((Ca + BM - 1) / BM) * ((16 * Cb + BN - 1) / BN);
    int splits = (4 * kNumSMs + tiles - 1) / tiles;
    const int maxs = (int)((K + 4 * BK - 1) / (4 * BK));
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    return splits;
}

int launch_wgrad_unpack(const float* part, float* dw4, int Ca, int Cb, int splits, int accumulate, cudaStream_t st) {
    if (Cb % kUnpackBT != 0) {
        const size_t total = (size_t)Ca * 16 * Cb;
        int nb = (int)((total + 255) / 256);
        if (nb > 4 * kNumSMs) nb = 4 * kNumSMs;
        launch_pdl(wgrad_unpack_small_kernel, dim3(nb), dim3(256), 0, st, part, dw4, Ca, Cb, splits, accumulate);
        JCK_LAUNCH_CHECK("wgrad_unpack");
        return JCK_OK;
    }
    int blocks = Ca * (Cb / kUnpackBT);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    launch_pdl(wgrad_unpack_kernel, dim3(blocks), dim3(256), 0, st, part, dw4, Ca, Cb, splits, accumulate);
    JCK_LAUNCH_CHECK("wgrad_unpack");
    return JCK_OK;
}

template <typename T>
int simt_down(const void* in, const void* w, void* out, float* stats, int B, int Hs, int Ws, int Ca, int Cb,
              int ipg, cudaStream_t st) {
    DownProblem<T> p{(const T*)in, (const T*)w, (T*)out, stats, B, Hs, Ws, Ca, Cb, ipg};
    return launch(p, B * Hs * Ws, Ca, 1, 0, st, "simt_down");
}
template <typename T>
int simt_up(const void* in, const void* w, void* out, float* stats, int B, int Hs, int Ws, int Ca, int Cb,
            int ipg, cudaStream_t st) {
    UpProblem<T> p{(const T*)in, (const T*)w, (T*)out, stats, B, Hs, Ws, Ca, Cb, ipg};
    return launch(p, B * Hs * Ws, Cb, 4, 0, st, "simt_up");
}
template <typename T>
int simt_wgrad(const void* small, const void* large, float* part, int splits, int B, int Hs, int Ws, int Ca,
               int Cb, cudaStream_t st) {
    const int K = B * Hs * Ws;
    int per = (K + splits - 1) / splits;
    per = (per + BK - 1) / BK * BK;
    WgradProblem<T> p{(const T*)small, (const T*)large, part, B, Hs, Ws, Ca, Cb, per};
    return launch(p, Ca, 16 * Cb, splits, per, st, "simt_wgrad");
}

template int simt_down<float>(const void*, const void*, void*, float*, int, int, int, int, int, int, cudaStream_t);
template int simt_down<__nv_bfloat16>(const void*, const void*, void*, float*, int, int, int, int, int, int, cudaStream_t);
template int simt_up<float>(const void*, const void*, void*, float*, int, int, int, int, int, int, cudaStream_t);
template int simt_up<__nv_bfloat16>(const void*, const void*, void*, float*, int, int, int, int, int, int, cudaStream_t);
template int simt_wgrad<float>(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t);
template int simt_wgrad<__nv_bfloat16>(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t);

}  // namespace jck

using namespace jck;

extern "C" int jck_fc_fwd(const float* x, const void* w, void* out, float* stats, int M, int N, int K, int C,
                          int dtype, void* stream) {
    JCK_REQUIRE(x && w && out && M > 0 && N > 0 && K > 0 && C > 0 && N % C == 0, "fc_fwd: bad argument");
    if (dtype == JCK_F32) {
        FcFwdProblem<float> p{x, (const float*)w, (float*)out, stats, M, N, K, C};
        return launch(p, M, N, 1, 0, as_stream(stream), "fc_fwd");
    } else if (dtype == JCK_BF16) {
        FcFwdProblem<__nv_bfloat16> p{x, (const __nv_bfloat16*)w, (__nv_bfloat16*)out, stats, M, N, K, C};
        return launch(p, M, N, 1, 0, as_stream(stream), "fc_fwd");
    }
    return set_error(JCK_E_BADARG, "fc_fwd: dtype %d", dtype);
}

extern "C" int jck_fc_wgrad(const void* dy, const float* x, float* dw, int M, int N, int K, int accumulate,
                            int dtype, void* stream) {
    JCK_REQUIRE(dy && x && dw && M > 0 && N > 0 && K > 0, "fc_wgrad: bad argument");
    if (dtype == JCK_F32) {
        FcWgradProblem<float> p{(const float*)dy, x, dw, M, N, K, accumulate};
        return launch(p, N, K, 1, 0, as_stream(stream), "fc_wgrad");
    } else if (dtype == JCK_BF16) {
        FcWgradProblem<__nv_bfloat16> p{(const __nv_bfloat16*)dy, x, dw, M, N, K, accumulate};
        return launch(p, N, K, 1, 0, as_stream(stream), "fc_wgrad");
    }
    return set_error(JCK_E_BADARG, "fc_wgrad: dtype %d", dtype);
}

extern "C" int jck_dense(const void* A, int a_dt, long long sam, long long sak, const void* B, int b_dt, long long sbn,
                         long long sbk, void* C, int c_dt, long long ldc, int M, int N, int K, int accumulate, void* stream) {
    JCK_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "dense: bad argument");
    JCK_REQUIRE((a_dt | 1) == 1 && (b_dt | 1) == 1 && (c_dt | 1) == 1, "dense: bad dtype");
    // The CGAN head's small products (bias sums, 256 x 200 weight gradients ...) have a handful of output tiles and a
    // contraction over the batch: one CTA per tile walks it serially (~0.5 us per 16-wide step).  Split the contraction
    // over the grid instead; the splits add their partials atomically into C (zeroed first unless accumulating).
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int splits = 1, k_per_split = 0;
    if (c_dt == JCK_F32 && tiles * 2 <= kNumSMs && K >= 8 * BK) {
        splits = kNumSMs / tiles;
        if (splits > K / (2 * BK)) splits = K / (2 * BK);
        k_per_split = ((K + splits - 1) / splits + BK - 1) / BK * BK;
        splits = (K + k_per_split - 1) / k_per_split;
    }
    const int atomic = splits > 1;
    if (atomic && !accumulate) {
        cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, as_stream(stream));
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "dense memset: %s", cudaGetErrorString(e));
    }
    if (sak == 1) {
        DenseProblem<true> p{A, B, C, a_dt, b_dt, c_dt, sam, sak, sbn, sbk, ldc, M, N, K, accumulate, atomic};
        return launch(p, M, N, splits, k_per_split, as_stream(stream), "dense");
    }
    DenseProblem<false> p{A, B, C, a_dt, b_dt, c_dt, sam, sak, sbn, sbk, ldc, M, N, K, accumulate, atomic};
    return launch(p, M, N, splits, k_per_split, as_stream(stream), "dense");
}
