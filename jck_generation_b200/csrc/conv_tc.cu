// tcgen05 / TMEM / TMA implicit-GEMM kernels for the 4x4 stride-2 layers (bf16 in, fp32 accumulate).
//
// All three GEMM-shaped operations of a layer (Ca = low-res channels, Cb = high-res channels):
//   down  : out_small[pix][a]   = sum_{tap,b} in_large[shift_tap(pix)][b] * Wdown[a][tap][b]
//   up    : out_large[2pix+ph][b] = sum_{t,a} in_small[shift_t(pix)][a]   * Wup[ph][b][t][a]
//   wgrad : dW[a][tap][b]       = sum_pix   small[pix][a] * large[shift_tap(pix)][b]
// The im2col never exists: because activations are NHWC, one filter tap of a 128-pixel output patch is
// a dense [128 pixels][64 channels] box of the input tensor at a shifted coordinate, which TMA tiled
// mode fetches in one instruction and zero-fills outside the image (= the conv padding).  The stride-2
// side is addressed through a 5-D "parity" view  (2C | W/2 | 2 | H/2 | B)  so that a tap is again a
// dense box.  Tiles land 128B-swizzled in shared memory exactly as UMMA wants them (K-major for
// down/up, MN-major for wgrad where the contraction runs over pixels).
//
// Kernel anatomy (one output tile per CTA): warp 0 = TMA producer, warp 1 = TMEM allocator + single
// thread MMA issuer, warps 2..5 = epilogue (tcgen05.ld -> BatchNorm partial sums by warp shuffles ->
// bf16 -> global).  smem ring: full/empty mbarriers; accumulator handed over by tcgen05.commit.
#include <stdlib.h>
#include "tc_common.cuh"

namespace jck {

// SIMT fallbacks / helpers from conv_simt.cu
template <typename T> int simt_down(const void*, const void*, void*, float*, int, int, int, int, int, int, cudaStream_t);
template <typename T> int simt_up(const void*, const void*, void*, float*, int, int, int, int, int, int, cudaStream_t);
template <typename T> int simt_wgrad(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t);
int simt_wgrad_splits(int B, int Hs, int Ws, int Ca, int Cb);
int launch_wgrad_unpack(const float* part, float* dw4, int Ca, int Cb, int splits, int accumulate, cudaStream_t st);

namespace {
using namespace tc;

// ------------------------------------------------------------------------------------------------
// Tensor maps (driver entry point resolved through the runtime: the library never links libcuda,
// so it still loads on a machine without a driver).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return set_error(JCK_E_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims,
                    strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(JCK_E_DRIVER, "cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
    return JCK_OK;
}

// plain NHWC view (C | W | H | B), box (64 | bw | bh | nb)
int map_small(CUtensorMap* m, const void* p, int C, int W, int H, int B, int bw, int bh, int nb) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)nb};
    return encode(m, p, 4, dims, str, box);
}
// parity view of an NHWC tensor with H2 x W2 pixels: (2C | W2/2 | 2 | H2/2 | B), box (64 | bw | 1 | bh | nb).
// element (px*C + c, xh, py, yh, n)  ==  pixel (n, 2*yh + py, 2*xh + px), channel c.
int map_large(CUtensorMap* m, const void* p, int C, int W2, int H2, int B, int bw, int bh, int nb) {
    cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)W2 / 2, 2, (cuuint64_t)H2 / 2, (cuuint64_t)B};
    cuuint64_t str[4] = {(cuuint64_t)2 * C * 2, (cuuint64_t)W2 * C * 2, (cuuint64_t)2 * W2 * C * 2,
                         (cuuint64_t)H2 * W2 * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, 1, (cuuint32_t)bh, (cuuint32_t)nb};
    return encode(m, p, 5, dims, str, box);
}
// row-major matrix [rows][cols] bf16, box (64 | box_rows)
int map_matrix(CUtensorMap* m, const void* p, int rows, int cols, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    return encode(m, p, 2, dims, str, box);
}

// row-major fp32 matrix [rows][cols], box (32 floats = one 128-byte swizzle row | box_rows): TMA stores of partials
int map_matrix_f32(CUtensorMap* m, const void* p, long long rows, int cols, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    return encode(m, p, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

// Patch matrix [B*Hs*Ws][64] bf16 made by jck_p4_to_patches: row = output pixel, 64 = (ky, kx, c4).
// (A 5-D tensor map with overlapping strides over the P4 image itself would avoid materialising it, but
// cuTensorMapEncodeTiled-built maps of that shape faulted on sm_100a / driver 580 -- see DESIGN.md.)
int map_rows64(CUtensorMap* m, const void* p, long long rows, int box_rows) {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    return encode(m, p, 2, dims, str, box);
}

struct PatchGeom { int bw, bh, nb; };
// rows of a tile = bw x bh x nb output pixels (x fastest), bw*bh*nb == npix
bool patch_geom(int Hs, int Ws, int npix, PatchGeom* g) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (!pow2(Hs) || !pow2(Ws) || Ws > npix) return false;
    g->bw = Ws;
    g->bh = Hs < npix / g->bw ? Hs : npix / g->bw;
    g->nb = npix / (g->bw * g->bh);
    return g->bw <= 256 && g->bh <= 256 && g->nb <= 256 && g->bw * g->bh * g->nb == npix;
}

// ------------------------------------------------------------------------------------------------
// down / up
// ------------------------------------------------------------------------------------------------
struct ConvTcParams {
    int B, Hs, Ws, Ca, Cb;     // layer geometry
    int bw, bh, nb;            // tile patch
    int tiles_x, tiles_y;      // patches per image row / column
    int ipg;                   // images per BatchNorm group
};

constexpr int kConvThreads = 192;
constexpr int kTileM = 128;
constexpr int kBK = 64;  // bf16 elements per K step = one 128-byte swizzle row

// MODE: which implicit GEMM the tile loop walks
constexpr int kDown = 0;      // 16 taps x Cb/64 chunks, A through the parity view of the large tensor
constexpr int kUpM = 1;       // one output parity: 4 taps x Ca/64 chunks, A = shifted small tensor
constexpr int kEdgeDown = 2;  // (patch-matrix image edge; superseded by edge_down_direct_kernel, kept for the generic tile walker)
constexpr int kEdgeUp = 3;    // image edge: 9 input shifts x Ca/64, N = 16 (4 parities x 4 channels) -> P4 image

// ------------------------------------------------------------------------------------------------
// Single-CTA persistent tile walker (used for the image-edge up conv; the trunk layers run on the CTA-pair kernel
// below): a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the smem
// ring keeps streaming across tile boundaries and the accumulator is double-buffered in TMEM, so the
// epilogue of tile i (tcgen05.ld, BatchNorm partial sums, stores) overlaps the MMAs of tile i+1 and the
// per-CTA set-up (TMEM alloc, barrier init, descriptor prefetch) is paid once per SM instead of per tile.
// Tile order: n fastest, so CTAs running side by side share the activation tile in L2.
// ------------------------------------------------------------------------------------------------
template <int BN_, int STAGES>
struct PersistSmem {
    static constexpr int kABytes = kTileM * kBK * 2;
    static constexpr int kBBytes = BN_ * kBK * 2;
    static constexpr int kStage = kABytes + kBBytes;
    static constexpr int kBarOff = STAGES * kStage;
    static constexpr int kRedOff = kBarOff + 512;
    static constexpr int kCoefOff = kRedOff + 2 * 4 * 2 * BN_ * 4;
    static constexpr int kTotal = kCoefOff + 2 * 4 * BN_ * 4 + 1024;
};

// Epilogue mode 1 = BatchNorm-backward reduction fused into the input-gradient convolution that produces
// d/d(activation): with the layer's saved raw output y and its statistics, the epilogue forms
//   g = acc * act'(y*scale + shift),   sums += (sum g, sum g * xhat),   xhat = (y - mean) * rstd
// from the fp32 accumulators and stores g, so the separate reduce pass over (da, y) disappears.
struct BnBwdEpi {
    const __nv_bfloat16* y;   // saved raw conv output, same shape as `out`
    const float* ss;          // [groups][2C] scale | shift
    const float* mr;          // [groups][2C] mean | rstd
    float slope;              // LeakyReLU slope (0 = ReLU)
};

template <int BN_, int STAGES, int MODE, int EPI, int EPIM>   // EPI = epilogue warps: 4 (one per TMEM lane quarter) or 8 (two, splitting the columns)
__global__ void __launch_bounds__(64 + 32 * EPI, EPI == 4 ? 2 : 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                       __nv_bfloat16* __restrict__ out, float* __restrict__ stats, const ConvTcParams p,
                       const int n_tiles, const int total_tiles, const BnBwdEpi bb) {
    pdl_trigger();
    constexpr bool kUp = (MODE == kUpM);
    constexpr int kAccCols = BN_ < 32 ? 32 : BN_;
    using L = PersistSmem<BN_, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;     // [2] accumulator ready
    uint64_t* tempty = tfull + 2;         // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* red = reinterpret_cast<float*>(smem + L::kRedOff);  // [2][4 warps][2][BN_]
    float* coef = reinterpret_cast<float*>(smem + L::kCoefOff); // [2][scale|shift|mean|rstd][BN_]  (EPIM == 1)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cin = (MODE == kUpM || MODE == kEdgeUp) ? p.Ca : p.Cb;
    const int Cout = (MODE == kUpM) ? p.Cb : p.Ca;
    const int cchunks = MODE == kEdgeDown ? 1 : Cin / kBK;
    const int ksteps = MODE == kEdgeDown ? 1 : (MODE == kUpM ? 4 : (MODE == kEdgeUp ? 9 : 16)) * cchunks;
    const int tiles_img = p.tiles_x * p.tiles_y;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 32 * EPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kAccCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // everything above is on-chip set-up; global memory is first touched below

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = t % n_tiles;
                int rest = t / n_tiles;
                const int phase = kUp ? (rest & 3) : 0;
                if (kUp) rest >>= 2;
                const int mt = rest;
                const int x0 = (mt % p.tiles_x) * p.bw, y0 = ((mt / p.tiles_x) % p.tiles_y) * p.bh;
                const int n0 = (mt / tiles_img) * p.nb;
                const int py = phase >> 1, px = phase & 1;
                for (int ks = 0; ks < ksteps; ++ks, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* sa = smem + s * L::kStage;
                    uint8_t* sb = sa + L::kABytes;
                    mbar_arrive_expect_tx(&full[s], L::kStage);
                    const int tap = ks / cchunks, cc = ks - tap * cchunks;
                    if (MODE == kEdgeDown) {
                        tma_load_2d(sa, &mapA, &full[s], 0, mt * kTileM);
                        tma_load_2d(sb, &mapB, &full[s], 0, nt * BN_);
                    } else if (MODE == kEdgeUp) {
                        const int di = tap / 3 - 1, dj = tap % 3 - 1;
                        tma_load_4d(sa, &mapA, &full[s], cc * kBK, x0 + dj, y0 + di, n0);
                        tma_load_2d(sb, &mapB, &full[s], tap * p.Ca + cc * kBK, 0);
                    } else if (!kUp) {
                        const int ky = tap >> 2, kx = tap & 3;
                        const int dy = (ky - 1) >> 1, qy = (ky - 1) & 1;
                        const int dx = (kx - 1) >> 1, qx = (kx - 1) & 1;
                        tma_load_5d(sa, &mapA, &full[s], qx * p.Cb + cc * kBK, x0 + dx, qy, y0 + dy, n0);
                        tma_load_2d(sb, &mapB, &full[s], tap * p.Cb + cc * kBK, nt * BN_);
                    } else {
                        const int dy = up_d(py, tap >> 1), dx = up_d(px, tap & 1);
                        tma_load_4d(sa, &mapA, &full[s], cc * kBK, x0 + dx, y0 + dy, n0);
                        tma_load_2d(sb, &mapB, &full[s], tap * p.Ca + cc * kBK, phase * p.Cb + nt * BN_);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = make_idesc(BN_, 0, 0);
        int it = 0, lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int acc = lt & 1;
            mbar_wait(&tempty[acc], ((lt >> 1) & 1) ^ 1);
            fence_after_sync();
            const uint32_t tmem_d = tmem_base + acc * kAccCols;
            for (int ks = 0; ks < ksteps; ++ks, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                fence_after_sync();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + s * L::kStage);
                    const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_bf16(tmem_d, make_sdesc(a_addr + k * 32, 0, 1024), make_sdesc(b_addr + k * 32, 0, 1024), idesc,
                                  (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                    if (ks == ksteps - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue (warps 2..) ----------------
        const int wq = warp & 3;                     // TMEM lane quarter
        const int half = (warp - 2) >> 2;            // with 8 warps: which half of the 32-column chunks
        constexpr int kChunkStep = EPI / 4;
        const int r = wq * 32 + lane;
        const int xl = r % p.bw, yl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
        int lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int nt = t % n_tiles;
            int rest = t / n_tiles;
            const int phase = kUp ? (rest & 3) : 0;
            if (kUp) rest >>= 2;
            const int mt = rest;
            const int x0 = (mt % p.tiles_x) * p.bw, y0 = ((mt / p.tiles_x) % p.tiles_y) * p.bh;
            const int n0 = (mt / tiles_img) * p.nb;
            const int py = phase >> 1, px = phase & 1;
            const int n = n0 + nl;
            const bool valid = n < p.B;
            const int acc = lt & 1;
            const uint32_t tmem_d = tmem_base + acc * kAccCols + ((uint32_t)(wq * 32) << 16);
            float* cf = coef + (lt & 1) * (4 * BN_);
            if constexpr (EPIM == 1) {
                // this tile's BatchNorm coefficients -> smem while the MMAs of the tile are still running
                const size_t gofs = (size_t)(n0 / p.ipg) * 2 * Cout + nt * BN_;
                for (int i = threadIdx.x - 64; i < 4 * BN_; i += 32 * EPI) {
                    const int which = i / BN_, cc = i - which * BN_;
                    cf[i] = ((which < 2) ? bb.ss : bb.mr)[gofs + (which & 1) * Cout + cc];
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI) : "memory");
            }
            mbar_wait(&tfull[acc], (lt >> 1) & 1);
            fence_after_sync();
            if constexpr (MODE == kEdgeUp) {
                float v[16];
                if (half == 0) { tmem_ld16(tmem_d, v); tmem_ld_wait(); }
                fence_before_sync();
                mbar_arrive(&tempty[acc]);
                if (valid && half == 0) {
                    const int Wp = 2 * p.Ws + 2, Hp = 2 * p.Hs + 2;
#pragma unroll
                    for (int qy = 0; qy < 2; ++qy) {
                        const size_t o = (((size_t)n * Hp + 2 * (y0 + yl) + qy + 1) * Wp + 2 * (x0 + xl) + 1) * 4;
                        uint2 u0, u1;
                        u0.x = pack_bf16x2(v[qy * 8 + 0], v[qy * 8 + 1]);
                        u0.y = pack_bf16x2(v[qy * 8 + 2], v[qy * 8 + 3]);
                        u1.x = pack_bf16x2(v[qy * 8 + 4], v[qy * 8 + 5]);
                        u1.y = pack_bf16x2(v[qy * 8 + 6], v[qy * 8 + 7]);
                        *reinterpret_cast<uint2*>(out + o) = u0;
                        *reinterpret_cast<uint2*>(out + o + 4) = u1;
                    }
                }
            } else {
                size_t pix;
                if (!kUp) pix = ((size_t)n * p.Hs + (y0 + yl)) * p.Ws + (x0 + xl);
                else pix = ((size_t)n * 2 * p.Hs + 2 * (y0 + yl) + py) * (2 * p.Ws) + 2 * (x0 + xl) + px;
                __nv_bfloat16* orow = out + pix * Cout + nt * BN_;
                float* rbuf = red + (lt & 1) * (4 * 2 * BN_);
                constexpr int kChunks = BN_ / 32;
#pragma unroll 1
                for (int c = half; c < kChunks; c += kChunkStep) {
                    float v[32];
                    tmem_ld32(tmem_d + c * 32, v);
                    tmem_ld_wait();
                    if (c + kChunkStep >= kChunks) {    // this warp's share of the accumulator is read: hand it back
                        fence_before_sync();
                        mbar_arrive(&tempty[acc]);
                    }
                    float sq[32];
                    if constexpr (EPIM == 1) {
                        uint4 yr[4];
                        if (valid) {
                            const uint4* ys = reinterpret_cast<const uint4*>(bb.y + pix * Cout + nt * BN_ + c * 32);
#pragma unroll
                            for (int q = 0; q < 4; ++q) yr[q] = __ldg(ys + q);
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q) yr[q] = make_uint4(0, 0, 0, 0);
                        }
                        const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(yr);
                        const float* c0 = cf + c * 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float yv = __bfloat162float(yb[i]);
                            const float pre = fmaf(yv, c0[i], c0[BN_ + i]);
                            const float gg = pre > 0.f ? v[i] : v[i] * bb.slope;
                            v[i] = gg;
                            sq[i] = gg * (yv - c0[2 * BN_ + i]) * c0[3 * BN_ + i];
                        }
                    }
                    if (valid) {
                        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 u;
                            u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                            u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                            u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                            u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                            dst[q] = u;
                        }
                    }
                    if (stats != nullptr) {
                        if constexpr (EPIM == 0) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
                        }
                        const float s1 = warp_transpose_sum(v, lane);
                        const float s2 = warp_transpose_sum(sq, lane);
                        rbuf[(wq * 2 + 0) * BN_ + c * 32 + lane] = s1;
                        rbuf[(wq * 2 + 1) * BN_ + c * 32 + lane] = s2;
                    }
                }
                if (stats != nullptr) {
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI) : "memory");   // epilogue warps only
                    const int e = threadIdx.x - 64;
                    float* sp = stats + (size_t)(n0 / p.ipg) * 2 * Cout + nt * BN_;
                    for (int col = e; col < 2 * BN_; col += 32 * EPI) {
                        const int which = col / BN_, cc = col % BN_;
                        const float s = rbuf[(0 * 2 + which) * BN_ + cc] + rbuf[(1 * 2 + which) * BN_ + cc] +
                                        rbuf[(2 * 2 + which) * BN_ + cc] + rbuf[(3 * 2 + which) * BN_ + cc];
                        atomicAdd(sp + which * Cout + cc, s);
                    }
                }
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 2 * kAccCols);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the measured bound of the kernels above is the L2 -> shared-memory
// stream (~6300 B/clk chip-wide, 128x128 tiles move 32 KB per 2.1 MFLOP).  Here two CTAs of a 2-CTA cluster,
// on the two SMs of one TPC, compute a 256-pixel x BN tile with ONE MMA of M = 256: each CTA stages its own 128
// pixel rows of A but only HALF of the weight tile (the tensor core reads both halves across the pair), so the
// bytes per flop drop by 25 % (BN = 128) / 17 % (BN = 64) at the same shared-memory footprint -- which buys a
// fourth pipeline stage.  The leader (cluster rank 0) issues the MMAs and owns the `full` and `tempty`
// barriers; TMA completions of both CTAs land on the leader's `full`, tcgen05.commit multicasts `empty` /
// `tfull` to both CTAs, both CTAs' epilogue warps release the accumulator on the leader's `tempty`.
// ------------------------------------------------------------------------------------------------
template <int BN_, int STAGES>
struct PairSmem {
    static constexpr int kABytes = kTileM * kBK * 2;
    static constexpr int kBBytes = (BN_ / 2) * kBK * 2;
    static constexpr int kStage = kABytes + kBBytes;
    static constexpr int kBarOff = STAGES * kStage;
    static constexpr int kRedOff = kBarOff + 512;
    static constexpr int kCoefOff = kRedOff + 2 * 4 * 2 * BN_ * 4;
    static constexpr int kTotal = kCoefOff + 2 * 4 * BN_ * 4 + 1024;
};

template <int BN_, int STAGES, int MODE, int EPIM>
__global__ void __launch_bounds__(192, BN_ == 256 ? 1 : 2)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ stats, const ConvTcParams p,
                    const int n_tiles, const int total_pairs, const BnBwdEpi bb) {
    pdl_trigger();
    static_assert(MODE == kDown || MODE == kUpM, "pair kernel: trunk layers only");
    constexpr bool kUp = (MODE == kUpM);
    constexpr int kAccCols = BN_;
    constexpr int EPI = 4;
    using L = PairSmem<BN_, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* red = reinterpret_cast<float*>(smem + L::kRedOff);
    float* coef = reinterpret_cast<float*>(smem + L::kCoefOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int Cin = kUp ? p.Ca : p.Cb;
    const int Cout = kUp ? p.Cb : p.Ca;
    const int cchunks = Cin / kBK;
    const int ksteps = (kUp ? 4 : 16) * cchunks;
    const int tiles_img = p.tiles_x * p.tiles_y;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * 32 * EPI); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta(tmem_slot, 2 * kAccCols);
    fence_before_sync();
    cluster_sync_all();                    // barriers of BOTH CTAs are initialised before anyone signals across
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // everything above is on-chip set-up; global memory is first touched below

    auto decode = [&](int tp, int& nt, int& phase, int& x0, int& y0, int& n0) {
        nt = tp % n_tiles;
        int rest = tp / n_tiles;
        phase = kUp ? (rest & 3) : 0;
        if (kUp) rest >>= 2;
        const int mt = 2 * rest + rank;
        x0 = (mt % p.tiles_x) * p.bw;
        y0 = ((mt / p.tiles_x) % p.tiles_y) * p.bh;
        n0 = (mt / tiles_img) * p.nb;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer (both CTAs) ----------------
            int it = 0;
            for (int tp = cid; tp < total_pairs; tp += ncl) {
                int nt, phase, x0, y0, n0;
                decode(tp, nt, phase, x0, y0, n0);
                const int py = phase >> 1, px = phase & 1;
                for (int ks = 0; ks < ksteps; ++ks, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* sa = smem + s * L::kStage;
                    uint8_t* sb = sa + L::kABytes;
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * L::kStage);   // both CTAs' bytes
                    const int tap = ks / cchunks, cc = ks - tap * cchunks;
                    if (!kUp) {
                        const int ky = tap >> 2, kx = tap & 3;
                        const int dy = (ky - 1) >> 1, qy = (ky - 1) & 1;
                        const int dx = (kx - 1) >> 1, qx = (kx - 1) & 1;
                        tma_load_5d_2cta(sa, &mapA, &full[s], qx * p.Cb + cc * kBK, x0 + dx, qy, y0 + dy, n0);
                        tma_load_2d_2cta(sb, &mapB, &full[s], tap * p.Cb + cc * kBK, nt * BN_ + rank * (BN_ / 2));
                    } else {
                        const int dy = up_d(py, tap >> 1), dx = up_d(px, tap & 1);
                        tma_load_4d_2cta(sa, &mapA, &full[s], cc * kBK, x0 + dx, y0 + dy, n0);
                        tma_load_2d_2cta(sb, &mapB, &full[s], tap * p.Ca + cc * kBK, phase * p.Cb + nt * BN_ + rank * (BN_ / 2));
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            // ---------------- MMA issuer: ONE thread of the leader CTA ----------------
            // (no per-stage warp reconvergence; the operand descriptors of a stage differ from those of stage 0 only in the
            // start-address field, so they are formed by one 64-bit add)
            constexpr uint32_t idesc = make_idesc(BN_, 0, 0, 256);
            const uint64_t a_desc0 = make_sdesc(smem_u32(smem), 0, 1024);
            const uint64_t b_desc0 = make_sdesc(smem_u32(smem) + L::kABytes, 0, 1024);
            int s = 0;
            uint32_t full_par = 0;
            int lt = 0;
            for (int tp = cid; tp < total_pairs; tp += ncl, ++lt) {
                const int acc = lt & 1;
                mbar_wait(&tempty[acc], ((lt >> 1) & 1) ^ 1);
                fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * kAccCols;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full[s], full_par);
                    fence_after_sync();
                    const uint64_t so = (uint64_t)((s * L::kStage) >> 4);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_bf16_2cta(tmem_d, a_desc0 + so + (uint64_t)(k * 2), b_desc0 + so + (uint64_t)(k * 2), idesc,
                                       (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2cta(&empty[s], 3);
                    if (ks == ksteps - 1) umma_commit_2cta(&tfull[acc], 3);
                    if (++s == STAGES) { s = 0; full_par ^= 1; }
                }
            }
        }
    } else {
        // ---------------- epilogue (warps 2..5 of both CTAs: this CTA's 128 rows) ----------------
        const int wq = warp & 3;
        const int r = wq * 32 + lane;
        const int xl = r % p.bw, yl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
        // BatchNorm partial sums stay in registers across this CTA's tiles (thread e owns columns e, e + 128 of the
        // [sum | sum of squares] x BN_ block) and are flushed with one atomic per column when the (group, n-tile)
        // they belong to changes -- a cluster keeps its n-tile when the cluster count divides by n_tiles -- instead
        // of 2*BN_ same-address atomics per tile
        float acc_stat[(2 * BN_ + 127) / 128];
#pragma unroll
        for (int j = 0; j < (2 * BN_ + 127) / 128; ++j) acc_stat[j] = 0.f;
        long long acc_key = -1;
        auto flush_stats = [&]() {
            if (acc_key < 0) return;
            float* sp = stats + (size_t)acc_key;
#pragma unroll
            for (int j = 0; j < (2 * BN_ + 127) / 128; ++j) {
                const int col = (threadIdx.x - 64) + j * 128;
                if (col < 2 * BN_) atomicAdd(sp + (col / BN_) * Cout + (col % BN_), acc_stat[j]);
                acc_stat[j] = 0.f;
            }
        };
        int lt = 0;
        for (int tp = cid; tp < total_pairs; tp += ncl, ++lt) {
            int nt, phase, x0, y0, n0;
            decode(tp, nt, phase, x0, y0, n0);
            const int py = phase >> 1, px = phase & 1;
            const int n = n0 + nl;
            const bool valid = n < p.B;
            const bool tile_live = n0 < p.B;            // the odd pair member past the last tile holds zeros only
            const int acc = lt & 1;
            const uint32_t tmem_d = tmem_base + acc * kAccCols + ((uint32_t)(wq * 32) << 16);
            float* cf = coef + (lt & 1) * (4 * BN_);
            if constexpr (EPIM == 1) {
                if (tile_live) {
                    const size_t gofs = (size_t)(n0 / p.ipg) * 2 * Cout + nt * BN_;
                    for (int i = threadIdx.x - 64; i < 4 * BN_; i += 32 * EPI) {
                        const int which = i / BN_, cc = i - which * BN_;
                        cf[i] = ((which < 2) ? bb.ss : bb.mr)[gofs + (which & 1) * Cout + cc];
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI) : "memory");
            }
            mbar_wait(&tfull[acc], (lt >> 1) & 1);
            fence_after_sync();
            size_t pix;
            if (!kUp) pix = ((size_t)n * p.Hs + (y0 + yl)) * p.Ws + (x0 + xl);
            else pix = ((size_t)n * 2 * p.Hs + 2 * (y0 + yl) + py) * (2 * p.Ws) + 2 * (x0 + xl) + px;
            __nv_bfloat16* orow = out + pix * Cout + nt * BN_;
            float* rbuf = red + (lt & 1) * (4 * 2 * BN_);
            constexpr int kChunks = BN_ / 32;
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c) {
                float v[32];
                tmem_ld32(tmem_d + c * 32, v);
                tmem_ld_wait();
                if (c == kChunks - 1) {                 // accumulator read: hand it back to the leader's MMA warp
                    fence_before_sync();
                    mbar_arrive_leader(&tempty[acc]);
                }
                float sq[32];
                if constexpr (EPIM == 1) {
                    uint4 yr[4];
                    if (valid) {
                        const uint4* ys = reinterpret_cast<const uint4*>(bb.y + pix * Cout + nt * BN_ + c * 32);
#pragma unroll
                        for (int q = 0; q < 4; ++q) yr[q] = __ldg(ys + q);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) yr[q] = make_uint4(0, 0, 0, 0);
                    }
                    const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(yr);
                    const float* c0 = cf + c * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float yv = __bfloat162float(yb[i]);
                        const float pre = fmaf(yv, c0[i], c0[BN_ + i]);
                        const float gg = pre > 0.f ? v[i] : v[i] * bb.slope;
                        v[i] = gg;
                        sq[i] = gg * (yv - c0[2 * BN_ + i]) * c0[3 * BN_ + i];
                    }
                }
                if (valid) {
                    uint4 u[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        u[q].x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                        u[q].y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                        u[q].z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                        u[q].w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                    }
                    st_global_256(orow + c * 32, u[0], u[1]);          // two full 32-byte sectors per 32 channels
                    st_global_256(orow + c * 32 + 16, u[2], u[3]);
                }
                if (stats != nullptr) {
                    if constexpr (EPIM == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
                    }
                    const float s1 = warp_transpose_sum(v, lane);
                    const float s2 = warp_transpose_sum(sq, lane);
                    rbuf[(wq * 2 + 0) * BN_ + c * 32 + lane] = s1;
                    rbuf[(wq * 2 + 1) * BN_ + c * 32 + lane] = s2;
                }
            }
            if (stats != nullptr) {
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI) : "memory");
                if (tile_live) {
                    const long long key = (long long)(n0 / p.ipg) * 2 * Cout + nt * BN_;   // offset of this tile's block
                    if (key != acc_key) {
                        flush_stats();
                        acc_key = key;
                    }
#pragma unroll
                    for (int j = 0; j < (2 * BN_ + 127) / 128; ++j) {
                        const int col = (threadIdx.x - 64) + j * 128;
                        if (col < 2 * BN_) {
                            const int which = col / BN_, cc = col % BN_;
                            acc_stat[j] += rbuf[(0 * 2 + which) * BN_ + cc] + rbuf[(1 * 2 + which) * BN_ + cc] +
                                           rbuf[(2 * 2 + which) * BN_ + cc] + rbuf[(3 * 2 + which) * BN_ + cc];
                        }
                    }
                }
            }
        }
        if (stats != nullptr) flush_stats();
    }

    fence_before_sync();
    cluster_sync_all();                    // the peer may still be signalling this CTA's barriers / reading its smem
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc_2cta(tmem_base, 2 * kAccCols);
    }
}

template <int BN_, int STAGES, int MODE, int EPIM>
int launch_pair_cfg(const CUtensorMap& mA, const CUtensorMap& mB, void* out, float* stats, const ConvTcParams& p,
                    int m_tiles, int n_tiles, const BnBwdEpi& bb, cudaStream_t st) {
    using L = PairSmem<BN_, STAGES>;
    static DeviceOnce configured;
    auto kern = conv_tc_pair_kernel<BN_, STAGES, MODE, EPIM>;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "conv_tc_pair smem attr: %s", cudaGetErrorString(e));
        configured.mark();
    }
    const int total_pairs = ((m_tiles + 1) / 2) * n_tiles * (MODE == kUpM ? 4 : 1);
    const int max_clusters = BN_ == 256 ? kNumSMs / 2 : kNumSMs;            // BN = 256 fills TMEM: one CTA per SM; else two
    const int clusters = total_pairs < max_clusters ? total_pairs : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mA, mB, (__nv_bfloat16*)out, stats, p, n_tiles, total_pairs, bb);
    if (e != cudaSuccess) return set_error(JCK_E_CUDA, "conv_tc_pair launch: %s", cudaGetErrorString(e));
    JCK_LAUNCH_CHECK(MODE == kUpM ? "conv_up_tc_pair" : "conv_down_tc_pair");
    return JCK_OK;
}

template <int BN_, int STAGES, int MODE, int EPI, int CTAS_PER_SM, int EPIM>
int launch_persist_cfg(const CUtensorMap& mA, const CUtensorMap& mB, void* out, float* stats, const ConvTcParams& p,
                       int m_tiles, int n_tiles, const BnBwdEpi& bb, cudaStream_t st) {
    using L = PersistSmem<BN_, STAGES>;
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_persist_kernel<BN_, STAGES, MODE, EPI, EPIM>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "conv_tc_persist smem attr: %s", cudaGetErrorString(e));
        configured.mark();
    }
    const int total = m_tiles * n_tiles * (MODE == kUpM ? 4 : 1);
    const int cap = kNumSMs * CTAS_PER_SM;
    const int grid = total < cap ? total : cap;
    cudaError_t e = launch_pdl(conv_tc_persist_kernel<BN_, STAGES, MODE, EPI, EPIM>, dim3(grid), dim3(64 + 32 * EPI), L::kTotal, st,
                               mA, mB, (__nv_bfloat16*)out, stats, p, n_tiles, total, bb);
    if (e != cudaSuccess) return set_error(JCK_E_CUDA, "conv_tc_persist launch: %s", cudaGetErrorString(e));
    return JCK_OK;
}

// two co-resident CTAs per SM, four epilogue warps each: one CTA's epilogue and TMA latency hide behind the other's MMAs
template <int BN_, int STAGES, int MODE>
int launch_conv_tc_persist(const CUtensorMap& mA, const CUtensorMap& mB, void* out, float* stats, const ConvTcParams& p,
                           int m_tiles, int n_tiles, cudaStream_t st) {
    const BnBwdEpi none{nullptr, nullptr, nullptr, 0.f};
    int rc = launch_persist_cfg<BN_, STAGES, MODE, 4, 2, 0>(mA, mB, out, stats, p, m_tiles, n_tiles, none, st);
    if (rc) return rc;
    JCK_LAUNCH_CHECK("edge_up_tc_persist");
    return JCK_OK;
}

// ------------------------------------------------------------------------------------------------
// Up conv with an in-shared-memory input WINDOW and a RESIDENT filter bank (the 64-channel layer: G.conv4 forward,
// D.conv2 input gradient; Ca <= 128 -> Cb = 64).
// The tap-streaming kernels above re-load the activation tile once per (output parity, tap): 16 boxes of 16 KB per
// 128 input pixels and 64-channel chunk, which pins the 64-channel layer (only 4 KB of weights per box to amortise
// them) to the L2 -> shared-memory stream at ~600-700 TFLOP/s.  Here
//  * a tile is 8 x 16 input pixels of ONE image and its 10 x 18 pixel neighbourhood (1-pixel halo; out-of-image pixels
//    are zero-filled by TMA = the padding) is loaded ONCE per channel chunk: 180 rows of 128 bytes, 128B-swizzled.
//    Every one of the 16 (parity, tap) products reads its shifted 8 x 16 sub-window straight from that buffer through
//    the MMA's shared-memory descriptor:   operand row (yl, xl) = window row (yl + 1 + dy) * 10 + (xl + 1 + dx),
//    i.e. start address = window + ((1 + dy) * 10 + 1 + dx) * 128 B with the 8-row groups (one image row each) 1280 B
//    apart (the descriptor's stride-byte-offset field).  The swizzle is a function of the absolute shared-memory
//    address on sm_100a (see incep.cu), so a start address off the 1 KB swizzle atom needs no base-offset bits.
//    Activation traffic drops 11x (23 KB instead of 256 KB per tile and chunk);
//  * two CTAs of a cluster (cta_group::2, M = 256) take the two 8-pixel halves of a 16-pixel-wide strip, and each keeps
//    ITS HALF of the layer's whole filter bank (all 16 (parity, tap) x chunk tiles of [32 out][64 k], 128 KB) resident
//    for its entire tile walk: after the prologue the only shared-memory traffic from L2 is the 46 KB window per tile,
//    and the main loop is 128 back-to-back MMAs per tile with no operand barrier in between;
//  * products that read the SAME window shift are issued as ONE MMA over the concatenated output columns of the parities
//    that use it: the centre shift serves all four parities (N = 256), three of the four edge shifts serve two parities
//    whose accumulators are adjacent in the column order P00 | P01 | P11 | P10 (N = 128), the rest are N = 64 -- ten
//    MMAs per 16-deep K step instead of sixteen N = 64 ones (measured: a stream of N = 64 MMAs keeps the tensor pipe
//    at 40 % -- the per-instruction A-operand fetch is not amortised);
//  * windows and the four 64-column parity accumulators are double buffered (2 x 256 TMEM columns), so the epilogue of
//    tile i (each thread writes its input pixel's 2 x 2 output block: two runs of 256 contiguous bytes; BatchNorm
//    sums over the four parities before ONE pair of warp transposes per 32 channels) overlaps the MMAs of tile i + 1.
// One CTA pair per SM pair (the resident bank + two windows fill the 227 KB).
// ------------------------------------------------------------------------------------------------
constexpr int kWinTW = 8, kWinTH = 16;                       // tile: 8 x 16 input pixels = 128 MMA rows per CTA
constexpr int kWinBW = kWinTW + 2, kWinBH = kWinTH + 2;      // with the halo
constexpr int kWinBytes = kWinBW * kWinBH * 128;             // 23,040 B per 64-channel chunk
constexpr int kWinPitch = 23 * 1024;                         // chunk buffers on 1 KB boundaries
constexpr int kWinMaxChunks = 2;                             // Ca <= 128
constexpr int kWinWTile = 32 * 128;                          // this CTA's half of one [64 out][64 k] weight tile
constexpr int kWinWOff = 2 * kWinMaxChunks * kWinPitch;      // after the two window buffers
constexpr int kWinBarOff = kWinWOff + 16 * kWinMaxChunks * kWinWTile;
constexpr int kWinRedOff = kWinBarOff + 256;
constexpr int kWinCoefOff = kWinRedOff + 4 * 2 * 64 * 4;     // [scale | shift | mean | rstd][64]  (fused BatchNorm-backward epilogue)
constexpr int kWinSmem = kWinCoefOff + 4 * 64 * 4 + 1024;

struct ConvWinParams { int B, Hs, Ws, Ca, tiles_x, tiles_img, ipg; };

// accumulator column order P00 | P01 | P11 | P10 (a Gray cycle: neighbours share an output-row or output-column parity, so
// the parities that read one edge shift of the window are adjacent) -> output parity index py * 2 + px
__host__ __device__ __forceinline__ int win_slot_phase(int slot) { return slot == 2 ? 3 : (slot == 3 ? 2 : slot); }

// EPIM = 1: the BatchNorm-backward reduction of the layer this convolution's output feeds (see BnBwdEpi) in the epilogue:
// g = acc * act'(y * scale + shift) is stored instead of acc and stats receives (sum g, sum g * xhat) per group.  The saved
// y has the output's layout, so a thread reads exactly the 64-byte runs it is about to write; the epilogue of this kernel
// waits for the MMAs most of the time (it is not the critical path), which is what makes the fusion pay here.
template <int EPIM>
__global__ void __launch_bounds__(320, 1)
conv_up_win_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ stats, const ConvWinParams p, const int total_pairs,
                   const BnBwdEpi bb) {
    pdl_trigger();
    constexpr int Cb = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kWinBarOff);   // [2]  leader: both CTAs' window bytes have landed
    uint64_t* a_empty = a_full + 2;                                      // [2]  each CTA: the tile's MMAs have read the window
    uint64_t* tfull = a_empty + 2;                                       // [2]  each CTA: accumulators complete
    uint64_t* tempty = tfull + 2;                                        // [2]  leader: accumulators drained by both epilogues
    uint64_t* w_full = tempty + 2;                                       // leader: both CTAs' filter-bank halves have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* red = reinterpret_cast<float*>(smem + kWinRedOff);            // [4 warps][2][64]
    float* cf = reinterpret_cast<float*>(smem + kWinCoefOff);            // [4][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int cchunks = p.Ca / kBK;
    const int wtiles = 16 * cchunks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&a_full[a], 1);
            mbar_init(&a_empty[a], 1);
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 2 * 256);
        }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
    fence_before_sync();
    cluster_sync_all();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // tile of pair-tile tp for this CTA: the two CTAs take x-adjacent tiles (tiles_x is even)
    auto decode = [&](int tp, int& n, int& x0, int& y0) {
        const int t = 2 * tp + rank;
        n = t / p.tiles_img;
        const int rem = t - n * p.tiles_img;
        x0 = (rem % p.tiles_x) * kWinTW;
        y0 = (rem / p.tiles_x) * kWinTH;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer (both CTAs) ----------------
            // resident filter bank, per channel chunk 64 KB = this CTA's half (cta_group::2: rows [r N/2, (r+1) N/2) of every
            // MMA's B tile) of the ten MMA groups, as sixteen [32 out][64 k] boxes -- see the group table at the MMA issuer
            if (rank == 0) mbar_arrive_expect_tx(w_full, 2 * wtiles * kWinWTile);
            for (int cc = 0; cc < cchunks; ++cc) {
                for (int q = 0; q < 16; ++q) {
                    int slot, half, tap;               // accumulator slot (column order), which 32-row half of it, filter tap
                    if (q < 4) { slot = 2 * rank + (q >> 1); half = q & 1; tap = 0; }             // centre shift, N = 256
                    else if (q < 6) { slot = 0 + rank; half = q & 1; tap = 2; }                 // (-1, 0): slots 0,1
                    else if (q < 8) { slot = 1 + rank; half = q & 1; tap = 1; }                 // (0, +1): slots 1,2
                    else if (q < 10) { slot = 2 + rank; half = q & 1; tap = 2; }                // (+1, 0): slots 2,3
                    else if (q == 10) { slot = 3; half = rank; tap = 1; }                       // (0, -1): slot 3
                    else if (q == 11) { slot = 0; half = rank; tap = 1; }                       // (0, -1): slot 0
                    else { slot = q - 12; half = rank; tap = 3; }                               // the four corners
                    const int phase = win_slot_phase(slot);
                    tma_load_2d_2cta(smem + kWinWOff + (cc * 16 + q) * kWinWTile, &mapB, w_full, tap * p.Ca + cc * kBK,
                                     phase * Cb + half * 32);
                }
            }
            int lt = 0;
            for (int tp = cid; tp < total_pairs; tp += ncl, ++lt) {
                int n, x0, y0;
                decode(tp, n, x0, y0);
                const int buf = lt & 1;
                mbar_wait(&a_empty[buf], ((lt >> 1) & 1) ^ 1);     // the MMAs of the tile two back have read this window
                if (rank == 0) mbar_arrive_expect_tx(&a_full[buf], 2 * cchunks * kWinBytes);
                for (int cc = 0; cc < cchunks; ++cc)
                    tma_load_4d_2cta(smem + (buf * kWinMaxChunks + cc) * kWinPitch, &mapA, &a_full[buf], cc * kBK, x0 - 1, y0 - 1, n);
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            // ---------------- MMA issuer: ONE thread of the leader CTA, 128 MMAs per tile back to back ----------------
            constexpr uint32_t idesc64 = make_idesc(64, 0, 0, 256), idesc128 = make_idesc(128, 0, 0, 256),
                               idesc256 = make_idesc(256, 0, 0, 256);
            const uint64_t a_desc0 = make_sdesc(smem_u32(smem), 0, kWinBW * 128);
            const uint64_t b_desc0 = make_sdesc(smem_u32(smem + kWinWOff), 0, 1024);
            mbar_wait(w_full, 0);
            int lt = 0;
            for (int tp = cid; tp < total_pairs; tp += ncl, ++lt) {
                const int buf = lt & 1;
                mbar_wait(&tempty[buf], ((lt >> 1) & 1) ^ 1);      // accumulator set drained by both CTAs' epilogues
                mbar_wait(&a_full[buf], (lt >> 1) & 1);
                fence_after_sync();
                const uint32_t tmem_d = tmem_base + buf * 256;
                for (int cc = 0; cc < cchunks; ++cc) {
                    const uint64_t a_c = a_desc0 + (uint64_t)(((buf * kWinMaxChunks + cc) * kWinPitch) >> 4);
                    const uint64_t b_c = b_desc0 + (uint64_t)((cc * 16 * kWinWTile) >> 4);
                    // group: window shift (dy, dx) | first accumulator slot | N | offset of its B tile in the chunk's bank
                    auto group = [&](int dy, int dx, int slot, int n, int b_off, uint32_t idesc_n, bool first) {
                        const uint64_t a_d = a_c + (uint64_t)((((1 + dy) * kWinBW + 1 + dx) * 128) >> 4);
                        const uint64_t b_d = b_c + (uint64_t)(b_off >> 4);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            umma_bf16_2cta(tmem_d + slot * Cb, a_d + (uint64_t)(k * 2), b_d + (uint64_t)(k * 2), idesc_n,
                                           (first && cc == 0 && k == 0) ? 0u : 1u);
                        (void)n;
                    };
                    group(0, 0, 0, 256, 0, idesc256, true);                 // centre: P00 | P01 | P11 | P10, tap (0,0)
                    group(-1, 0, 0, 128, 4 * kWinWTile, idesc128, false);   // rows above: P00 | P01, tap (1,0)
                    group(0, 1, 1, 128, 6 * kWinWTile, idesc128, false);    // right: P01 | P11, tap (0,1)
                    group(1, 0, 2, 128, 8 * kWinWTile, idesc128, false);    // below: P11 | P10, tap (1,0)
                    group(0, -1, 3, 64, 10 * kWinWTile, idesc64, false);    // left: P10, tap (0,1)
                    group(0, -1, 0, 64, 11 * kWinWTile, idesc64, false);    // left: P00, tap (0,1)
                    group(-1, -1, 0, 64, 12 * kWinWTile, idesc64, false);   // corners, tap (1,1)
                    group(-1, 1, 1, 64, 13 * kWinWTile, idesc64, false);
                    group(1, 1, 2, 64, 14 * kWinWTile, idesc64, false);
                    group(1, -1, 3, 64, 15 * kWinWTile, idesc64, false);
                }
                umma_commit_2cta(&a_empty[buf], 3);
                umma_commit_2cta(&tfull[buf], 3);
            }
        }
    } else {
        // ---------------- epilogue: EIGHT warps.  A thread = one input pixel (yl, xl) of this CTA's tile = one TMEM lane, its 2 x 2
        // output block, and ONE 32-channel half of the 64 output channels (warps 2..5 the first half, 6..9 the second: two warps
        // share a TMEM lane quarter, warp % 4).  Four warps could not keep up with the MMAs once the epilogue also does the
        // BatchNorm-backward arithmetic (~10 operations per element). ----------------
        const int wq = warp & 3;
        const int c = (warp - 2) >> 2;                       // which 32-channel half
        const int r = wq * 32 + lane;
        const int yl = r >> 3, xl = r & 7;
        const int e = threadIdx.x - 64;                      // 0..255; threads < 128 flush column e of [sum | sum of squares] x 64
        float acc_stat = 0.f;
        int acc_group = -1, cf_group = -1;
        int lt = 0;
        for (int tp = cid; tp < total_pairs; tp += ncl, ++lt) {
            int n, x0, y0;
            decode(tp, n, x0, y0);
            const int buf = lt & 1;
            const int grp = n / p.ipg;
            if constexpr (EPIM == 1) {
                if (grp != cf_group) {                       // this group's BatchNorm coefficients (rarely changes: tiles walk images in order)
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    {
                        const int which = e / Cb, ch = e - which * Cb;
                        cf[e] = ((which < 2) ? bb.ss : bb.mr)[(size_t)grp * 2 * Cb + (which & 1) * Cb + ch];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    cf_group = grp;
                }
            }
            const uint32_t tmem_d = tmem_base + buf * 256 + ((uint32_t)(wq * 32) << 16) + c * 32;
            const size_t Wo = 2 * (size_t)p.Ws;
            const size_t o00_off = (((size_t)n * 2 * p.Hs + 2 * (y0 + yl)) * Wo + 2 * (x0 + xl)) * Cb + c * 32;
            __nv_bfloat16* o00 = out + o00_off;
            auto out_off = [&](int slot) {
                const int phase = win_slot_phase(slot);
                return ((size_t)(phase >> 1) * Wo + (phase & 1)) * Cb;
            };
            if constexpr (EPIM == 1) {
                // the NEXT tile's saved y -> L2 now (its registers-loads then pay an L2 hit, not an HBM round trip)
                if (tp + ncl < total_pairs) {
                    int n2, x2, y2;
                    decode(tp + ncl, n2, x2, y2);
                    const __nv_bfloat16* yn = bb.y + (((size_t)n2 * 2 * p.Hs + 2 * (y2 + yl)) * Wo + 2 * (x2 + xl)) * Cb + c * 32;
#pragma unroll
                    for (int slot = 0; slot < 4; ++slot)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(yn + out_off(slot)));
                }
            }
            // EPIM = 1: the saved y of this thread's four 64-byte runs.  The first two are requested BEFORE the accumulator wait
            // (they do not depend on the MMAs), each register buffer is re-filled two iterations ahead of its next use (ten warps
            // per CTA cap a thread at 168 registers: two buffers, not four)
            uint4 yr[2][4];
            if constexpr (EPIM == 1) {
#pragma unroll
                for (int slot = 0; slot < 2; ++slot) {
                    const __nv_bfloat16* ys = bb.y + o00_off + out_off(slot);
                    ld_global_nc_256(ys, yr[slot][0], yr[slot][1]);
                    ld_global_nc_256(ys + 16, yr[slot][2], yr[slot][3]);
                }
            }
            mbar_wait(&tfull[buf], (lt >> 1) & 1);
            fence_after_sync();
            float s1[32], s2[32];
#pragma unroll
            for (int slot = 0; slot < 4; ++slot) {
                float v[32];
                tmem_ld32(tmem_d + slot * Cb, v);
                tmem_ld_wait();
                if (slot == 3) {                             // last read of this accumulator set: hand it back to the leader
                    fence_before_sync();
                    mbar_arrive_leader(&tempty[buf]);
                }
                if constexpr (EPIM == 1) {
                    const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(yr[slot & 1]);
                    const float4* sc4 = reinterpret_cast<const float4*>(cf + c * 32);          // 128-bit broadcast reads
                    const float4* sh4 = reinterpret_cast<const float4*>(cf + Cb + c * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 sc = sc4[q], sh = sh4[q];
                        const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = q * 4 + j;
                            const float yv = __bfloat162float(yb[i]);
                            const float pre = fmaf(yv, scv[j], shv[j]);
                            const float gg = pre > 0.f ? v[i] : v[i] * bb.slope;
                            v[i] = gg;
                            // sum g (y - mean) rstd = rstd (sum g y - mean sum g): mean and rstd once per column below
                            if (slot == 0) { s1[i] = gg; s2[i] = gg * yv; } else { s1[i] += gg; s2[i] = fmaf(gg, yv, s2[i]); }
                        }
                    }
                    if (slot < 2) {
                        const __nv_bfloat16* ys = bb.y + o00_off + out_off(slot + 2);
                        ld_global_nc_256(ys, yr[slot & 1][0], yr[slot & 1][1]);
                        ld_global_nc_256(ys + 16, yr[slot & 1][2], yr[slot & 1][3]);
                    }
                } else if (stats != nullptr) {
                    if (slot == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { s1[i] = v[i]; s2[i] = v[i] * v[i]; }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { s1[i] += v[i]; s2[i] = fmaf(v[i], v[i], s2[i]); }
                    }
                }
                uint4 u[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    u[q].x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                    u[q].y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                    u[q].z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                    u[q].w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                }
                __nv_bfloat16* orow = o00 + out_off(slot);
                st_global_256(orow, u[0], u[1]);
                st_global_256(orow + 16, u[2], u[3]);
            }
            if (stats != nullptr) {
                const float a1 = warp_transpose_sum(s1, lane);
                float a2 = warp_transpose_sum(s2, lane);
                if constexpr (EPIM == 1) a2 = (a2 - cf[2 * Cb + c * 32 + lane] * a1) * cf[3 * Cb + c * 32 + lane];
                red[(wq * 2 + 0) * Cb + c * 32 + lane] = a1;
                red[(wq * 2 + 1) * Cb + c * 32 + lane] = a2;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (e < 2 * Cb) {
                    if (grp != acc_group) {
                        if (acc_group >= 0) atomicAdd(stats + (size_t)acc_group * 2 * Cb + e, acc_stat);
                        acc_group = grp;
                        acc_stat = 0.f;
                    }
                    const int which = e / Cb, cc = e % Cb;
                    acc_stat += red[(0 * 2 + which) * Cb + cc] + red[(1 * 2 + which) * Cb + cc] +
                                red[(2 * 2 + which) * Cb + cc] + red[(3 * 2 + which) * Cb + cc];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");   // red[] is rewritten by the next tile
            }
        }
        if (stats != nullptr && e < 2 * Cb && acc_group >= 0) atomicAdd(stats + (size_t)acc_group * 2 * Cb + e, acc_stat);
    }

    fence_before_sync();
    cluster_sync_all();                    // the peer may still be signalling this CTA's barriers / reading its smem
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

bool conv_up_win_supported(int B, int Hs, int Ws, int Ca, int Cb) {
    static const bool enabled = [] { const char* e = getenv("JCK_UP_WIN"); return !(e && e[0] == '0'); }();
    return enabled && B > 0 && Cb == 64 && Ca % kBK == 0 && Ca / kBK <= kWinMaxChunks && Ws % (2 * kWinTW) == 0 &&
           Hs % kWinTH == 0;
}

int conv_up_win(const void* in, const void* w, void* out, float* stats, int B, int Hs, int Ws, int Ca, int ipg, cudaStream_t st,
                const BnBwdEpi* bb) {
    CUtensorMap mA, mB;
    int rc;
    if ((rc = map_small(&mA, in, Ca, Ws, Hs, B, kWinBW, kWinBH, 1))) return rc;
    if ((rc = map_matrix(&mB, w, 4 * 64, 4 * Ca, 32))) return rc;
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(conv_up_win_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWinSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_up_win_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWinSmem);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "conv_up_win smem attr: %s", cudaGetErrorString(e));
        configured.mark();
    }
    ConvWinParams p{B, Hs, Ws, Ca, Ws / kWinTW, (Ws / kWinTW) * (Hs / kWinTH), ipg};
    const int total_pairs = B * p.tiles_img / 2;
    const int clusters = total_pairs < kNumSMs / 2 ? total_pairs : kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = kWinSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    const BnBwdEpi none{nullptr, nullptr, nullptr, 0.f};
    cudaError_t e = bb ? cudaLaunchKernelEx(&cfg, conv_up_win_kernel<1>, mA, mB, (__nv_bfloat16*)out, stats, p, total_pairs, *bb)
                       : cudaLaunchKernelEx(&cfg, conv_up_win_kernel<0>, mA, mB, (__nv_bfloat16*)out, stats, p, total_pairs, none);
    if (e != cudaSuccess) return set_error(JCK_E_CUDA, "conv_up_win launch: %s", cudaGetErrorString(e));
    JCK_LAUNCH_CHECK("conv_up_win");
    return JCK_OK;
}

bool tc_conv_supported(int B, int Hs, int Ws, int Ca, int Cb, int ipg, bool up, PatchGeom* g) {
    const int Cin = up ? Ca : Cb, Cout = up ? Cb : Ca;
    if (Cin % 64 != 0 || Cout % 64 != 0) return false;
    if (!patch_geom(Hs, Ws, kTileM, g)) return false;
    if (ipg < B && (ipg % g->nb) != 0) return false;   // a tile must not straddle BatchNorm groups
    return B > 0;
}

template <bool kUp>
int conv_tc(const void* in, const void* w, void* out, float* stats, int B, int Hs, int Ws, int Ca, int Cb, int ipg,
            cudaStream_t st, const BnBwdEpi* bb = nullptr) {
    PatchGeom g;
    if (!tc_conv_supported(B, Hs, Ws, Ca, Cb, ipg, kUp, &g))
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv tc: unsupported shape B=%d Hs=%d Ws=%d Ca=%d Cb=%d", B, Hs, Ws, Ca, Cb);
    if (kUp && conv_up_win_supported(B, Hs, Ws, Ca, Cb)) return conv_up_win(in, w, out, stats, B, Hs, Ws, Ca, ipg, st, bb);
    const int Cout = kUp ? Cb : Ca;
    const int bn = (Cout % 128 == 0) ? 128 : 64;
    ConvTcParams p{B, Hs, Ws, Ca, Cb, g.bw, g.bh, g.nb, Ws / g.bw, Hs / g.bh, ipg};
    const int m_tiles = p.tiles_x * p.tiles_y * ((B + g.nb - 1) / g.nb);
    CUtensorMap mA, mB;
    int rc;
    // cta_group::2: every CTA of a pair stages half of the weight tile -> B box of bn / 2 rows
    if (!kUp) {
        if ((rc = map_large(&mA, in, Cb, 2 * Ws, 2 * Hs, B, g.bw, g.bh, g.nb))) return rc;
        if ((rc = map_matrix(&mB, w, Ca, 16 * Cb, bn / 2))) return rc;
    } else {
        if ((rc = map_small(&mA, in, Ca, Ws, Hs, B, g.bw, g.bh, g.nb))) return rc;
        if ((rc = map_matrix(&mB, w, 4 * Cb, 4 * Ca, bn / 2))) return rc;
    }
    const BnBwdEpi none{nullptr, nullptr, nullptr, 0.f};
    constexpr int M_ = kUp ? kUpM : kDown;
    // JCK_BN256: 0 = never, 1 = always, unset = where it measured faster (>= 1024 images: 5-10 % on the 256- / 512-channel
    // layers of the discriminator's three-group forward; +-1 us at 512 images)
    static const int bn256 = [] { const char* e = getenv("JCK_BN256"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    if (Cout % 256 == 0 && (bn256 == 1 || (bn256 < 0 && B >= 1024))) {
        // 256-column tiles: 131 FLOP per byte staged instead of 87; both accumulators fill TMEM, one CTA pair per SM pair
        CUtensorMap mB2;
        if (!kUp) { if ((rc = map_matrix(&mB2, w, Ca, 16 * Cb, 128))) return rc; }
        else { if ((rc = map_matrix(&mB2, w, 4 * Cb, 4 * Ca, 128))) return rc; }
        if (bb) return launch_pair_cfg<256, 6, M_, 1>(mA, mB2, out, stats, p, m_tiles, Cout / 256, *bb, st);
        return launch_pair_cfg<256, 6, M_, 0>(mA, mB2, out, stats, p, m_tiles, Cout / 256, none, st);
    }
    if (bn == 128) {
        if (bb) return launch_pair_cfg<128, 4, M_, 1>(mA, mB, out, stats, p, m_tiles, Cout / 128, *bb, st);
        return launch_pair_cfg<128, 4, M_, 0>(mA, mB, out, stats, p, m_tiles, Cout / 128, none, st);
    }
    if (bb) return launch_pair_cfg<64, 4, M_, 1>(mA, mB, out, stats, p, m_tiles, Cout / 64, *bb, st);
    return launch_pair_cfg<64, 4, M_, 0>(mA, mB, out, stats, p, m_tiles, Cout / 64, none, st);
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[a][tap, b] = sum over pixels small[pix][a] * large[shift_tap(pix)][b]   (split over pixels)
// Both operands are MN-major (channels contiguous, contraction over rows).  One CTA owns 128 `a`
// channels x (G taps x BNW `b` channels) = 512 TMEM columns and a contiguous range of 64-pixel K steps;
// partial tiles are summed into a [Ca][16*Cb] workspace by TMA reducing stores; wgrad_unpack transposes it into the
// reference's [Ca][Cb][4][4] layout.
// ------------------------------------------------------------------------------------------------
struct WgradTcParams {
    int B, Hs, Ws, Ca, Cb;
    int kbw, kbh, knb;         // 64-pixel K patch
    int kx_tiles, ky_tiles;    // patches per image
    int total_steps, steps_per_split;
    int b_tiles;               // Cb / BNW
};

// Pipeline: two stages of 64 pixels (80 KB each).  Four stages of 32 pixels (same 160 KB, deeper prefetch) measured SLOWER
// (c3 at 512 images: 52 vs 44 us; c4: 58 vs 48 us): twice the barrier round trips and TMA boxes per MMA outweigh the prefetch.
constexpr int kWgradKPix = 64;                           // K step of the image-edge weight gradient (wgrad_edge_direct_kernel)
constexpr int kWgradStages = 2;
constexpr int kWgTcKPix = 64;
constexpr int kWgradABytes = 2 * kWgTcKPix * 128;       // two 64-channel atoms of `a`
constexpr int kWgradBBytes = 8 * kWgTcKPix * 128;       // G * BNW / 64 = 8 atoms
constexpr int kWgradStage = kWgradABytes + kWgradBBytes; // 80 KB
constexpr int kWgradSmem = kWgradStages * kWgradStage + 256 + 1024;

template <int BNW>  // 64 (G = 8 taps) or 128 (G = 4 taps)
__global__ void __launch_bounds__(kConvThreads)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapS, const __grid_constant__ CUtensorMap mapL,
                const __grid_constant__ CUtensorMap mapP, const WgradTcParams p) {
    pdl_trigger();
    constexpr int G = 512 / BNW;
    constexpr int ATOMS_B = BNW / 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgradStages * kWgradStage);
    uint64_t* empty = full + kWgradStages;
    uint64_t* tmem_full = empty + kWgradStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int a_tile = blockIdx.y / p.b_tiles, b_tile = blockIdx.y % p.b_tiles;
    const int tap0 = blockIdx.z * G;
    const int step_beg = split * p.steps_per_split;
    const int step_end = min(p.total_steps, step_beg + p.steps_per_split);
    const int nsteps = max(0, step_end - step_beg);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapS);
        prefetch_tmap(&mapL);
        prefetch_tmap(&mapP);
        for (int s = 0; s < kWgradStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // everything above is on-chip set-up; global memory is first touched below

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % kWgradStages;
                const uint32_t ph = (it / kWgradStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* sa = smem + s * kWgradStage;
                uint8_t* sb = sa + kWgradABytes;
                mbar_arrive_expect_tx(&full[s], kWgradStage);
                const int st = step_beg + it;
                const int xs = (st % p.kx_tiles) * p.kbw;
                const int ys = ((st / p.kx_tiles) % p.ky_tiles) * p.kbh;
                const int ns = (st / (p.kx_tiles * p.ky_tiles)) * p.knb;
                tma_load_4d(sa, &mapS, &full[s], a_tile * 128, xs, ys, ns);
                tma_load_4d(sa + kWgTcKPix * 128, &mapS, &full[s], a_tile * 128 + 64, xs, ys, ns);
#pragma unroll 1
                for (int g = 0; g < G; ++g) {
                    const int tap = tap0 + g;
                    const int ky = tap >> 2, kx = tap & 3;
                    const int dy = (ky - 1) >> 1, qy = (ky - 1) & 1;
                    const int dx = (kx - 1) >> 1, qx = (kx - 1) & 1;
                    for (int at = 0; at < ATOMS_B; ++at)
                        tma_load_5d(sb + (g * ATOMS_B + at) * (kWgTcKPix * 128), &mapL, &full[s],
                                    qx * p.Cb + b_tile * BNW + at * 64, xs + dx, qy, ys + dy, ns);
                }
            }
        }
    } else if (warp == 1) {
        // Every MMA is N = 128: with BNW = 64 the 64-channel atoms of two consecutive taps sit 8 KB apart in the stage and
        // their accumulator columns are adjacent (tap g at column g * 64), so one instruction covers the pair -- the
        // leading-dimension byte offset of the MN-major descriptor is exactly that atom distance.  (N = 64 instructions ran the
        // 64-channel layer at 690 TFLOP/s against 1,010 for the 128-channel ones: half the math per issued instruction.)
        constexpr uint32_t idesc = make_idesc(128, 1, 1);
        constexpr int kGroups = 512 / 128;
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % kWgradStages;
            const uint32_t ph = (it / kWgradStages) & 1;
            mbar_wait(&full[s], ph);
            fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(smem + s * kWgradStage);
                const uint32_t b_addr = a_addr + kWgradABytes;
#pragma unroll 1
                for (int g = 0; g < kGroups; ++g) {
#pragma unroll
                    for (int k = 0; k < kWgTcKPix / 16; ++k) {
                        const uint64_t da = make_sdesc(a_addr + k * 2048, kWgTcKPix * 128, 1024);
                        const uint64_t db = make_sdesc(b_addr + g * 2 * (kWgTcKPix * 128) + k * 2048,
                                                       kWgTcKPix * 128, 1024);
                        umma_bf16(tmem_base + g * 128, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);
                if (it == nsteps - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        // The 128 x 512 fp32 partial tile (256 KB) leaves through shared memory: the pipeline stages are idle once the
        // last MMA has retired, so each 32-column chunk is staged as a 128B-swizzled [128 rows][32 floats] block
        // (conflict-free 16-byte stores) and written by ONE TMA store as full 128-byte lines -- eight chunks per
        // round.  (Thread-per-row float4 stores wrote 16-byte pieces of 32 different lines per instruction.)
        const int wq = warp & 3;
        const int r = wq * 32 + lane;                               // row of the tile = TMEM lane
        const int e = threadIdx.x - 64;
        if (nsteps > 0) {
            mbar_wait(tmem_full, 0);
            fence_after_sync();
        }
        constexpr int kChunks = 16;                                 // 512 columns / 32
        constexpr int kRound = 8;                                   // 8 x 16 KB staging blocks = 128 KB of the 160 KB ring
#pragma unroll 1
        for (int c0 = 0; c0 < kChunks; c0 += kRound) {
            if (c0 > 0) {
                if (e == 0) tma_store_wait_read0();                 // previous round's stores have read their blocks
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
#pragma unroll 1
            for (int j = 0; j < kRound; ++j) {
                const int c = c0 + j;                               // column chunk: g = c / (BNW/32), cc = c % (BNW/32)
                float v[32];
                if (nsteps > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + c * 32, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                uint8_t* blk = smem + j * (128 * 128) + r * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(blk + ((q ^ (r & 7)) << 4)) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e == 0) {
#pragma unroll 1
                for (int j = 0; j < kRound; ++j) {
                    const int c = c0 + j, g = c / (BNW / 32), cc = c % (BNW / 32);
                    tma_reduce_add_2d(&mapP, smem + j * (128 * 128), (tap0 + g) * p.Cb + b_tile * BNW + cc * 32, a_tile * 128);
                }
                tma_store_commit();
            }
        }
        if (e == 0) tma_store_wait_all();
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

struct WgradPlan { bool ok; int bnw, G, splits, steps_per_split, total_steps; PatchGeom g; };

WgradPlan wgrad_plan(int B, int Hs, int Ws, int Ca, int Cb) {
    WgradPlan pl{};
    pl.ok = false;
    if (Ca % 128 != 0 || Cb % 64 != 0) return pl;
    if (!patch_geom(Hs, Ws, kWgTcKPix, &pl.g)) return pl;
    pl.bnw = (Cb % 128 == 0) ? 128 : 64;
    pl.G = 512 / pl.bnw;
    const int tiles = (Ca / 128) * (Cb / pl.bnw) * (16 / pl.G);
    pl.total_steps = (Ws / pl.g.bw) * (Hs / pl.g.bh) * ((B + pl.g.nb - 1) / pl.g.nb);
    int splits = kNumSMs / tiles;
    if (splits < 1) splits = 1;
    if (splits > pl.total_steps) splits = pl.total_steps;
    pl.steps_per_split = (pl.total_steps + splits - 1) / splits;
    pl.splits = (pl.total_steps + pl.steps_per_split - 1) / pl.steps_per_split;
    pl.ok = true;
    return pl;
}

int wgrad_tc(const void* small, const void* large, float* part, const WgradPlan& pl, int B, int Hs, int Ws, int Ca,
             int Cb, cudaStream_t st) {
    CUtensorMap mS, mL, mP;
    int rc;
    if ((rc = map_small(&mS, small, Ca, Ws, Hs, B, pl.g.bw, pl.g.bh, pl.g.nb))) return rc;
    if ((rc = map_large(&mL, large, Cb, 2 * Ws, 2 * Hs, B, pl.g.bw, pl.g.bh, pl.g.nb))) return rc;
    // all splits reduce into ONE [Ca][16*Cb] fp32 matrix with TMA reducing stores (the adds resolve in L2): the 148 partial
    // tiles of 256 KB (38 MB written and read back per layer) never reach HBM
    if ((rc = map_matrix_f32(&mP, part, (long long)Ca, 16 * Cb, 128))) return rc;
    {
        cudaError_t e = cudaMemsetAsync(part, 0, (size_t)Ca * 16 * Cb * sizeof(float), st);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "wgrad_tc memset: %s", cudaGetErrorString(e));
    }
    WgradTcParams p{B, Hs, Ws, Ca, Cb, pl.g.bw, pl.g.bh, pl.g.nb, Ws / pl.g.bw, Hs / pl.g.bh,
                    pl.total_steps, pl.steps_per_split, Cb / pl.bnw};
    dim3 grid(pl.splits, (Ca / 128) * (Cb / pl.bnw), 16 / pl.G);
    static DeviceOnce cfg64, cfg128;
    if (pl.bnw == 64) {
        if (!cfg64.done()) {
            cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem);
            if (e != cudaSuccess) return set_error(JCK_E_CUDA, "wgrad_tc smem attr: %s", cudaGetErrorString(e));
            cfg64.mark();
        }
        launch_pdl(wgrad_tc_kernel<64>, dim3(grid), dim3(kConvThreads), kWgradSmem, st, mS, mL, mP, p);
    } else {
        if (!cfg128.done()) {
            cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem);
            if (e != cudaSuccess) return set_error(JCK_E_CUDA, "wgrad_tc smem attr: %s", cudaGetErrorString(e));
            cfg128.mark();
        }
        launch_pdl(wgrad_tc_kernel<128>, dim3(grid), dim3(kConvThreads), kWgradSmem, st, mS, mL, mP, p);
    }
    JCK_LAUNCH_CHECK("wgrad_tc");
    return JCK_OK;
}

// ------------------------------------------------------------------------------------------------
// image-edge wgrad: D[a][(ky,kx,c4)] = sum over pixels small[pix][a] * patch[pix][(ky,kx,c4)]
// Ca = 64 -> M = 64: issued as an M = 128 MMA whose second 64-row atom is a block of zeros (LBO points at it);
// TMEM lanes 64..127 then hold zeros that the epilogue ignores.  N = 64 = the whole patch, K = 64 pixels
// per step, split over the grid; partials [split][64][64] are reduced by edge_wgrad_unpack.
// The patch operand is built in shared memory from the padded image itself (see
// edge_down_direct_kernel): per 64-pixel K step the producer bulk-copies the 2*rps + 2 image rows next to the TMA
// box of `small`, warps 2..5 re-pack them into the swizzled [64 pixels][64] tile (half a pixel row per thread),
// and the MMA warp consumes the stage once both the TMA bytes and the re-pack have landed.
struct EdgeWgradDirectParams { int total_steps, steps_per_split, Hs, Ws, rps, raw_bytes, raw_stride; };
constexpr int kEdgeWDStages = 4;

__global__ void __launch_bounds__(kConvThreads)
wgrad_edge_direct_kernel(const __grid_constant__ CUtensorMap mapS, const __nv_bfloat16* __restrict__ img,
                         float* __restrict__ part, const EdgeWgradDirectParams p) {
    pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kTile = kWgradKPix * 128;                        // 8 KB
    uint8_t* sSmall = smem;                                        // [stages][8 KB]
    uint8_t* sPatch = sSmall + kEdgeWDStages * kTile;              // [stages][8 KB]
    uint8_t* sZero = sPatch + kEdgeWDStages * kTile;               // 8 KB of zeros (upper 64-row atom of the M = 128 MMA)
    uint8_t* sRaw = sZero + kTile;                                 // [stages][raw_stride]
    uint64_t* full = reinterpret_cast<uint64_t*>(sRaw + kEdgeWDStages * p.raw_stride);
    uint64_t* empty = full + kEdgeWDStages;
    uint64_t* ready = empty + kEdgeWDStages;
    uint64_t* tmem_full = ready + kEdgeWDStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int step_beg = split * p.steps_per_split;
    const int nsteps = max(0, min(p.total_steps, step_beg + p.steps_per_split) - step_beg);
    const int Wp = 2 * p.Ws + 2, Hp = 2 * p.Hs + 2;
    for (int i = threadIdx.x; i < kTile / 16; i += kConvThreads) reinterpret_cast<uint4*>(sZero)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapS);
        for (int s = 0; s < kEdgeWDStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 128); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 64);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % kEdgeWDStages;
                mbar_wait(&empty[s], ((it / kEdgeWDStages) & 1) ^ 1);
                const int st = step_beg + it;
                const int row = st * p.rps;                         // global output row index (n * Hs + oy0)
                const int n = row / p.Hs, oy0 = row - n * p.Hs;
                mbar_arrive_expect_tx(&full[s], kTile + p.raw_bytes);
                tma_load_2d(sSmall + s * kTile, &mapS, &full[s], 0, st * kWgradKPix);
                bulk_load_1d(sRaw + s * p.raw_stride, img + ((size_t)n * Hp + 2 * oy0) * Wp * 4, p.raw_bytes, &full[s]);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(64, 1, 1);
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % kEdgeWDStages;
            const uint32_t ph = (it / kEdgeWDStages) & 1;
            mbar_wait(&full[s], ph);                               // TMA bytes of `small`
            mbar_wait(&ready[s], ph);                              // re-packed patches
            fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(sSmall + s * kTile);
                const uint32_t b_addr = smem_u32(sPatch + s * kTile);
                const uint32_t lbo = smem_u32(sZero) - a_addr;
#pragma unroll
                for (int k = 0; k < kWgradKPix / 16; ++k)
                    umma_bf16(tmem_base, make_sdesc(a_addr + k * 2048, lbo, 1024), make_sdesc(b_addr + k * 2048, 0, 1024),
                              idesc, (it > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty[s]);
                if (it == nsteps - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        const int e = threadIdx.x - 64;                            // 0..127
        const int r = e >> 1, half = e & 1;                        // pixel row of the K step, which two filter rows
        const int ox = r % p.Ws, oyl = r / p.Ws;
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % kEdgeWDStages;
            mbar_wait(&full[s], (it / kEdgeWDStages) & 1);
            const uint8_t* raw = sRaw + s * p.raw_stride;
            uint8_t* dst = sPatch + s * kTile + r * 128;
            uint4 q[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int ky = 2 * half + j;
                const uint4* src = reinterpret_cast<const uint4*>(raw + ((size_t)(2 * oyl + ky) * Wp + 2 * ox) * 8);
                q[2 * j] = src[0];
                q[2 * j + 1] = src[1];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * half + j;
                *reinterpret_cast<uint4*>(dst + ((c ^ (r & 7)) << 4)) = q[j];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&ready[s]);
        }
        const int wq = warp & 3;
        if (wq < 2) {                                              // TMEM lanes 0..63 carry the 64 real rows
            const int a = wq * 32 + lane;
            if (nsteps > 0) { mbar_wait(tmem_full, 0); fence_after_sync(); }
            float* drow = part + ((size_t)split * 64 + a) * 64;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                float v[32];
                if (nsteps > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + c * 32, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                float4* d4 = reinterpret_cast<float4*>(drow + c * 32);
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) d4[qq] = make_float4(v[qq * 4], v[qq * 4 + 1], v[qq * 4 + 2], v[qq * 4 + 3]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 64);
    }
}

// dw4[a][c][ky][kx] (+)= sum_split part[split][a][ky*16 + kx*4 + c], c < nc
__global__ void edge_wgrad_unpack_kernel(const float* __restrict__ part, float* __restrict__ dw4, int nc, int splits,
                                         int accumulate) {
    pdl_entry();
    const int total = 64 * nc * 16;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int kx = idx & 3, ky = (idx >> 2) & 3, c = (idx >> 4) % nc, a = (idx >> 4) / nc;
        const int col = ky * 16 + kx * 4 + c;
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[((size_t)z * 64 + a) * 64 + col];
        dw4[idx] = accumulate ? dw4[idx] + s : s;
    }
}

// ------------------------------------------------------------------------------------------------
// image-edge down conv straight from the padded 4-channel image (no patch matrix in HBM).
// A 128-pixel output tile (bh rows x Ws columns of one image) needs 2*bh + 2 CONSECUTIVE rows of the P4 image:
// one contiguous run of (2*bh + 2) * (2*Ws + 2) * 8 bytes (5280 B for a 4 x 32 tile vs the 16 KB patch tile), so
//   warp 0     : one 1-D bulk copy per tile into a small raw ring (runs ahead of the consumers),
//   warps 2..5 : thread r re-packs the four 32-byte runs of ITS output pixel from the raw rows into row r of the
//                128B-swizzled K-major A tile (conflict-free 16-byte shared loads / stores), fences the async proxy,
//   warp 1     : ONE K = 64 step of tcgen05.mma against the weight tile (loaded once per CTA),
//   warps 2..5 : epilogue (tcgen05.ld, BatchNorm partial sums, bf16 rows) -- then the next tile.
// The per-tile chain is serial inside a CTA (the MMA is 4 instructions); three co-resident CTAs per SM overlap it.
// HBM traffic per output pixel: 8 B of image (L2-shared between neighbouring tiles) + 128 B written, instead of
// 128 B (patches written) + 128 B (patches read) + 128 B.
// ------------------------------------------------------------------------------------------------
constexpr int kEdgeRawStages = 2;
struct EdgeDirectParams {
    int B, Hs, Ws, bh, tiles_y, ipg, raw_bytes, raw_stride;
};

__global__ void __launch_bounds__(192, 3)
edge_down_direct_kernel(const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
                        const __nv_bfloat16* __restrict__ img, float* __restrict__ stats, const EdgeDirectParams p,
                        const int total_tiles) {
    pdl_entry();
    constexpr int BN_ = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                   // 2 x [128][64] bf16, 128B swizzle (double-buffered)
    uint8_t* sB = smem + 2 * kTileM * 128;                // [64][64] bf16, 128B swizzle
    uint8_t* sOut = sB + BN_ * 128;                       // [128][64] bf16 output tile, 128B swizzle, stored by TMA
    uint8_t* sRaw = sOut + kTileM * 128;                  // kEdgeRawStages x raw_stride
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(sRaw + kEdgeRawStages * p.raw_stride);
    uint64_t* raw_empty = raw_full + kEdgeRawStages;
    uint64_t* b_full = raw_empty + kEdgeRawStages;
    uint64_t* a_ready = b_full + 1;                       // [2]
    uint64_t* tfull = a_ready + 2;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);
    float* red = reinterpret_cast<float*>(tmem_slot + 4);  // [4 warps][2][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wp = 2 * p.Ws + 2, Hp = 2 * p.Hs + 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapB);
        prefetch_tmap(&mapOut);
        for (int s = 0; s < kEdgeRawStages; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 128); }
        mbar_init(b_full, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&a_ready[a], 128); mbar_init(&tfull[a], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN_);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(b_full, BN_ * 128);
            tma_load_2d(sB, &mapB, b_full, 0, 0);
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int s = it % kEdgeRawStages;
                mbar_wait(&raw_empty[s], ((it / kEdgeRawStages) & 1) ^ 1);
                const int n = t / p.tiles_y, oy0 = (t - n * p.tiles_y) * p.bh;
                const __nv_bfloat16* src = img + ((size_t)n * Hp + 2 * oy0) * Wp * 4;
                mbar_arrive_expect_tx(&raw_full[s], p.raw_bytes);
                bulk_load_1d(sRaw + s * p.raw_stride, src, p.raw_bytes, &raw_full[s]);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BN_, 0, 0);
        mbar_wait(b_full, 0);
        int lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int buf = lt & 1;
            mbar_wait(&a_ready[buf], (lt >> 1) & 1);
            fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(sA + buf * kTileM * 128), b_addr = smem_u32(sB);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)
                    umma_bf16(tmem_base + buf * BN_, make_sdesc(a_addr + k * 32, 0, 1024), make_sdesc(b_addr + k * 32, 0, 1024),
                              idesc, k > 0 ? 1u : 0u);
                umma_commit(&tfull[buf]);
            }
            __syncwarp();
        }
    } else {
        const int wq = warp & 3;
        const int r = wq * 32 + lane;                      // tile row = output pixel = TMEM lane
        const int ox = r % p.Ws, oyl = r / p.Ws;
        const int e = threadIdx.x - 64;                    // 0..127: the (sum | sum of squares, channel) column this thread flushes
        float acc_stat = 0.f;
        int acc_group = -1;
        // re-pack tile `lt` (the lt-th tile of this CTA) into A buffer lt & 1 and hand it to the MMA warp
        auto repack = [&](int lt) {
            const int s = lt % kEdgeRawStages;
            mbar_wait(&raw_full[s], (lt / kEdgeRawStages) & 1);
            const uint8_t* raw = sRaw + s * p.raw_stride;
            uint4 q[8];
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const uint4* src = reinterpret_cast<const uint4*>(raw + ((size_t)(2 * oyl + ky) * Wp + 2 * ox) * 8);
                q[2 * ky] = src[0];
                q[2 * ky + 1] = src[1];
            }
            mbar_arrive(&raw_empty[s]);                    // raw slot is in registers
            uint8_t* dst = sA + (lt & 1) * kTileM * 128 + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + ((c ^ (r & 7)) << 4)) = q[c];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
            fence_before_sync();                           // this thread's earlier tcgen05.ld of that TMEM buffer is done
            mbar_arrive(&a_ready[lt & 1]);
        };
        int lt = 0;
        if (blockIdx.x < total_tiles) repack(0);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            // tile lt+1 is re-packed (and its MMA runs) while tile lt's accumulator is drained below; its A / TMEM
            // buffer (lt+1)&1 was last used by tile lt-1, whose epilogue this thread finished in the previous iteration
            if (t + (int)gridDim.x < total_tiles) repack(lt + 1);
            const int n = t / p.tiles_y, oy0 = (t - n * p.tiles_y) * p.bh;
            const int buf = lt & 1;
            mbar_wait(&tfull[buf], (lt >> 1) & 1);
            fence_after_sync();
            const uint32_t tmem_d = tmem_base + buf * BN_ + ((uint32_t)(wq * 32) << 16);
            // the output tile = 128 consecutive pixels x 64 channels = 16 KB contiguous in HBM: stage it (swizzled,
            // conflict-free 16-byte shared stores) and let ONE TMA store write full lines, instead of 128 threads
            // each dribbling 16-byte pieces of their own row
            if (e == 0) tma_store_wait_read0();                // previous tile's store has finished reading sOut
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN_ / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_d + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    uint4 u;
                    u.x = pack_bf16x2(v[qq * 8 + 0], v[qq * 8 + 1]);
                    u.y = pack_bf16x2(v[qq * 8 + 2], v[qq * 8 + 3]);
                    u.z = pack_bf16x2(v[qq * 8 + 4], v[qq * 8 + 5]);
                    u.w = pack_bf16x2(v[qq * 8 + 6], v[qq * 8 + 7]);
                    *reinterpret_cast<uint4*>(sOut + r * 128 + (((c * 4 + qq) ^ (r & 7)) << 4)) = u;
                }
                if (stats != nullptr) {
                    float sq[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
                    const float s1 = warp_transpose_sum(v, lane);
                    const float s2 = warp_transpose_sum(sq, lane);
                    red[(wq * 2 + 0) * BN_ + c * 32 + lane] = s1;
                    red[(wq * 2 + 1) * BN_ + c * 32 + lane] = s2;
                }
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e == 0) {
                tma_store_2d(&mapOut, sOut, 0, (n * p.Hs + oy0) * p.Ws);
                tma_store_commit();
            }
            if (stats != nullptr) {
                // per-CTA running sums, flushed with ONE atomic per column when the BatchNorm group changes / at the
                // end (a CTA walks its tiles in image order), instead of 128 same-address atomics per tile;
                // red[] is safe until the next tile's first bar.sync
                const int which = e / BN_, cc = e % BN_, g = n / p.ipg;
                if (g != acc_group) {
                    if (acc_group >= 0) atomicAdd(stats + (size_t)acc_group * 2 * BN_ + which * BN_ + cc, acc_stat);
                    acc_group = g;
                    acc_stat = 0.f;
                }
                acc_stat += red[(0 * 2 + which) * BN_ + cc] + red[(1 * 2 + which) * BN_ + cc] +
                            red[(2 * 2 + which) * BN_ + cc] + red[(3 * 2 + which) * BN_ + cc];
            }
        }
        if (e == 0) tma_store_wait_all();
        if (stats != nullptr && acc_group >= 0)
            atomicAdd(stats + (size_t)acc_group * 2 * BN_ + (e / BN_) * BN_ + (e % BN_), acc_stat);
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 2 * BN_);
    }
}

struct EdgePlan { bool ok; int splits, steps_per_split, total_steps; };
EdgePlan edge_wgrad_plan(int B, int Hs, int Ws) {
    EdgePlan pl{};
    pl.ok = ((long long)B * Hs * Ws) % kWgradKPix == 0;
    if (!pl.ok) return pl;
    pl.total_steps = (int)((long long)B * Hs * Ws / kWgradKPix);
    int splits = kNumSMs < pl.total_steps ? kNumSMs : pl.total_steps;
    pl.steps_per_split = (pl.total_steps + splits - 1) / splits;
    pl.splits = (pl.total_steps + pl.steps_per_split - 1) / pl.steps_per_split;
    return pl;
}


// ------------------------------------------------------------------------------------------------
// Image-edge up conv, scatter form (G.conv5 forward, D.conv1 input gradient; Ca = 64 -> nc <= 4 image channels):
//     img[2i - 1 + ky][2j - 1 + kx][c] += x[i][j][:] . w[:][c][ky][kx]
// The 9-shift gather form above re-reads every activation tile nine times (its bound is the L2 -> shared-memory
// stream, 3.4x the HBM time).  Here each tile of 4 input rows x 32 pixels is loaded ONCE and multiplied by the whole
// filter bank in one K = 64, N = 64 = (ky, kx, c4) MMA -- B is the edge-down weight matrix w_down_e[ca][(ky,kx,c4)] read
// as an MN-major operand, so no second packing exists -- and the 16 tap images are folded in the epilogue: columns
// across neighbouring lanes by shuffle (a warp owns one input row), rows across warps through 8 KB of shared memory.
// A tile of input rows 3t-1 .. 3t+2 (rows outside the image arrive as TMA zero fill) completes output rows
// 6t-1 .. 6t+4, so tiles are independent (11 per 32-row image: 1.375x the activation bytes instead of 9x) and every
// output pixel is written exactly once, as 8-byte (c0 c1 c2 0) records of the JCK_IMG_P4 layout.
// ------------------------------------------------------------------------------------------------
constexpr int kEUStages = 4;
constexpr int kEUABytes = kTileM * kBK * 2;                  // 16 KB: 4 rows x 32 pixels x 64 channels
constexpr int kEUBOff = kEUStages * kEUABytes;               // weights, 8 KB
constexpr int kEUXOff = kEUBOff + 64 * 128;                  // 2 x [128 threads][4 float4] row-exchange buffers
constexpr int kEUBarOff = kEUXOff + 2 * 128 * 64;
constexpr int kEUSmem = kEUBarOff + 256 + 1024;

__global__ void __launch_bounds__(192, 2)
edge_up_scatter_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                       __nv_bfloat16* __restrict__ img, const int B, const int Hs, const int tiles_img, const int total_tiles) {
    pdl_trigger();
    constexpr int Ws = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kEUBarOff);
    uint64_t* empty = full + kEUStages;
    uint64_t* tfull = empty + kEUStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < kEUStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(wfull, 64 * 128);
            tma_load_2d(smem + kEUBOff, &mapB, wfull, 0, 0);
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int s = it % kEUStages;
                mbar_wait(&empty[s], ((it / kEUStages) & 1) ^ 1);
                mbar_arrive_expect_tx(&full[s], kEUABytes);
                const int n = t / tiles_img, y0 = 3 * (t % tiles_img) - 1;
                tma_load_4d(smem + s * kEUABytes, &mapA, &full[s], 0, 0, y0, n);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(64, 0, 1);
        mbar_wait(wfull, 0);
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int s = it % kEUStages, acc = it & 1;
            mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
            mbar_wait(&full[s], (it / kEUStages) & 1);
            fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(smem + s * kEUABytes), b_addr = smem_u32(smem + kEUBOff);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)
                    umma_bf16(tmem_base + acc * 64, make_sdesc(a_addr + k * 32, 0, 1024), make_sdesc(b_addr + k * 2048, kBK * 128, 1024),
                              idesc, k > 0 ? 1u : 0u);
                umma_commit(&empty[s]);
                umma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else {
        const int r = warp & 3;                    // local input row = TMEM lane quarter; lane = input column j
        const int e = r * 32 + lane;
        const int Hp = 2 * Hs + 2, Wp = 2 * Ws + 2;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const int n = t / tiles_img, y0 = 3 * (t % tiles_img) - 1;
            mbar_wait(&tfull[acc], (it >> 1) & 1);
            fence_after_sync();
            float v[64];                           // v[(ky*4 + kx)*4 + c]
            {
                float lo[32], hi[32];
                const uint32_t ta = tmem_base + acc * 64 + ((uint32_t)(r * 32) << 16);
                tmem_ld32(ta, lo);
                tmem_ld32(ta + 32, hi);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { v[i] = lo[i]; v[32 + i] = hi[i]; }
            }
            fence_before_sync();
            mbar_arrive(&tempty[acc]);
            // columns: X = 2j (b = 0) takes kx = 1 of column j and kx = 3 of column j-1; X = 2j+1 takes kx = 2 of j, kx = 0 of j+1
            float R[4][2][4];
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float left = __shfl_up_sync(0xffffffffu, v[(ky * 4 + 3) * 4 + c], 1);
                    float right = __shfl_down_sync(0xffffffffu, v[(ky * 4 + 0) * 4 + c], 1);
                    if (lane == 0) left = 0.f;
                    if (lane == 31) right = 0.f;
                    R[ky][0][c] = v[(ky * 4 + 1) * 4 + c] + left;
                    R[ky][1][c] = v[(ky * 4 + 2) * 4 + c] + right;
                }
            }
            // rows: Y = 2i (a = 0) takes ky = 1 of row i and ky = 3 of row i-1; Y = 2i+1 takes ky = 2 of i and ky = 0 of i+1
            float4* xb = reinterpret_cast<float4*>(smem + kEUXOff + (it & 1) * (128 * 64));
            xb[e * 4 + 0] = make_float4(R[3][0][0], R[3][0][1], R[3][0][2], R[3][0][3]);
            xb[e * 4 + 1] = make_float4(R[3][1][0], R[3][1][1], R[3][1][2], R[3][1][3]);
            xb[e * 4 + 2] = make_float4(R[0][0][0], R[0][0][1], R[0][0][2], R[0][0][3]);
            xb[e * 4 + 3] = make_float4(R[0][1][0], R[0][1][1], R[0][1][2], R[0][1][3]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int i = y0 + r;
            if (r >= 1) {                          // a = 0: output row 2i
                const int Y = 2 * i;
                if (Y >= 0 && Y < 2 * Hs) {
                    const float4 u0 = xb[(e - 32) * 4 + 0], u1 = xb[(e - 32) * 4 + 1];
                    uint2 p0, p1;
                    p0.x = pack_bf16x2(R[1][0][0] + u0.x, R[1][0][1] + u0.y);
                    p0.y = pack_bf16x2(R[1][0][2] + u0.z, R[1][0][3] + u0.w);
                    p1.x = pack_bf16x2(R[1][1][0] + u1.x, R[1][1][1] + u1.y);
                    p1.y = pack_bf16x2(R[1][1][2] + u1.z, R[1][1][3] + u1.w);
                    __nv_bfloat16* o = img + (((size_t)n * Hp + Y + 1) * Wp + 2 * lane + 1) * 4;
                    *reinterpret_cast<uint2*>(o) = p0;
                    *reinterpret_cast<uint2*>(o + 4) = p1;
                }
            }
            if (r <= 2) {                          // a = 1: output row 2i + 1
                const int Y = 2 * i + 1;
                if (Y >= 0 && Y < 2 * Hs) {
                    const float4 u0 = xb[(e + 32) * 4 + 2], u1 = xb[(e + 32) * 4 + 3];
                    uint2 p0, p1;
                    p0.x = pack_bf16x2(R[2][0][0] + u0.x, R[2][0][1] + u0.y);
                    p0.y = pack_bf16x2(R[2][0][2] + u0.z, R[2][0][3] + u0.w);
                    p1.x = pack_bf16x2(R[2][1][0] + u1.x, R[2][1][1] + u1.y);
                    p1.y = pack_bf16x2(R[2][1][2] + u1.z, R[2][1][3] + u1.w);
                    __nv_bfloat16* o = img + (((size_t)n * Hp + Y + 1) * Wp + 2 * lane + 1) * 4;
                    *reinterpret_cast<uint2*>(o) = p0;
                    *reinterpret_cast<uint2*>(o + 4) = p1;
                }
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 128);
    }
}

bool want_tc(int dtype, int algo) { return dtype == JCK_BF16 && algo != JCK_ALGO_SIMT; }

// ------------------------------------------------------------------------------------------------
// Dense GEMM on tcgen05 (the CGAN discriminator head, model/CGAN.py:118-120 Linear(16*C4 + E -> 256), forward,
// input gradient, weight gradient and their second-order twins):   C[m][n] (+)= sum_k A(m,k) * B(n,k),  bf16 -> fp32.
// Either operand may be K-major (A[m*lda + k]: one TMA box of 128 rows x 64 k, the conv kernels' operand form) or
// MN-major (A[k*lda + m]: two boxes of 64 k-rows x 64 m, the weight-gradient form), so x.W^T, g.W and g^T.x all run
// without a transposed copy.  One CTA = one 128 x 128 tile over a contiguous range of 64-wide K steps (split-K over
// blockIdx.x, so the 8192-long contraction of the forward product fills the chip); 4-stage TMA ring, accumulator
// in TMEM, thread-per-row epilogue straight to C (fp32, fp32 +=, or bf16) or to a per-split fp32 partial that
// gemm_reduce_kernel sums.  Ragged M / N / K are zero-filled by TMA on the way in and masked on the way out.
// ------------------------------------------------------------------------------------------------
struct GemmParams {
    int M, N, K;
    long long ldc;             // row pitch of C (elements)
    int ksteps, steps_per_split;
    int n_tiles;
    int mode;                  // 0: C fp32 =, 1: C fp32 +=, 2: C bf16 =, 3: fp32 partial [split][M][N]
    float* stats;              // optional (single split only): stats[n % sC] += sum_m C, stats[sC + n % sC] += sum_m C^2
    int stats_C;
    int stages;                // depth of the TMA ring: 4 for long contractions, 2 for short ones (3 CTAs per SM instead of 1)
};
constexpr int kGemmMaxStages = 4;
constexpr int kGemmStage = 2 * kTileM * kBK * 2;   // A 16 KB + B 16 KB
constexpr int gemm_smem(int stages) { return 1024 + stages * kGemmStage + 1024; }   // barriers | ring | alignment slack

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(kConvThreads)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void* __restrict__ Cout,
               const GemmParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kGemmMaxStages;
    uint64_t* tmem_full = empty + kGemmMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    uint8_t* ring = smem + 1024;
    const int stages = p.stages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int m0 = (blockIdx.y / p.n_tiles) * kTileM, n0 = (blockIdx.y % p.n_tiles) * 128;
    const int step_beg = split * p.steps_per_split;
    const int nsteps = max(0, min(p.ksteps, step_beg + p.steps_per_split) - step_beg);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();     // everything above is on-chip set-up; global memory is first touched below

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % stages;
                mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
                uint8_t* sa = ring + s * kGemmStage;
                uint8_t* sb = sa + kTileM * kBK * 2;
                mbar_arrive_expect_tx(&full[s], kGemmStage);
                const int k0 = (step_beg + it) * kBK;
                if (A_MN) {
                    tma_load_2d(sa, &mapA, &full[s], m0, k0);
                    tma_load_2d(sa + kBK * 128, &mapA, &full[s], m0 + 64, k0);
                } else {
                    tma_load_2d(sa, &mapA, &full[s], k0, m0);
                }
                if (B_MN) {
                    tma_load_2d(sb, &mapB, &full[s], n0, k0);
                    tma_load_2d(sb + kBK * 128, &mapB, &full[s], n0 + 64, k0);
                } else {
                    tma_load_2d(sb, &mapB, &full[s], k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(128, A_MN, B_MN);
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % stages;
            mbar_wait(&full[s], (it / stages) & 1);
            fence_after_sync();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(ring + s * kGemmStage);
                const uint32_t b_addr = a_addr + kTileM * kBK * 2;
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t da = A_MN ? make_sdesc(a_addr + k * 2048, kBK * 128, 1024) : make_sdesc(a_addr + k * 32, 0, 1024);
                    const uint64_t db = B_MN ? make_sdesc(b_addr + k * 2048, kBK * 128, 1024) : make_sdesc(b_addr + k * 32, 0, 1024);
                    umma_bf16(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty[s]);
                if (it == nsteps - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        const int wq = warp & 3;
        const int m = m0 + wq * 32 + lane;
        if (nsteps > 0) {
            mbar_wait(tmem_full, 0);
            fence_after_sync();
        }
        float* cf = reinterpret_cast<float*>(Cout);
        __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(Cout);
        long long row;
        if (p.mode == 3) row = ((long long)split * p.M + m) * p.N;
        else row = (long long)m * p.ldc;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            float v[32];
            if (nsteps > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + c * 32, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            const int nb = n0 + c * 32;
            if (p.stats != nullptr) {
                // BatchNorm statistics of the product (G.conv1: channel = n % C) from the fp32 accumulators; rows past M
                // are zero (TMA zero-fill), so they add nothing
                float s1v[32], s2v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { s1v[i] = v[i]; s2v[i] = v[i] * v[i]; }
                const float s1 = warp_transpose_sum(s1v, lane);
                const float s2 = warp_transpose_sum(s2v, lane);
                if (nb + lane < p.N) {
                    const int ch = (nb + lane) % p.stats_C;
                    atomicAdd(p.stats + ch, s1);
                    atomicAdd(p.stats + p.stats_C + ch, s2);
                }
            }
            if (m >= p.M || nb >= p.N) continue;
            const bool whole = nb + 32 <= p.N;
            if (p.mode == 2) {
                if (whole && (((row + nb) & 7) == 0)) {
                    uint4* dst = reinterpret_cast<uint4*>(cb + row + nb);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                        u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                        u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                        u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                        dst[q] = u;
                    }
                } else {
                    for (int i = 0; i < 32 && nb + i < p.N; ++i) cb[row + nb + i] = __float2bfloat16_rn(v[i]);
                }
            } else if (whole && (((row + nb) & 3) == 0)) {
                float4* dst = reinterpret_cast<float4*>(cf + row + nb);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 o = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    if (p.mode == 1) {
                        const float4 old = dst[q];
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    dst[q] = o;
                }
            } else {
                for (int i = 0; i < 32 && nb + i < p.N; ++i) {
                    if (p.mode == 1) cf[row + nb + i] += v[i];
                    else cf[row + nb + i] = v[i];
                }
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 128);
    }
}

// C[m][n] (=, +=) sum over splits of part[s][m][n]; mode as GemmParams (0 fp32 =, 1 fp32 +=, 2 bf16 =)
__global__ void gemm_reduce_kernel(const float* __restrict__ part, void* __restrict__ Cout, int M, int N, long long ldc,
                                   int splits, int mode) {
    pdl_entry();
    const long long total = (long long)M * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += part[(long long)s * total + i];
        const long long o = (i / N) * ldc + (i % N);
        if (mode == 2) reinterpret_cast<__nv_bfloat16*>(Cout)[o] = __float2bfloat16_rn(acc);
        else if (mode == 1) reinterpret_cast<float*>(Cout)[o] += acc;
        else reinterpret_cast<float*>(Cout)[o] = acc;
    }
}

// operand maps: K-major [rows][ld] -> dims (K | rows), box (64 | 128);  MN-major [K][ld] -> dims (rows | K), box (64 | 64)
int map_gemm_operand(CUtensorMap* m, const void* p, int mn_major, int rows, int K, long long ld) {
    cuuint64_t dims[2];
    cuuint32_t box[2];
    if (mn_major) { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; box[0] = 64; box[1] = 64; }
    else { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; box[0] = 64; box[1] = 128; }
    cuuint64_t str[1] = {(cuuint64_t)ld * 2};
    return encode(m, p, 2, dims, str, box);
}

struct GemmPlan { int m_tiles, n_tiles, ksteps, splits, steps_per_split; };
GemmPlan gemm_plan(int M, int N, int K) {
    GemmPlan pl;
    pl.m_tiles = (M + kTileM - 1) / kTileM;
    pl.n_tiles = (N + 127) / 128;
    pl.ksteps = (K + kBK - 1) / kBK;
    const int tiles = pl.m_tiles * pl.n_tiles;
    int splits = tiles >= kNumSMs ? 1 : kNumSMs / tiles;
    const int max_splits = pl.ksteps / 4 > 0 ? pl.ksteps / 4 : 1;     // at least 4 K steps per CTA
    if (splits > max_splits) splits = max_splits;
    pl.steps_per_split = (pl.ksteps + splits - 1) / splits;
    pl.splits = (pl.ksteps + pl.steps_per_split - 1) / pl.steps_per_split;
    return pl;
}

template <int A_MN, int B_MN>
int launch_gemm(const CUtensorMap& mA, const CUtensorMap& mB, void* C, const GemmParams& p, dim3 grid, cudaStream_t st) {
    static DeviceOnce cfg;
    if (!cfg.done()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             gemm_smem(kGemmMaxStages));
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "gemm_tc smem attr: %s", cudaGetErrorString(e));
        cfg.mark();
    }
    launch_pdl(gemm_tc_kernel<A_MN, B_MN>, dim3(grid), dim3(kConvThreads), gemm_smem(p.stages), st, mA, mB, C, p);
    JCK_LAUNCH_CHECK("gemm_tc");
    return JCK_OK;
}

}  // namespace

// 2-D bf16 tensor map (dims / box innermost first, 128-byte swizzle) for the other translation units (incep.cu)
int encode_bf16_2d(CUtensorMap* m, const void* base, unsigned long long inner, unsigned long long rows,
                   unsigned long long pitch_bytes, unsigned box_inner, unsigned box_rows, int swizzle64) {
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t str[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    return encode(m, base, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}
}  // namespace jck

using namespace jck;

extern "C" int jck_conv_down(const void* in_large, const void* w_down, void* out_small, float* stats, int B, int Hs,
                             int Ws, int Ca, int Cb, int imgs_per_group, int dtype, int algo, void* stream) {
    JCK_REQUIRE(in_large && w_down && out_small && B > 0 && Hs > 0 && Ws > 0 && Ca > 0 && Cb > 0, "conv_down: bad argument");
    if (imgs_per_group <= 0) imgs_per_group = B;
    cudaStream_t st = as_stream(stream);
    if (want_tc(dtype, algo)) {
        PatchGeom g;
        if (tc_conv_supported(B, Hs, Ws, Ca, Cb, imgs_per_group, false, &g))
            return conv_tc<false>(in_large, w_down, out_small, stats, B, Hs, Ws, Ca, Cb, imgs_per_group, st);
        if (algo == JCK_ALGO_TC)
            return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_down: no tcgen05 tile for Ca=%d Cb=%d Hs=%d Ws=%d", Ca, Cb, Hs, Ws);
    } else if (algo == JCK_ALGO_TC) {
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_down: tcgen05 path is bf16 only");
    }
    if (dtype == JCK_F32) return simt_down<float>(in_large, w_down, out_small, stats, B, Hs, Ws, Ca, Cb, imgs_per_group, st);
    if (dtype == JCK_BF16) return simt_down<__nv_bfloat16>(in_large, w_down, out_small, stats, B, Hs, Ws, Ca, Cb, imgs_per_group, st);
    return set_error(JCK_E_BADARG, "conv_down: dtype %d", dtype);
}

extern "C" int jck_conv_up(const void* in_small, const void* w_up, void* out_large, float* stats, int B, int Hs, int Ws,
                           int Ca, int Cb, int imgs_per_group, int dtype, int algo, void* stream) {
    JCK_REQUIRE(in_small && w_up && out_large && B > 0 && Hs > 0 && Ws > 0 && Ca > 0 && Cb > 0, "conv_up: bad argument");
    if (imgs_per_group <= 0) imgs_per_group = B;
    cudaStream_t st = as_stream(stream);
    if (want_tc(dtype, algo)) {
        PatchGeom g;
        if (tc_conv_supported(B, Hs, Ws, Ca, Cb, imgs_per_group, true, &g))
            return conv_tc<true>(in_small, w_up, out_large, stats, B, Hs, Ws, Ca, Cb, imgs_per_group, st);
        if (algo == JCK_ALGO_TC)
            return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_up: no tcgen05 tile for Ca=%d Cb=%d Hs=%d Ws=%d", Ca, Cb, Hs, Ws);
    } else if (algo == JCK_ALGO_TC) {
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_up: tcgen05 path is bf16 only");
    }
    if (dtype == JCK_F32) return simt_up<float>(in_small, w_up, out_large, stats, B, Hs, Ws, Ca, Cb, imgs_per_group, st);
    if (dtype == JCK_BF16) return simt_up<__nv_bfloat16>(in_small, w_up, out_large, stats, B, Hs, Ws, Ca, Cb, imgs_per_group, st);
    return set_error(JCK_E_BADARG, "conv_up: dtype %d", dtype);
}

// Input-gradient convolutions with the BatchNorm-backward reduction of the layer BELOW fused into the epilogue
// (tcgen05 / bf16 only; the fp32 parity mode keeps the separate jck_bn_act_bwd_reduce pass).
extern "C" int jck_conv_up_bnbwd(const void* in_small, const void* w_up, const void* y_saved, const float* scale_shift,
                                 const float* mean_rstd, float slope, void* out_g, float* sums, int B, int Hs, int Ws,
                                 int Ca, int Cb, int imgs_per_group, int dtype, void* stream) {
    JCK_REQUIRE(in_small && w_up && y_saved && scale_shift && mean_rstd && out_g && sums && B > 0, "conv_up_bnbwd: bad argument");
    if (imgs_per_group <= 0) imgs_per_group = B;
    PatchGeom g;
    if (dtype != JCK_BF16 || !tc_conv_supported(B, Hs, Ws, Ca, Cb, imgs_per_group, true, &g))
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_up_bnbwd: bf16 tcgen05 shapes only (Ca=%d Cb=%d Hs=%d)", Ca, Cb, Hs);
    const BnBwdEpi bb{(const __nv_bfloat16*)y_saved, scale_shift, mean_rstd, slope};
    return conv_tc<true>(in_small, w_up, out_g, sums, B, Hs, Ws, Ca, Cb, imgs_per_group, as_stream(stream), &bb);
}

extern "C" int jck_conv_down_bnbwd(const void* in_large, const void* w_down, const void* y_saved, const float* scale_shift,
                                   const float* mean_rstd, float slope, void* out_g, float* sums, int B, int Hs, int Ws,
                                   int Ca, int Cb, int imgs_per_group, int dtype, void* stream) {
    JCK_REQUIRE(in_large && w_down && y_saved && scale_shift && mean_rstd && out_g && sums && B > 0, "conv_down_bnbwd: bad argument");
    if (imgs_per_group <= 0) imgs_per_group = B;
    PatchGeom g;
    if (dtype != JCK_BF16 || !tc_conv_supported(B, Hs, Ws, Ca, Cb, imgs_per_group, false, &g))
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_down_bnbwd: bf16 tcgen05 shapes only (Ca=%d Cb=%d Hs=%d)", Ca, Cb, Hs);
    const BnBwdEpi bb{(const __nv_bfloat16*)y_saved, scale_shift, mean_rstd, slope};
    return conv_tc<false>(in_large, w_down, out_g, sums, B, Hs, Ws, Ca, Cb, imgs_per_group, as_stream(stream), &bb);
}

static bool wgrad_uses_tc(int B, int Hs, int Ws, int Ca, int Cb, int dtype, int algo, WgradPlan* pl) {
    if (!want_tc(dtype, algo)) return false;
    *pl = wgrad_plan(B, Hs, Ws, Ca, Cb);
    return pl->ok;
}

extern "C" size_t jck_conv_wgrad_workspace_bytes(int B, int Hs, int Ws, int Ca, int Cb, int dtype, int algo) {
    WgradPlan pl;
    int splits = wgrad_uses_tc(B, Hs, Ws, Ca, Cb, dtype, algo, &pl) ? 1 : simt_wgrad_splits(B, Hs, Ws, Ca, Cb);
    return (size_t)splits * Ca * 16 * Cb * sizeof(float);
}

extern "C" int jck_conv_wgrad(const void* small, const void* large, float* dw4, void* workspace, size_t workspace_bytes,
                              int B, int Hs, int Ws, int Ca, int Cb, int accumulate, int dtype, int algo, void* stream) {
    JCK_REQUIRE(small && large && dw4 && workspace && B > 0 && Hs > 0 && Ws > 0 && Ca > 0 && Cb > 0, "conv_wgrad: bad argument");
    cudaStream_t st = as_stream(stream);
    const size_t need = jck_conv_wgrad_workspace_bytes(B, Hs, Ws, Ca, Cb, dtype, algo);
    JCK_REQUIRE(workspace_bytes >= need, "conv_wgrad: workspace %zu < %zu bytes", workspace_bytes, need);
    WgradPlan pl;
    int rc, splits;
    if (wgrad_uses_tc(B, Hs, Ws, Ca, Cb, dtype, algo, &pl)) {
        splits = 1;             // the tcgen05 kernel's splits have already been summed (TMA reducing stores)
        rc = wgrad_tc(small, large, (float*)workspace, pl, B, Hs, Ws, Ca, Cb, st);
    } else {
        if (algo == JCK_ALGO_TC) return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_wgrad: no tcgen05 path for this shape/dtype");
        splits = simt_wgrad_splits(B, Hs, Ws, Ca, Cb);
        if (dtype == JCK_F32) rc = simt_wgrad<float>(small, large, (float*)workspace, splits, B, Hs, Ws, Ca, Cb, st);
        else if (dtype == JCK_BF16) rc = simt_wgrad<__nv_bfloat16>(small, large, (float*)workspace, splits, B, Hs, Ws, Ca, Cb, st);
        else return set_error(JCK_E_BADARG, "conv_wgrad: dtype %d", dtype);
    }
    if (rc) return rc;
    return launch_wgrad_unpack((const float*)workspace, dw4, Ca, Cb, splits, accumulate, st);
}

// ------------------------------------------------------------------------------------------------
// image-edge entry points (bf16, JCK_IMG_P4 image layout)
// ------------------------------------------------------------------------------------------------
extern "C" int jck_edge_down_img(const void* img_p4, const void* w_down_e, void* out_small, float* stats, int B, int Hs, int Ws,
                                 int Ca, int imgs_per_group, void* stream) {
    JCK_REQUIRE(img_p4 && w_down_e && out_small && B > 0 && Hs > 0 && Ws > 0, "edge_down_img: bad argument");
    if (imgs_per_group <= 0) imgs_per_group = B;
    PatchGeom g;
    if (Ca != 64 || !patch_geom(Hs, Ws, kTileM, &g) || g.bw != Ws || g.nb != 1)
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "edge_down_img: Ca=%d Hs=%d Ws=%d (needs Ca = 64, Ws <= 128, Hs*Ws >= 128)", Ca, Hs, Ws);
    EdgeDirectParams p{B, Hs, Ws, g.bh, Hs / g.bh, imgs_per_group, (2 * g.bh + 2) * (2 * Ws + 2) * 8, 0};
    p.raw_stride = (p.raw_bytes + 127) & ~127;
    const int smem = 3 * kTileM * 128 + 64 * 128 + kEdgeRawStages * p.raw_stride + 256 + 4 * 2 * 64 * 4 + 1024;
    JCK_REQUIRE(smem <= 72 * 1024, "edge_down_img: tile too large for the raw ring (%d bytes)", smem);
    CUtensorMap mB, mOut;
    int rc;
    if ((rc = map_matrix(&mB, w_down_e, Ca, 64, 64))) return rc;
    if ((rc = map_rows64(&mOut, out_small, (long long)B * Hs * Ws, kTileM))) return rc;
    static DeviceOnce cfg;
    if (!cfg.done()) {
        cudaError_t e = cudaFuncSetAttribute(edge_down_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "edge_down_img smem attr: %s", cudaGetErrorString(e));
        cfg.mark();
    }
    const int total = B * p.tiles_y;
    const int grid = total < 3 * kNumSMs ? total : 3 * kNumSMs;
    launch_pdl(edge_down_direct_kernel, dim3(grid), dim3(192), smem, as_stream(stream), mB, mOut, (const __nv_bfloat16*)img_p4, stats, p, total);
    JCK_LAUNCH_CHECK("edge_down_img");
    return JCK_OK;
}

extern "C" int jck_edge_up(const void* in_small, const void* w_up9, void* img_p4, int B, int Hs, int Ws, int Ca, void* stream) {
    JCK_REQUIRE(in_small && w_up9 && img_p4 && B > 0 && Hs > 0 && Ws > 0, "edge_up: bad argument");
    PatchGeom g;
    if (Ca % 64 != 0 || !patch_geom(Hs, Ws, kTileM, &g))
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "edge_up: Ca=%d Hs=%d Ws=%d", Ca, Hs, Ws);
    ConvTcParams p{B, Hs, Ws, Ca, 4, g.bw, g.bh, g.nb, Ws / g.bw, Hs / g.bh, B};
    const int m_tiles = p.tiles_x * p.tiles_y * ((B + g.nb - 1) / g.nb);
    CUtensorMap mA, mB;
    int rc;
    if ((rc = map_small(&mA, in_small, Ca, Ws, Hs, B, g.bw, g.bh, g.nb))) return rc;
    if ((rc = map_matrix(&mB, w_up9, 16, 9 * Ca, 16))) return rc;
    return launch_conv_tc_persist<16, 4, kEdgeUp>(mA, mB, img_p4, nullptr, p, m_tiles, 1, as_stream(stream));
}

extern "C" size_t jck_edge_wgrad_workspace_bytes(int B, int Hs, int Ws, int Ca) {
    EdgePlan pl = edge_wgrad_plan(B, Hs, Ws);
    return pl.ok ? (size_t)pl.splits * 64 * 64 * sizeof(float) : 0;
}

extern "C" int jck_edge_wgrad_img(const void* small, const void* img_p4, float* dw4, void* workspace, size_t workspace_bytes,
                                  int B, int Hs, int Ws, int Ca, int nc, int accumulate, void* stream) {
    JCK_REQUIRE(small && img_p4 && dw4 && workspace && B > 0 && nc > 0 && nc <= 4, "edge_wgrad_img: bad argument");
    EdgePlan pl = edge_wgrad_plan(B, Hs, Ws);
    if (Ca != 64 || !pl.ok || Ws > kWgradKPix || kWgradKPix % Ws != 0 || Hs % (kWgradKPix / Ws) != 0)
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "edge_wgrad_img: Ca=%d Hs=%d Ws=%d", Ca, Hs, Ws);
    JCK_REQUIRE(workspace_bytes >= (size_t)pl.splits * 64 * 64 * sizeof(float), "edge_wgrad_img: workspace too small");
    cudaStream_t st = as_stream(stream);
    CUtensorMap mS;
    int rc;
    if ((rc = map_rows64(&mS, small, (long long)B * Hs * Ws, kWgradKPix))) return rc;
    const int rps = kWgradKPix / Ws;
    EdgeWgradDirectParams p{pl.total_steps, pl.steps_per_split, Hs, Ws, rps, (2 * rps + 2) * (2 * Ws + 2) * 8, 0};
    p.raw_stride = (p.raw_bytes + 127) & ~127;
    const int smem = (2 * kEdgeWDStages + 1) * kWgradKPix * 128 + kEdgeWDStages * p.raw_stride + 256 + 1024;
    JCK_REQUIRE(smem <= 100 * 1024, "edge_wgrad_img: raw ring too large (%d bytes)", smem);
    static DeviceOnce cfg;
    if (!cfg.done()) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_edge_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "edge_wgrad_img smem attr: %s", cudaGetErrorString(e));
        cfg.mark();
    }
    launch_pdl(wgrad_edge_direct_kernel, dim3(pl.splits), dim3(kConvThreads), smem, st, mS, (const __nv_bfloat16*)img_p4, (float*)workspace, p);
    JCK_LAUNCH_CHECK("edge_wgrad_img");
    launch_pdl(edge_wgrad_unpack_kernel, dim3((64 * nc * 16 + 255) / 256), dim3(256), 0, st, (const float*)workspace, dw4, nc, pl.splits, accumulate);
    JCK_LAUNCH_CHECK("edge_wgrad_unpack");
    return JCK_OK;
}

extern "C" size_t jck_gemm_tc_workspace_bytes(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const GemmPlan pl = gemm_plan(M, N, K);
    return pl.splits > 1 ? (size_t)pl.splits * M * N * sizeof(float) : 0;
}

extern "C" int jck_gemm_tc(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, void* C,
                           int c_dtype, long long ldc, int M, int N, int K, int accumulate, float* stats, int stats_channels,
                           void* workspace, size_t workspace_bytes, void* stream) {
    JCK_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gemm_tc: bad argument");
    JCK_REQUIRE(c_dtype == JCK_F32 || (c_dtype == JCK_BF16 && !accumulate), "gemm_tc: C must be fp32, or bf16 without accumulate");
    if (lda % 8 != 0 || ldb % 8 != 0 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15))
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "gemm_tc: operand pitch / base not 16-byte aligned (lda=%lld ldb=%lld)", lda, ldb);
    cudaStream_t st = as_stream(stream);
    const GemmPlan pl = gemm_plan(M, N, K);
    const int final_mode = c_dtype == JCK_BF16 ? 2 : (accumulate ? 1 : 0);
    JCK_REQUIRE(pl.splits == 1 || (workspace && workspace_bytes >= (size_t)pl.splits * M * N * sizeof(float)),
                "gemm_tc: workspace too small");
    CUtensorMap mA, mB;
    int rc;
    if ((rc = map_gemm_operand(&mA, A, a_mn_major, M, K, lda))) return rc;
    if ((rc = map_gemm_operand(&mB, B, b_mn_major, N, K, ldb))) return rc;
    JCK_REQUIRE(!stats || (pl.splits == 1 && stats_channels > 0), "gemm_tc: statistics need a single-split plan");
    GemmParams p{M, N, K, ldc, pl.ksteps, pl.steps_per_split, pl.n_tiles, pl.splits > 1 ? 3 : final_mode, stats, stats_channels,
                 pl.steps_per_split >= kGemmMaxStages ? kGemmMaxStages : 2};
    void* dst = pl.splits > 1 ? workspace : C;
    dim3 grid(pl.splits, pl.m_tiles * pl.n_tiles);
    if (!a_mn_major && !b_mn_major) rc = launch_gemm<0, 0>(mA, mB, dst, p, grid, st);
    else if (!a_mn_major && b_mn_major) rc = launch_gemm<0, 1>(mA, mB, dst, p, grid, st);
    else if (a_mn_major && !b_mn_major) rc = launch_gemm<1, 0>(mA, mB, dst, p, grid, st);
    else rc = launch_gemm<1, 1>(mA, mB, dst, p, grid, st);
    if (rc) return rc;
    if (pl.splits > 1) {
        const long long total = (long long)M * N;
        long long blocks = (total + 255) / 256;
        if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
        launch_pdl(gemm_reduce_kernel, dim3((int)blocks), dim3(256), 0, st, (const float*)workspace, C, M, N, ldc, pl.splits, final_mode);
        JCK_LAUNCH_CHECK("gemm_reduce");
    }
    return JCK_OK;
}

extern "C" int jck_edge_up_scatter(const void* in_small, const void* w_down_e, void* img_p4, int B, int Hs, int Ws, int Ca,
                                   void* stream) {
    JCK_REQUIRE(in_small && w_down_e && img_p4 && B > 0 && Hs > 0, "edge_up_scatter: bad argument");
    if (Ca != 64 || Ws != 32)
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "edge_up_scatter: Ca=%d Ws=%d (needs 64 channels, 32-pixel rows)", Ca, Ws);
    CUtensorMap mA, mB;
    int rc;
    if ((rc = map_small(&mA, in_small, Ca, Ws, Hs, B, Ws, 4, 1))) return rc;
    if ((rc = map_gemm_operand(&mB, w_down_e, 1, 64, 64, 64))) return rc;
    const int tiles_img = (2 * Hs + 6) / 6;                // tile t completes output rows 6t-1 .. 6t+4
    const int total = B * tiles_img;
    static DeviceOnce cfg;
    if (!cfg.done()) {
        cudaError_t e = cudaFuncSetAttribute(edge_up_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEUSmem);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "edge_up_scatter smem attr: %s", cudaGetErrorString(e));
        cfg.mark();
    }
    const int grid = total < 2 * kNumSMs ? total : 2 * kNumSMs;
    launch_pdl(edge_up_scatter_kernel, dim3(grid), dim3(192), kEUSmem, as_stream(stream), mA, mB, (__nv_bfloat16*)img_p4, B, Hs,
               tiles_img, total);
    JCK_LAUNCH_CHECK("edge_up_scatter");
    return JCK_OK;
}
