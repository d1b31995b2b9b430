// Inception-v3 feature extractor of the reference's metrics.py (FID / IS evaluation, metrics.py:46-52,80-93) on sm_100a:
//   * conv_gemm_kernel   every BasicConv2d (conv, eval-mode BatchNorm folded to scale/shift, ReLU) and the final Linear as an
//                        implicit GEMM on tcgen05: activations are NHWC bf16 buffers that carry their own zero border, so
//                        for a stride-1 kh x kw convolution filter tap (ky, kx) of 128 consecutive buffer positions is the
//                        SAME 128 rows shifted by a constant ((ky - pady) * Wq + (kx - padx)) -- one 2-D TMA box per tap,
//                        no im2col, any kernel shape (1x1, 3x3, 5x5, 1x7, 7x1, 1x3, 3x1).  Positions on the border compute
//                        garbage that the epilogue masks; valid rows are re-mapped into the (differently padded, possibly
//                        wider = concatenated) output buffer.  Persistent CTAs, TMA ring, double-buffered TMEM accumulator.
//   * im2col_kernel      the five stride-2 convolutions (and the 3-channel stem) as explicit patch matrices for the same GEMM
//   * pool3_kernel       max 3x3 s2 / avg 3x3 s1 p1 (count_include_pad) into a channel slice of the block output
//   * global_avgpool     8x8 -> 1x1 (pool3 features, fp32 + bf16)
//   * resize_norm_kernel the eval branch's pre-processing (dcgan_trainer.py:203-207: 0.5x+0.5, F.resize to 299x299
//                        bilinear, ImageNet normalise) fused, NCHW fp32 -> NHWC bf16
//   * inception_score    metrics.py:96-110: softmax, split marginals, KL, exp(mean) per split
#include <stdlib.h>
#include "tc_common.cuh"

namespace jck {

int encode_bf16_2d(CUtensorMap* m, const void* base, unsigned long long inner, unsigned long long rows,
                   unsigned long long pitch_bytes, unsigned box_inner, unsigned box_rows, int swizzle64);

namespace {
using namespace tc;

constexpr int kCGThreads = 192;
constexpr int kCGMaxStages = 8;
constexpr int kCGMaxTaps = 96;      // 5 x 5 taps x 3 (split-precision mode: hi*hi, hi*lo, lo*hi)
constexpr int kCGAccCols = 256;

// K-major operand tile [rows][kw bf16], kw = 64 (128-byte swizzle, layout 2) or 32 (64-byte swizzle, layout 4): 8-row groups
// of rows x (2 kw) bytes, +32 bytes per UMMA_K = 16 step
// (the address-independent bits are built once per kernel: the single issuing thread must not spend its cycles on them)
__device__ __forceinline__ uint64_t cg_sdesc_hi(int kw) {
    uint64_t d = 0;
    d |= (uint64_t)(((uint32_t)(kw * 16) >> 4) & 0x3FFFu) << 32;       // SBO: 8 rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(kw == 64 ? 2 : 4) << 61;
    return d;
}
__device__ __forceinline__ uint64_t cg_sdesc(uint64_t hi, uint32_t smem_addr) { return hi | (uint64_t)((smem_addr >> 4) & 0x3FFFu); }

struct CGParams {
    int M, N, BN, n_tiles, total_tiles;
    int csteps, ksteps, Cp;
    int kw, a_bytes;           // K chunk per stage: 64 channels (16 KB of A) or, for C <= 32, 32 channels (8 KB, 64-byte swizzle)
    int shift[kCGMaxTaps];
    int Hq, Wq, oy0, ox0, Ho, Wo, Hob, Wob, opy, opx, c_off, relu, out_f32;
    long long ldc;
    long long lo_out;          // split-precision output (bf16 only): element offset of the LOW plane of `out`; 0 = plain bf16
    int stages, stage_bytes, vec_coef;
    // window mode (conv_gemm_window_kernel): the rows [m0 + smin, m0 + 128 + smax) of a tile are loaded ONCE per 64-channel
    // chunk and every tap reads them at a row offset through its shared-memory descriptor
    int smin, wboxes, win_bytes, wstages;
    int resident;              // 1: the whole weight matrix (ksteps boxes of BN x 64) stays in shared memory; the ring carries A only
    const float* scale;
    const float* bias;
};

// 16 consecutive bf16 outputs of one row: 256-bit, 2 x 128-bit or element stores, whatever the address allows
__device__ __forceinline__ void cg_store16(__nv_bfloat16* o, const float (&v)[16]) {
    uint4 u0, u1;
    u0.x = pack_bf16x2(v[0], v[1]); u0.y = pack_bf16x2(v[2], v[3]);
    u0.z = pack_bf16x2(v[4], v[5]); u0.w = pack_bf16x2(v[6], v[7]);
    u1.x = pack_bf16x2(v[8], v[9]); u1.y = pack_bf16x2(v[10], v[11]);
    u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
    const uintptr_t a = reinterpret_cast<uintptr_t>(o);
    if ((a & 31) == 0) {
        st_global_256(o, u0, u1);
    } else if ((a & 15) == 0) {
        reinterpret_cast<uint4*>(o)[0] = u0;
        reinterpret_cast<uint4*>(o)[1] = u1;
    } else {
        const __nv_bfloat16* h0 = reinterpret_cast<const __nv_bfloat16*>(&u0);
        const __nv_bfloat16* h1 = reinterpret_cast<const __nv_bfloat16*>(&u1);
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i] = h0[i]; o[8 + i] = h1[i]; }
    }
}

// Epilogue warps (2..5) of both conv kernels: TMEM accumulator -> scale / bias / ReLU -> masked, re-mapped store.
__device__ __forceinline__ void cg_epilogue(const CGParams& p, void* __restrict__ out, uint32_t tmem_base, uint64_t* tfull,
                                            uint64_t* tempty, int warp, int lane) {
    const int wq = warp & 3;
    const int plane = p.Hq * p.Wq;
    const int chunks = p.BN >> 4;
    int lt = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
        const int nt = t % p.n_tiles, mt = t / p.n_tiles;
        const int m = mt * 128 + wq * 32 + lane, n0 = nt * p.BN;
        const int b = m / plane, r = m - b * plane;
        const int Y = r / p.Wq, X = r - Y * p.Wq;
        const int oy = Y - p.oy0, ox = X - p.ox0;
        const bool valid = m < p.M && oy >= 0 && oy < p.Ho && ox >= 0 && ox < p.Wo;
        const long long orow = (((long long)b * p.Hob + oy + p.opy) * p.Wob + ox + p.opx) * p.ldc + p.c_off;
        const int acc = lt & 1;
        const uint32_t tmem_d = tmem_base + acc * kCGAccCols + ((uint32_t)(wq * 32) << 16);
        mbar_wait(&tfull[acc], (lt >> 1) & 1);
        fence_after_sync();
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
            const int nb = n0 + c * 16;
            const bool full16 = nb + 16 <= p.N;
            // this chunk's BatchNorm coefficients (L1-resident after the CTA's first tile), issued before the TMEM read
            float4 sc[4], bi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                sc[q] = make_float4(1.f, 1.f, 1.f, 1.f);
                bi[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (full16 && p.vec_coef) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (p.scale) sc[q] = __ldg(reinterpret_cast<const float4*>(p.scale + nb) + q);
                    if (p.bias) bi[q] = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + q);
                }
            }
            float v[16];
            tmem_ld16(tmem_d + c * 16, v);
            tmem_ld_wait();
            if (c == chunks - 1) {
                fence_before_sync();
                mbar_arrive(&tempty[acc]);
            }
            if (!valid || nb >= p.N) continue;
            if (full16 && p.vec_coef) {
                const float* scf = reinterpret_cast<const float*>(sc);
                const float* bif = reinterpret_cast<const float*>(bi);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x = fmaf(v[i], scf[i], bif[i]);
                    v[i] = p.relu ? fmaxf(x, 0.f) : x;
                }
                if (p.out_f32) {
                    float* o = reinterpret_cast<float*>(out) + orow + nb;
                    if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            reinterpret_cast<float4*>(o)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = v[i];
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + orow + nb;
                    cg_store16(o, v);
                    if (p.lo_out) {
                        // split precision: the value is kept as hi + lo, hi = bf16(v), lo = bf16(v - hi) (16 mantissa bits)
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] -= __bfloat162float(__float2bfloat16_rn(v[i]));
                        cg_store16(o + p.lo_out, v);
                    }
                }
            } else {
                // ragged last chunk (the fc layer's N = 100) or unaligned coefficient vectors: element by element
                const int nv = min(16, p.N - nb);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (i < nv) {
                        const float s1 = p.scale ? __ldg(p.scale + nb + i) : 1.f;
                        const float b1 = p.bias ? __ldg(p.bias + nb + i) : 0.f;
                        float x = fmaf(v[i], s1, b1);
                        if (p.relu) x = fmaxf(x, 0.f);
                        if (p.out_f32) {
                            reinterpret_cast<float*>(out)[orow + nb + i] = x;
                        } else {
                            const __nv_bfloat16 h = __float2bfloat16_rn(x);
                            reinterpret_cast<__nv_bfloat16*>(out)[orow + nb + i] = h;
                            if (p.lo_out)
                                reinterpret_cast<__nv_bfloat16*>(out)[orow + nb + i + p.lo_out] = __float2bfloat16_rn(x - __bfloat162float(h));
                        }
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kCGThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void* __restrict__ out,
                 const __grid_constant__ CGParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    const int b_bytes = p.BN * 2 * p.kw;                               // one K step of the weight tile
    uint8_t* wres = smem + stages * p.stage_bytes;                     // resident weights (p.resident), else unused
    uint64_t* full = reinterpret_cast<uint64_t*>(wres + (p.resident ? p.ksteps * b_bytes : 0));
    uint64_t* empty = full + kCGMaxStages;
    uint64_t* tfull = empty + kCGMaxStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kCGAccCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = p.a_bytes + (p.resident ? 0 : b_bytes);
            if (p.resident) {                       // n_tiles == 1: every tile of this CTA uses the same weights
                mbar_arrive_expect_tx(wfull, (uint32_t)(p.ksteps * b_bytes));
                for (int ks = 0; ks < p.ksteps; ++ks) {
                    const int tap = ks / p.csteps, cc = ks - tap * p.csteps;
                    tma_load_2d(wres + ks * b_bytes, &mapB, wfull, tap * p.Cp + cc * p.kw, 0);
                }
            }
            int it = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int nt = t % p.n_tiles, mt = t / p.n_tiles;
                const int m0 = mt * 128, n0 = nt * p.BN;
                for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
                    const int s = it % stages;
                    mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
                    uint8_t* sa = smem + s * p.stage_bytes;
                    uint8_t* sb = sa + p.a_bytes;
                    mbar_arrive_expect_tx(&full[s], tx);
                    const int tap = ks / p.csteps, cc = ks - tap * p.csteps;
                    tma_load_2d(sa, &mapA, &full[s], cc * p.kw, m0 + p.shift[tap]);
                    if (!p.resident) tma_load_2d(sb, &mapB, &full[s], tap * p.Cp + cc * p.kw, n0);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc(p.BN, 0, 0);
        const uint64_t dhi = cg_sdesc_hi(p.kw);
        const bool wide = p.kw == 64;
        if (p.resident) {
            mbar_wait(wfull, 0);
            fence_after_sync();
        }
        int it = 0, lt = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
            const int acc = lt & 1;
            mbar_wait(&tempty[acc], ((lt >> 1) & 1) ^ 1);
            fence_after_sync();
            const uint32_t tmem_d = tmem_base + acc * kCGAccCols;
            for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
                const int s = it % stages;
                mbar_wait(&full[s], (it / stages) & 1);
                fence_after_sync();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + s * p.stage_bytes);
                    const uint32_t b_addr = p.resident ? smem_u32(wres + ks * b_bytes) : a_addr + p.a_bytes;
                    if (wide) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_d, cg_sdesc(dhi, a_addr + k * 32), cg_sdesc(dhi, b_addr + k * 32), idesc, (ks > 0 || k > 0) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            umma_bf16(tmem_d, cg_sdesc(dhi, a_addr + k * 32), cg_sdesc(dhi, b_addr + k * 32), idesc, (ks > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                    if (ks == p.ksteps - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        cg_epilogue(p, out, tmem_base, tfull, tempty, warp, lane);
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 2 * kCGAccCols);
    }
}

// ------------------------------------------------------------------------------------------------
// Window variant.  The kernel above re-reads the activations once per filter tap (nine shifted boxes for a 3 x 3), which
// makes every multi-tap layer L2 -> shared-memory bound (profiles/r01_ncu_incep.md).  Here the producer loads the row WINDOW
// [m0 + smin, m0 + 128 + smax) of a tile once per channel chunk (wboxes boxes of 128 rows), and tap t is the SAME shared
// memory read through a descriptor that starts (shift[t] - smin) rows further down.  Measured on sm_100a: the tensor core
// applies the 128-byte (64-byte) swizzle to the ABSOLUTE shared-memory address, exactly as TMA wrote it, so a start address
// that is not a multiple of the 8-row pattern needs NO base offset in the descriptor (leaving bits 49-51 zero gives exact
// results for every shift; setting them to (address >> 7) & 7 gives wrong ones).  K order is chunk-major / tap-minor; the
// weights are still indexed [tap][chunk].
// Weights: resident (as above) or streamed through the `full / empty` ring, one BN x kw tile per (chunk, tap).
// ------------------------------------------------------------------------------------------------
constexpr int kCGMaxWin = 4;

__global__ void __launch_bounds__(kCGThreads, 1)
conv_gemm_window_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, void* __restrict__ out,
                        const __grid_constant__ CGParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages, wstages = p.wstages;
    const int b_bytes = p.BN * 2 * p.kw;
    const int ntaps = p.ksteps / p.csteps;
    uint8_t* win = smem;                                              // [wstages][win_bytes]
    uint8_t* bring = smem + wstages * p.win_bytes;                    // resident: [ksteps][b_bytes]; else ring [stages][b_bytes]
    uint64_t* full = reinterpret_cast<uint64_t*>(bring + (p.resident ? p.ksteps : stages) * b_bytes);
    uint64_t* empty = full + kCGMaxStages;
    uint64_t* tfull = empty + kCGMaxStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint64_t* winf = wfull + 1;
    uint64_t* wine = winf + kCGMaxWin;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wine + kCGMaxWin);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&mapA);
        prefetch_tmap(&mapB);
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < wstages; ++s) { mbar_init(&winf[s], 1); mbar_init(&wine[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * kCGAccCols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            if (p.resident) {
                mbar_arrive_expect_tx(wfull, (uint32_t)(p.ksteps * b_bytes));
                for (int ks = 0; ks < p.ksteps; ++ks) {
                    const int tap = ks / p.csteps, cc = ks - tap * p.csteps;
                    tma_load_2d(bring + ks * b_bytes, &mapB, wfull, tap * p.Cp + cc * p.kw, 0);
                }
            }
            // the window of job (tile, chunk) j + 1 is requested BEFORE the weight tiles of job j, so that its latency hides
            // behind job j's MMAs instead of behind the weight ring's back-pressure
            int iw = 0, ib = 0;
            int tw = blockIdx.x, ccw = 0;
            auto issue_window = [&]() {
                if (tw >= p.total_tiles) return;
                const int m0w = (tw / p.n_tiles) * 128;
                const int ws = iw % wstages;
                mbar_wait(&wine[ws], ((iw / wstages) & 1) ^ 1);
                mbar_arrive_expect_tx(&winf[ws], (uint32_t)(p.wboxes * p.a_bytes));
                for (int bx = 0; bx < p.wboxes; ++bx)
                    tma_load_2d(win + ws * p.win_bytes + bx * p.a_bytes, &mapA, &winf[ws], ccw * p.kw, m0w + p.smin + bx * 128);
                ++iw;
                if (++ccw == p.csteps) { ccw = 0; tw += gridDim.x; }
            };
            issue_window();
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int n0 = (t % p.n_tiles) * p.BN;
                for (int cc = 0; cc < p.csteps; ++cc) {
                    issue_window();
                    if (!p.resident) {
                        for (int tap = 0; tap < ntaps; ++tap, ++ib) {
                            const int s = ib % stages;
                            mbar_wait(&empty[s], ((ib / stages) & 1) ^ 1);
                            mbar_arrive_expect_tx(&full[s], (uint32_t)b_bytes);
                            tma_load_2d(bring + s * b_bytes, &mapB, &full[s], tap * p.Cp + cc * p.kw, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc(p.BN, 0, 0);
        const uint64_t dhi = cg_sdesc_hi(p.kw);
        const int row_bytes = 2 * p.kw;
        const int nk = p.kw >> 4;
        if (p.resident) {
            mbar_wait(wfull, 0);
            fence_after_sync();
        }
        int iw = 0, ib = 0, lt = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++lt) {
            const int acc = lt & 1;
            mbar_wait(&tempty[acc], ((lt >> 1) & 1) ^ 1);
            fence_after_sync();
            const uint32_t tmem_d = tmem_base + acc * kCGAccCols;
            for (int cc = 0; cc < p.csteps; ++cc, ++iw) {
                const int ws = iw % wstages;
                mbar_wait(&winf[ws], (iw / wstages) & 1);
                fence_after_sync();
                const uint32_t w_addr = smem_u32(win + ws * p.win_bytes);
                for (int tap = 0; tap < ntaps; ++tap) {
                    int s = 0;
                    if (!p.resident) {
                        s = ib % stages;
                        mbar_wait(&full[s], (ib / stages) & 1);
                        fence_after_sync();
                        ++ib;
                    }
                    if (lane == 0) {
                        const uint32_t a_addr = w_addr + (uint32_t)((p.shift[tap] - p.smin) * row_bytes);
                        const uint32_t b_addr = smem_u32(bring + (p.resident ? (tap * p.csteps + cc) : s) * b_bytes);
                        for (int k = 0; k < nk; ++k)
                            umma_bf16(tmem_d, cg_sdesc(dhi, a_addr + k * 32), cg_sdesc(dhi, b_addr + k * 32), idesc,
                                      (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        if (!p.resident) umma_commit(&empty[s]);
                    }
                    __syncwarp();
                }
                if (lane == 0) {
                    umma_commit(&wine[ws]);
                    if (cc == p.csteps - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        cg_epilogue(p, out, tmem_base, tfull, tempty, warp, lane);
    }

    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tmem_dealloc(tmem_base, 2 * kCGAccCols);
    }
}

// ---- buffer geometry shared by the streaming kernels: logical pixel (y, x) of image b lives at
//      ((b * Hb + y + by) * Wb + x + bx) * ld + c_off
struct BufGeom { int Hb, Wb, by, bx, c_off; long long ld; };
__device__ __forceinline__ long long buf_at(const BufGeom& g, int b, int y, int x) {
    return (((long long)b * g.Hb + y + g.by) * g.Wb + x + g.bx) * g.ld + g.c_off;
}

// patches[m][k], m = (b, oy, ox), k = (ky * kw + kx) * C + c, zero outside the image and for k >= kh*kw*C
template <int VEC>
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, BufGeom g, int B, int H, int W, int C, int kh,
              int kw, int sy, int sx, int py, int px, int Ho, int Wo, int Kp) {
    pdl_entry();
    const int kv = Kp / VEC;
    const long long total = (long long)B * Ho * Wo * kv;
    const int K = kh * kw * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % kv) * VEC;
        const long long m = i / kv;
        const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), b = (int)(m / ((long long)Wo * Ho));
        __nv_bfloat16* o = out + m * Kp + k;
        if (VEC == 8) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (k < K) {
                const int tap = k / C, c = k - tap * C;
                const int iy = oy * sy + tap / kw - py, ix = ox * sx + tap % kw - px;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(reinterpret_cast<const uint4*>(x + buf_at(g, b, iy, ix) + c));
            }
            *reinterpret_cast<uint4*>(o) = v;
        } else {
            __nv_bfloat16 v = __float2bfloat16_rn(0.f);
            if (k < K) {
                const int tap = k / C, c = k - tap * C;
                const int iy = oy * sy + tap / kw - py, ix = ox * sx + tap % kw - px;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[buf_at(g, b, iy, ix) + c];
            }
            *o = v;
        }
    }
}

// 3x3 pooling, 8 channels per thread.  mode 0: max (no padding, window inside the image); 1: average with zero padding 1,
// divisor 9 (F.avg_pool2d default count_include_pad=True, torchvision inception.py branch_pool)
__global__ void __launch_bounds__(256)
pool3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, BufGeom gi, BufGeom go, int B, int H, int W, int C,
             int stride, int pad, int Ho, int Wo, int mode, long long x_lo, long long out_lo) {
    pdl_entry();
    const int cv = C / 8;
    const long long total = (long long)B * Ho * Wo * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * 8;
        const long long m = i / cv;
        const int ox = (int)(m % Wo), oy = (int)((m / Wo) % Ho), b = (int)(m / ((long long)Wo * Ho));
        // all nine 128-bit loads are issued before the first use (clamped addresses; windows cut by the border are masked)
        uint4 u[9], ul[9];
        bool ok[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iy = oy * stride + ky - pad, ix = ox * stride + kx - pad;
                ok[ky * 3 + kx] = iy >= 0 && iy < H && ix >= 0 && ix < W;
                const int cy = min(max(iy, 0), H - 1), cx = min(max(ix, 0), W - 1);
                const __nv_bfloat16* src = x + buf_at(gi, b, cy, cx) + c;
                u[ky * 3 + kx] = __ldg(reinterpret_cast<const uint4*>(src));
                ul[ky * 3 + kx] = x_lo ? __ldg(reinterpret_cast<const uint4*>(src + x_lo)) : make_uint4(0, 0, 0, 0);   // split: value = hi + lo
            }
        }
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = mode == 0 ? -INFINITY : 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&u[t]);
            const __nv_bfloat16* l = reinterpret_cast<const __nv_bfloat16*>(&ul[t]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float f = __bfloat162float(h[j]) + __bfloat162float(l[j]);
                if (mode == 0) a[j] = ok[t] ? fmaxf(a[j], f) : a[j];
                else a[j] += ok[t] ? f : 0.f;
            }
        }
        if (mode == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] *= (1.f / 9.f);
        }
        uint4 r;
        r.x = pack_bf16x2(a[0], a[1]); r.y = pack_bf16x2(a[2], a[3]); r.z = pack_bf16x2(a[4], a[5]); r.w = pack_bf16x2(a[6], a[7]);
        __nv_bfloat16* dst = out + buf_at(go, b, oy, ox) + c;
        *reinterpret_cast<uint4*>(dst) = r;
        if (out_lo) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] -= __bfloat162float(__float2bfloat16_rn(a[j]));
            r.x = pack_bf16x2(a[0], a[1]); r.y = pack_bf16x2(a[2], a[3]); r.z = pack_bf16x2(a[4], a[5]); r.w = pack_bf16x2(a[6], a[7]);
            *reinterpret_cast<uint4*>(dst + out_lo) = r;
        }
    }
}

// mean over the HW pixels of a dense [B][HW][C] bf16 tensor -> fp32 [B][C] and bf16 [B][C] (either nullable)
__global__ void __launch_bounds__(256)
global_avgpool_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int B,
                      int HW, int C, long long x_lo, long long out_lo) {
    pdl_entry();
    const long long total = (long long)B * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C), b = (int)(i / C);
        float s = 0.f;
        for (int q = 0; q < HW; ++q) {
            const long long at = ((long long)b * HW + q) * C + c;
            s += __bfloat162float(x[at]) + (x_lo ? __bfloat162float(x[at + x_lo]) : 0.f);
        }
        s /= (float)HW;
        if (out_f32) out_f32[i] = s;
        if (out_bf16) {
            const __nv_bfloat16 h = __float2bfloat16_rn(s);
            out_bf16[i] = h;
            if (out_lo) out_bf16[i + out_lo] = __float2bfloat16_rn(s - __bfloat162float(h));
        }
    }
}

// out[b][y][x][c] (NHWC bf16, channel pitch ldo, pad channels zero) = ((a * bilinear(in)[b][c][y][x] + bb) - mean[c]) / std[c]
// bilinear: half-pixel centres, source index clamped at 0 (torch upsample_bilinear2d, align_corners=False, no antialias)
__global__ void __launch_bounds__(256)
resize_norm_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int C, int Hi, int Wi, int Ho, int Wo, int ldo,
                   float a, float bb, float m0, float m1, float m2, float s0, float s1, float s2) {
    pdl_entry();
    const long long total = (long long)B * Ho * Wo;
    const float ry = (float)Hi / (float)Ho, rx = (float)Wi / (float)Wo;
    const float mean[3] = {m0, m1, m2}, istd[3] = {s0, s1, s2};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho), b = (int)(i / ((long long)Wo * Ho));
        const float fy = fmaxf(ry * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(rx * ((float)x + 0.5f) - 0.5f, 0.f);
        const int y0 = min((int)fy, Hi - 1), x0 = min((int)fx, Wi - 1);
        const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
        const float wy = fy - (float)y0, wx = fx - (float)x0;
        __nv_bfloat16* o = out + i * ldo;
        for (int c = 0; c < ldo; ++c) {
            float v = 0.f;
            if (c < C) {
                const float* pl = in + ((long long)b * C + c) * Hi * Wi;
                const float v00 = __ldg(pl + y0 * Wi + x0), v01 = __ldg(pl + y0 * Wi + x1);
                const float v10 = __ldg(pl + y1 * Wi + x0), v11 = __ldg(pl + y1 * Wi + x1);
                const float top = v00 + wx * (v01 - v00), bot = v10 + wx * (v11 - v10);
                v = top + wy * (bot - top);
                v = (a * v + bb - mean[c < 3 ? c : 0]) * istd[c < 3 ? c : 0];
            }
            o[c] = __float2bfloat16_rn(v);
        }
    }
}

// The stem in one pass: patch matrix of Conv2d_1a_3x3 (3x3, stride 2, no padding, 3 channels -> K = 27, row pitch 32) taken
// straight from the un-resized images: patches[(b, oy, ox)][(ky*3 + kx)*3 + c] = bf16(resize_norm(in)(b, c, 2oy + ky, 2ox + kx)),
// the same values resize_norm_kernel + im2col_kernel would produce (each resized pixel is rounded to bf16 first).
__global__ void __launch_bounds__(256)
stem_patches_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int Hi, int Wi, int Hr, int Wr, int Ho, int Wo,
                    float a, float bb, float m0, float m1, float m2, float s0, float s1, float s2, long long out_lo) {
    pdl_entry();
    const long long total = (long long)B * Ho * Wo;
    const float ry = (float)Hi / (float)Hr, rx = (float)Wi / (float)Wr;
    const float mean[3] = {m0, m1, m2}, istd[3] = {s0, s1, s2};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), b = (int)(i / ((long long)Wo * Ho));
        const float* img = in + (long long)b * 3 * Hi * Wi;
        __align__(16) __nv_bfloat16 row[32], row_lo[32];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float fy = fmaxf(ry * ((float)(2 * oy + ky) + 0.5f) - 0.5f, 0.f);
            const int y0 = min((int)fy, Hi - 1), y1 = min(y0 + 1, Hi - 1);
            const float wy = fy - (float)y0;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float fx = fmaxf(rx * ((float)(2 * ox + kx) + 0.5f) - 0.5f, 0.f);
                const int x0 = min((int)fx, Wi - 1), x1 = min(x0 + 1, Wi - 1);
                const float wx = fx - (float)x0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* pl = img + (long long)c * Hi * Wi;
                    const float v00 = __ldg(pl + y0 * Wi + x0), v01 = __ldg(pl + y0 * Wi + x1);
                    const float v10 = __ldg(pl + y1 * Wi + x0), v11 = __ldg(pl + y1 * Wi + x1);
                    const float top = v00 + wx * (v01 - v00), bot = v10 + wx * (v11 - v10);
                    const float v = top + wy * (bot - top);
                    const float val = (a * v + bb - mean[c]) * istd[c];
                    const __nv_bfloat16 h = __float2bfloat16_rn(val);
                    row[(ky * 3 + kx) * 3 + c] = h;
                    row_lo[(ky * 3 + kx) * 3 + c] = __float2bfloat16_rn(val - __bfloat162float(h));
                }
            }
        }
#pragma unroll
        for (int k = 27; k < 32; ++k) row[k] = row_lo[k] = __float2bfloat16_rn(0.f);
        uint4* o = reinterpret_cast<uint4*>(out + i * 32);
        const uint4* r = reinterpret_cast<const uint4*>(row);
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = r[q];
        if (out_lo) {
            uint4* ol = reinterpret_cast<uint4*>(out + out_lo + i * 32);
            const uint4* rl = reinterpret_cast<const uint4*>(row_lo);
#pragma unroll
            for (int q = 0; q < 4; ++q) ol[q] = rl[q];
        }
    }
}

// Inception score of one split per block (metrics.py:96-110): rows [r0, r1) of fp32 logits [n][d]:
//   p = softmax(row);  py = mean_rows p;  score = exp(mean_rows sum_j p_j log(p_j / py_j))
__global__ void __launch_bounds__(256)
inception_score_kernel(const float* __restrict__ logits, int n, int d, int splits, float* __restrict__ scores) {
    pdl_entry();
    extern __shared__ float sh[];
    float* py = sh;                       // [d]
    __shared__ double part[8];
    const int per = n / splits;
    const int r0 = blockIdx.x * per, r1 = r0 + per;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = threadIdx.x; j < d; j += blockDim.x) py[j] = 0.f;
    __syncthreads();
    for (int r = r0 + warp; r < r1; r += nw) {
        const float* row = logits + (long long)r * d;
        float mx = -INFINITY;
        for (int j = lane; j < d; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int j = lane; j < d; j += 32) s += expf(row[j] - mx);
        s = warp_sum(s);
        for (int j = lane; j < d; j += 32) atomicAdd(py + j, expf(row[j] - mx) / s);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += blockDim.x) py[j] /= (float)per;
    __syncthreads();
    double kl = 0.0;
    for (int r = r0 + warp; r < r1; r += nw) {
        const float* row = logits + (long long)r * d;
        float mx = -INFINITY;
        for (int j = lane; j < d; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int j = lane; j < d; j += 32) s += expf(row[j] - mx);
        s = warp_sum(s);
        float a = 0.f;
        for (int j = lane; j < d; j += 32) {
            const float pj = expf(row[j] - mx) / s;
            if (pj > 0.f) a += pj * logf(pj / py[j]);
        }
        a = warp_sum(a);
        kl += (double)a;
    }
    if (lane == 0) part[warp] = kl;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += part[w];
        scores[blockIdx.x] = per > 0 ? (float)exp(t / (double)per) : 0.f;
    }
}

int grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = 16LL * kNumSMs;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace
}  // namespace jck

using namespace jck;

extern "C" int jck_conv_gemm(const void* act, long long lda, const void* w, const float* scale, const float* bias, void* out,
                             long long ldc, const int* geom, int ngeom, void* stream) {
    JCK_REQUIRE(act && w && out && geom && ngeom >= 18, "conv_gemm: bad argument");
    const int M = geom[0], N = geom[1], C = geom[2], ntaps = geom[3];
    JCK_REQUIRE(M > 0 && N > 0 && C > 0 && ntaps > 0 && ntaps <= kCGMaxTaps && ngeom >= 18 + ntaps, "conv_gemm: bad geometry");
    if (lda % 8 != 0 || ((uintptr_t)act & 15) || ((uintptr_t)w & 15))
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "conv_gemm: operand pitch / base not 16-byte aligned (lda=%lld)", lda);
    CGParams p;
    p.M = M; p.N = N;
    // K chunk: 64 channels; layers with at most 32 channels (the stem) use 32-wide chunks so that no box is half zero-fill
    p.kw = (C <= 32 && getenv("JCK_CG_NO_K32") == nullptr) ? 32 : 64;
    p.a_bytes = 128 * p.kw * 2;
    p.csteps = (C + p.kw - 1) / p.kw;
    p.Cp = C <= 32 ? 32 : (C + 63) / 64 * 64;      // the weight matrix's pitch per tap (the caller's packing)
    p.ksteps = ntaps * p.csteps;
    p.Hq = geom[4]; p.Wq = geom[5]; p.oy0 = geom[6]; p.ox0 = geom[7]; p.Ho = geom[8]; p.Wo = geom[9];
    p.Hob = geom[10]; p.Wob = geom[11]; p.opy = geom[12]; p.opx = geom[13]; p.c_off = geom[14]; p.relu = geom[15];
    p.out_f32 = geom[16] == JCK_F32;
    const long long rows_a = (long long)geom[17];
    JCK_REQUIRE(p.Hq > 0 && p.Wq > 0 && M % (p.Hq * p.Wq) == 0 && rows_a > 0, "conv_gemm: row space %d is not B x %d x %d", M, p.Hq, p.Wq);
    for (int t = 0; t < ntaps; ++t) p.shift[t] = geom[18 + t];
    p.ldc = ldc;
    // optional trailing entry: split-precision output, the LOW plane lies geom[18 + ntaps] output rows behind the high one
    p.lo_out = (ngeom > 18 + ntaps) ? (long long)geom[18 + ntaps] * ldc : 0;
    JCK_REQUIRE(p.lo_out == 0 || (!p.out_f32 && p.lo_out > 0 && p.lo_out % 8 == 0), "conv_gemm: split output needs bf16 and an 8-element aligned plane offset");
    const int nt = (N + 255) / 256;
    p.BN = (((N + nt - 1) / nt) + 15) / 16 * 16;
    p.n_tiles = (N + p.BN - 1) / p.BN;
    const int m_tiles = (M + 127) / 128;
    p.total_tiles = m_tiles * p.n_tiles;
    // small filter banks stay resident in shared memory for the CTA's whole tile walk (the stem / 1x1 layers would otherwise
    // re-stream as many weight bytes as activation bytes); the ring then carries A only
    const int w_bytes = p.ksteps * p.BN * 2 * p.kw;
    p.resident = (p.n_tiles == 1 && w_bytes <= 96 * 1024 && getenv("JCK_CG_NO_RESIDENT") == nullptr) ? 1 : 0;
    p.stage_bytes = p.a_bytes + (p.resident ? 0 : p.BN * 2 * p.kw);
    p.stages = (200 * 1024 - (p.resident ? w_bytes : 0)) / p.stage_bytes;
    if (p.stages > kCGMaxStages) p.stages = kCGMaxStages;
    p.scale = scale;
    p.bias = bias;
    p.vec_coef = (((uintptr_t)scale & 15) == 0 && ((uintptr_t)bias & 15) == 0) ? 1 : 0;
    // always more than half of the SM's shared memory: ONE CTA per SM.  Two co-resident CTAs (possible with the 8 KB stages of
    // the 32-channel mode; with programmatic dependent launch also a CTA of the NEXT launch) would serialise on the 512 TMEM
    // columns each allocates and leave half of the SMs idle (measured: 6.0 -> 7.7 ms per 128 images)
    int smem = p.stages * p.stage_bytes + (p.resident ? w_bytes : 0) + 256 + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;
    CUtensorMap mA, mB;
    int rc;
    const int sw64 = p.kw == 32;
    if ((rc = encode_bf16_2d(&mA, act, (unsigned long long)C, (unsigned long long)rows_a, (unsigned long long)lda * 2, (unsigned)p.kw, 128,
                             sw64)))
        return rc;
    if ((rc = encode_bf16_2d(&mB, w, (unsigned long long)ntaps * p.Cp, (unsigned long long)N, (unsigned long long)ntaps * p.Cp * 2,
                             (unsigned)p.kw, (unsigned)p.BN, sw64)))
        return rc;
    static DeviceOnce cfg;
    if (!cfg.done()) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "conv_gemm smem attr: %s", cudaGetErrorString(e));
        cfg.mark();
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    // window kernel: multi-tap layers whose filter bank is resident (measured: the two 32-channel 3 x 3 stem layers 531 -> 480 and
    // 547 -> 496 us per 128 images); with streamed weights it measured SLOWER than the tap-streaming kernel (Conv2d_4a 338 -> 398
    // us) before the producer was reordered, so those layers stay on the kernel above unless JCK_CG_WINDOW=all; JCK_CG_WINDOW=0
    // turns the window kernel off
    const char* wenv = getenv("JCK_CG_WINDOW");
    const bool win_all = wenv && wenv[0] == 'a';
    const bool win_off = wenv && wenv[0] == '0';
    if (ntaps > 1 && !win_off && (p.resident || win_all)) {
        int smin = p.shift[0], smax = p.shift[0];
        for (int t = 1; t < ntaps; ++t) { smin = p.shift[t] < smin ? p.shift[t] : smin; smax = p.shift[t] > smax ? p.shift[t] : smax; }
        p.smin = smin;
        if (smax - smin > 8192) smax = smin + (1 << 20);     // split-precision taps reach into the low plane: no window fits (falls through)
        p.wboxes = (128 + (smax - smin) + 127) / 128;
        p.win_bytes = p.wboxes * p.a_bytes;
        const int b_bytes = p.BN * 2 * p.kw;
        p.stages = p.resident ? 1 : 4;                       // weight ring depth (unused when resident)
        const int b_region = p.resident ? w_bytes : p.stages * b_bytes;
        p.wstages = (200 * 1024 - b_region) / p.win_bytes;
        if (p.wstages > kCGMaxWin) p.wstages = kCGMaxWin;
        if (p.wstages >= 2) {
            int wsmem = p.wstages * p.win_bytes + b_region + 256 + 1024;
            if (wsmem < 120 * 1024) wsmem = 120 * 1024;
            launch_pdl(conv_gemm_window_kernel, dim3(grid), dim3(kCGThreads), wsmem, as_stream(stream), mA, mB, out, p);
            JCK_LAUNCH_CHECK("conv_gemm_window");
            return JCK_OK;
        }
        p.stages = (200 * 1024 - (p.resident ? w_bytes : 0)) / p.stage_bytes;      // does not fit: the tap-streaming kernel
        if (p.stages > kCGMaxStages) p.stages = kCGMaxStages;
    }
    launch_pdl(conv_gemm_kernel, dim3(grid), dim3(kCGThreads), smem, as_stream(stream), mA, mB, out, p);
    JCK_LAUNCH_CHECK("conv_gemm");
    return JCK_OK;
}

static BufGeom geom_of(const int* g, long long ld) { return BufGeom{g[0], g[1], g[2], g[3], g[4], ld}; }

extern "C" int jck_im2col(const void* x, const int* in_geom, long long ldx, void* patches, int B, int H, int W, int C, int kh, int kw,
                          int sy, int sx, int py, int px, int Ho, int Wo, int Kp, void* stream) {
    JCK_REQUIRE(x && in_geom && patches && B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && sy > 0 && sx > 0 && Ho > 0 && Wo > 0,
                "im2col: bad argument");
    JCK_REQUIRE(Kp >= kh * kw * C && Kp % 8 == 0, "im2col: Kp %d must be a multiple of 8 and >= %d", Kp, kh * kw * C);
    const BufGeom g = geom_of(in_geom, ldx);
    const bool vec = C % 8 == 0 && ldx % 8 == 0 && g.c_off % 8 == 0 && (((uintptr_t)x & 15) == 0);
    const long long total = (long long)B * Ho * Wo * (vec ? Kp / 8 : Kp);
    if (vec)
        launch_pdl(im2col_kernel<8>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)x,
                   (__nv_bfloat16*)patches, g, B, H, W, C, kh, kw, sy, sx, py, px, Ho, Wo, Kp);
    else
        launch_pdl(im2col_kernel<1>, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)x,
                   (__nv_bfloat16*)patches, g, B, H, W, C, kh, kw, sy, sx, py, px, Ho, Wo, Kp);
    JCK_LAUNCH_CHECK("im2col");
    return JCK_OK;
}

extern "C" int jck_pool3(const void* x, const int* in_geom, long long ldx, void* out, const int* out_geom, long long ldo, int B, int H,
                         int W, int C, int stride, int pad, int Ho, int Wo, int mode, void* stream) {
    return jck_pool3_split(x, in_geom, ldx, 0, out, out_geom, ldo, 0, B, H, W, C, stride, pad, Ho, Wo, mode, stream);
}

extern "C" int jck_pool3_split(const void* x, const int* in_geom, long long ldx, long long x_lo, void* out, const int* out_geom,
                               long long ldo, long long out_lo, int B, int H, int W, int C, int stride, int pad, int Ho, int Wo, int mode,
                               void* stream) {
    JCK_REQUIRE(x && out && in_geom && out_geom && B > 0 && H > 0 && W > 0 && C > 0 && (mode == 0 || mode == 1), "pool3: bad argument");
    JCK_REQUIRE(x_lo >= 0 && out_lo >= 0 && x_lo % 8 == 0 && out_lo % 8 == 0, "pool3: plane offsets must be multiples of 8 elements");
    const BufGeom gi = geom_of(in_geom, ldx), go = geom_of(out_geom, ldo);
    if (C % 8 != 0 || ldx % 8 != 0 || ldo % 8 != 0 || gi.c_off % 8 != 0 || go.c_off % 8 != 0)
        return set_error(JCK_E_UNSUPPORTED_SHAPE, "pool3: channels / pitches must be multiples of 8 (C=%d)", C);
    JCK_REQUIRE(mode == 1 || (pad == 0 && (Ho - 1) * stride + 3 <= H && (Wo - 1) * stride + 3 <= W), "pool3: max window leaves the image");
    const long long total = (long long)B * Ho * Wo * (C / 8);
    launch_pdl(pool3_kernel, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)x, (__nv_bfloat16*)out, gi,
               go, B, H, W, C, stride, pad, Ho, Wo, mode, x_lo, out_lo);
    JCK_LAUNCH_CHECK("pool3");
    return JCK_OK;
}

extern "C" int jck_global_avgpool(const void* x, float* out_f32, void* out_bf16, int B, int HW, int C, void* stream) {
    return jck_global_avgpool_split(x, 0, out_f32, out_bf16, 0, B, HW, C, stream);
}

extern "C" int jck_global_avgpool_split(const void* x, long long x_lo, float* out_f32, void* out_bf16, long long out_lo, int B, int HW,
                                        int C, void* stream) {
    JCK_REQUIRE(x && (out_f32 || out_bf16) && B > 0 && HW > 0 && C > 0 && x_lo >= 0 && out_lo >= 0, "global_avgpool: bad argument");
    launch_pdl(global_avgpool_kernel, dim3(grid_for((long long)B * C, 256)), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)x,
               out_f32, (__nv_bfloat16*)out_bf16, B, HW, C, x_lo, out_lo);
    JCK_LAUNCH_CHECK("global_avgpool");
    return JCK_OK;
}

extern "C" int jck_resize_norm(const float* in_nchw, void* out_nhwc, int B, int C, int Hi, int Wi, int Ho, int Wo, int ldo, float a,
                               float b, const float* mean3, const float* std3, void* stream) {
    JCK_REQUIRE(in_nchw && out_nhwc && mean3 && std3 && B > 0 && C > 0 && C <= 3 && ldo >= C && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0,
                "resize_norm: bad argument (host mean/std of 3 floats, C <= 3)");
    launch_pdl(resize_norm_kernel, dim3(grid_for((long long)B * Ho * Wo, 256)), dim3(256), 0, as_stream(stream), in_nchw,
               (__nv_bfloat16*)out_nhwc, B, C, Hi, Wi, Ho, Wo, ldo, a, b, mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1],
               1.f / std3[2]);
    JCK_LAUNCH_CHECK("resize_norm");
    return JCK_OK;
}

extern "C" int jck_stem_patches(const float* in_nchw, void* patches, int B, int Hi, int Wi, int Hr, int Wr, float a, float b,
                                const float* mean3, const float* std3, void* stream) {
    return jck_stem_patches_split(in_nchw, patches, 0, B, Hi, Wi, Hr, Wr, a, b, mean3, std3, stream);
}

extern "C" int jck_stem_patches_split(const float* in_nchw, void* patches, long long patches_lo, int B, int Hi, int Wi, int Hr, int Wr,
                                      float a, float b, const float* mean3, const float* std3, void* stream) {
    JCK_REQUIRE(in_nchw && patches && mean3 && std3 && B > 0 && Hi > 0 && Wi > 0 && Hr >= 3 && Wr >= 3 && patches_lo >= 0 &&
                patches_lo % 8 == 0, "stem_patches: bad argument");
    const int Ho = (Hr - 3) / 2 + 1, Wo = (Wr - 3) / 2 + 1;
    launch_pdl(stem_patches_kernel, dim3(grid_for((long long)B * Ho * Wo, 256)), dim3(256), 0, as_stream(stream), in_nchw,
               (__nv_bfloat16*)patches, B, Hi, Wi, Hr, Wr, Ho, Wo, a, b, mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1],
               1.f / std3[2], patches_lo);
    JCK_LAUNCH_CHECK("stem_patches");
    return JCK_OK;
}

extern "C" int jck_inception_score(const float* logits, int n, int d, int splits, float* scores, void* stream) {
    JCK_REQUIRE(logits && scores && n > 0 && d > 0 && splits > 0 && d <= 8192, "inception_score: bad argument");
    launch_pdl(inception_score_kernel, dim3(splits), dim3(256), (size_t)d * sizeof(float), as_stream(stream), logits, n, d, splits, scores);
    JCK_LAUNCH_CHECK("inception_score");
    return JCK_OK;
}
