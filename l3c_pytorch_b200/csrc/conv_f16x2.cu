// conv_f16x2.cu -- precision mode L3C_PREC_F16X2: the strict tensor-core mode.  Every operand x is carried as
// TWO FP16 numbers, hi = fp16(x) and lo = fp16((x - hi) * 2^11), so that x = hi + lo / 2^11 up to ~2^-22
// relative; a product is evaluated as
//        x * w  ~=  hi_x * hi_w  +  (hi_x * lo_w + lo_x * hi_w) / 2^11
// i.e. three tcgen05.mma (kind::f16) per K step into TWO fp32 accumulators in TMEM (the dropped lo*lo term is
// 2^-22 relative); the epilogue combines acc1 + acc2 * 2^-11 in fp32.  This is the error-compensated split
// SURVEY.md section 7 asks for ("3xTF32"), built from FP16 pieces: fp32-class results (tests: rtol 2e-5 against
// PyTorch fp32, the tolerance of the CUDA-core kernel) at several times the FFMA rate, so that the one-scale RGB
// baselines and small images -- where the 10-bit operands of the fast mode cost more than 1e-4 bpsp -- also run
// on the tensor cores.
//
// Reference layers: the same as conv_f16.cu (3x3 / dilated 3x3 with 64 input channels, 1x1 with Cin % 64 == 0).
// Layouts: split operand image [N][H][W][2*C] FP16 (hi planes in channels [0,C), lo planes in [C,2C)); split
// weight image 3x3: [9 taps][cout_pad][128] (hi | lo), 1x1: [2*Cin/64 chunks: hi..., lo...][cout_pad][64].
// Kernel structure: conv_f16.cu's (weights resident, one halo box per filter column and plane, warp-uniform MMA
// issue, 8 epilogue warps on 16x256b fragments); one stage per pipe, one pipe for dilations > 1 (shared memory:
// 144 KB of weights + 40-64 KB per stage).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace l3c {
namespace f16x2 {
using namespace tcx;

constexpr int TH = 8, TW = 16;
constexpr int W_TAP_BYTES = 64 * 128;
constexpr int W_PLANE_BYTES = 9 * W_TAP_BYTES;        // 73728 per plane (hi / lo)
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 32 * (4 + EPI_WARPS);         // warps 0,1 producers, 2,3 issuers (pipe 1 idle when n_pipes == 1)
constexpr int ACC_COLS = 128;                         // acc1 (64) | acc2 (64)
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;

struct Params {
    const float *bias;
    const float *residual;
    float *y;
    __half *yh;            // split image or null
    int N, H, W;
    int Cout, y_pitch, y_coff;
    int yh_pitch, yh_lo_off;
    int dilation;
    unsigned flags;
    int tiles_x, tiles_y;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// (hi, lo) pair images of two floats
__device__ __forceinline__ void split_h2(float a, float b, uint32_t &hi, uint32_t &lo) {
    hi = pack_h2(a, b);
    const float2 h = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    lo = pack_h2((a - h.x) * LO_SCALE, (b - h.y) * LO_SCALE);
}
__device__ __forceinline__ void quad_transpose(uint32_t (&v)[4], int tq) {
    const bool hi2 = (tq & 2) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint32_t send = hi2 ? v[i] : v[i + 2];
        const uint32_t got = __shfl_xor_sync(0xFFFFFFFFu, send, 2);
        if (hi2) v[i] = got; else v[i + 2] = got;
    }
    const bool hi1 = (tq & 1) != 0;
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        const uint32_t send = hi1 ? v[i] : v[i + 1];
        const uint32_t got = __shfl_xor_sync(0xFFFFFFFFu, send, 1);
        if (hi1) v[i] = got; else v[i + 1] = got;
    }
}

// barriers per pipe: full, empty, tfull[2], tempty[2]
constexpr int BARS_PER_PIPE = 6;

__global__ void __launch_bounds__(THREADS, 1)
conv3x3_f16x2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const Params p, const int n_pipes, const int a_bytes, const int ptiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // [weights hi 72 KB][weights lo 72 KB][pipe 0: hi box, lo box][pipe 1: ...][barriers]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * W_PLANE_BYTES + n_pipes * 2 * a_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * BARS_PER_PIPE + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t w_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    const uint32_t wbar = bar_base + 8u * (2 * BARS_PER_PIPE);
    const uint32_t tmem_cols = n_pipes == 2 ? 512u : 256u;

    if (threadIdx.x == 0) {
        for (int k = 0; k < 2; ++k) {
            const uint32_t b0 = bar_base + 8u * (BARS_PER_PIPE * k);
            mbar_init(b0, 1);                                              // full
            mbar_init(b0 + 8u, 1);                                         // empty
            for (int a = 0; a < 2; ++a) {
                mbar_init(b0 + 8u * (2 + a), 1);                           // tmem full
                mbar_init(b0 + 8u * (4 + a), EPI_WARPS);                   // tmem empty
            }
        }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int d = p.dilation;
    const int ct = blockIdx.y;
    const int lo_ch = 64;                                  // lo plane of a 64-channel operand image

    if (warp < 2) {
        // ===================== TMA producer of pipe `warp` =====================
        const int k = warp;
        if (k < n_pipes) {
            const uint32_t a_base = w_base + 2 * W_PLANE_BYTES + k * 2 * a_bytes;
            const uint32_t full = bar_base + 8u * (BARS_PER_PIPE * k), empty = full + 8u;
            if (k == 0 && lane == 0) {
                mbar_expect_tx(wbar, 2 * W_PLANE_BYTES);
                for (int tap = 0; tap < 9; ++tap) {
                    const int row = tap * (int)gridDim.y * 64 + ct * 64;
                    tma_load_2d(w_base + tap * W_TAP_BYTES, &map_w, wbar, 0, row);                      // hi
                    tma_load_2d(w_base + W_PLANE_BYTES + tap * W_TAP_BYTES, &map_w, wbar, 64, row);     // lo
                }
            }
            uint32_t phase = 0;
            for (int t = blockIdx.x + k * gridDim.x; t < ptiles; t += n_pipes * gridDim.x) {
                int q = t;
                const int tx = q % p.tiles_x; q /= p.tiles_x;
                const int ty = q % p.tiles_y; q /= p.tiles_y;
                const int n = q;
                for (int dx = 0; dx < 3; ++dx) {
                    if (lane == 0) {
                        mbar_wait(empty, phase ^ 1u);
                        mbar_expect_tx(full, 2 * a_bytes);
                        tma_load_4d(a_base, &map_x, full, 0, tx * TW + (dx - 1) * d, ty * TH - d, n);
                        tma_load_4d(a_base + a_bytes, &map_x, full, lo_ch, tx * TW + (dx - 1) * d, ty * TH - d, n);
                    }
                    __syncwarp();
                    phase ^= 1u;
                }
            }
        }
    } else if (warp < 4) {
        // ===================== MMA issuer of pipe `warp - 2` (warp-uniform code) =====================
        const int k = warp - 2;
        if (k < n_pipes) {
            const uint32_t a_base = w_base + 2 * W_PLANE_BYTES + k * 2 * a_bytes;
            const uint32_t full = bar_base + 8u * (BARS_PER_PIPE * k), empty = full + 8u;
            const uint32_t tfull0 = full + 16u, tempty0 = full + 32u;
            constexpr uint32_t IDESC = idesc_f16(64);
            uint32_t phase = 0, acc = 0, acc_phase = 0;
            mbar_wait(wbar, 0);
            for (int t = blockIdx.x + k * gridDim.x; t < ptiles; t += n_pipes * gridDim.x) {
                mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d1 = tmem_base + (k * 2 + acc) * ACC_COLS, d2 = d1 + 64u;
                for (int dx = 0; dx < 3; ++dx) {
                    mbar_wait(full, phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint64_t ah = make_desc(a_base + dy * d * (TW * 128));
                            const uint64_t al = make_desc(a_base + a_bytes + dy * d * (TW * 128));
                            const uint64_t bh = make_desc(w_base + (dy * 3 + dx) * W_TAP_BYTES);
                            const uint64_t bl = make_desc(w_base + W_PLANE_BYTES + (dy * 3 + dx) * W_TAP_BYTES);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint32_t first = (dx | dy | kk) != 0 ? 1u : 0u;
                                mma_f16(d1, ah + 2u * kk, bh + 2u * kk, IDESC, first);        // hi * hi
                                mma_f16(d2, ah + 2u * kk, bl + 2u * kk, IDESC, first);        // hi * lo
                                mma_f16(d2, al + 2u * kk, bh + 2u * kk, IDESC, 1u);           // lo * hi
                            }
                        }
                        mma_commit(empty);
                        if (dx == 2) mma_commit(tfull0 + 8u * acc);
                    }
                    __syncwarp();
                    phase ^= 1u;
                }
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===================== epilogue =====================
        const int e = warp - 4;
        const int quarter = warp & 3;
        const int cbeg = (e >> 2) * 32;
        const int tq = lane & 3, tr = lane >> 2;
        const bool relu = (p.flags & L3C_CONV_RELU) != 0;
        const bool shuffle = (p.flags & L3C_CONV_PIXEL_SHUFFLE2) != 0;
        const int cb = ct * 64 + cbeg;
        float2 bias2[4];
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) bias2[jb] = __ldg(reinterpret_cast<const float2 *>(p.bias + cb + 8 * jb + 2 * tq));
        int j = 0;
        for (int t = blockIdx.x; t < ptiles; t += gridDim.x, ++j) {
            const int k = j % n_pipes;
            const uint32_t acc = (uint32_t)(j / n_pipes) & 1u;
            const uint32_t acc_phase = (uint32_t)(j / (2 * n_pipes)) & 1u;
            const uint32_t tfull = bar_base + 8u * (BARS_PER_PIPE * k + 2 + acc), tempty = tfull + 16u;
            int q = t;
            const int tx = q % p.tiles_x; q /= p.tiles_x;
            const int ty = q % p.tiles_y; q /= p.tiles_y;
            const int n = q;
            size_t pix[2][2];
            bool ok[2][2];
#pragma unroll
            for (int lh = 0; lh < 2; ++lh)
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int oy = ty * TH + 2 * quarter + lh, ox = tx * TW + tr + 8 * rh;
                    ok[lh][rh] = (oy < p.H) && (ox < p.W);
                    pix[lh][rh] = ((size_t)n * p.H + oy) * p.W + ox;
                }
            float2 res[2][2][4];
            if (p.residual != nullptr && !shuffle) {
#pragma unroll
                for (int lh = 0; lh < 2; ++lh)
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                        for (int jb = 0; jb < 4; ++jb)
                            res[lh][rh][jb] = ok[lh][rh]
                                                  ? __ldg(reinterpret_cast<const float2 *>(
                                                        p.residual + pix[lh][rh] * p.y_pitch + p.y_coff + cb + 8 * jb + 2 * tq))
                                                  : make_float2(0.f, 0.f);
            }
            mbar_wait(tfull, acc_phase);
            tc_fence_after();
#pragma unroll
            for (int lh = 0; lh < 2; ++lh) {
                float v1[16], v2[16];
                const uint32_t taddr = tmem_base + (k * 2 + acc) * ACC_COLS + cbeg + ((uint32_t)(quarter * 32 + lh * 16) << 16);
                tmem_ld_16x256b_x4(taddr, v1);
                tmem_ld_16x256b_x4(taddr + 64u, v2);
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    float a[4], b[4];
#pragma unroll
                    for (int jb = 0; jb < 4; ++jb) {
                        a[jb] = __fmaf_rn(v2[4 * jb + 2 * rh], LO_INV, v1[4 * jb + 2 * rh]) + bias2[jb].x;
                        b[jb] = __fmaf_rn(v2[4 * jb + 2 * rh + 1], LO_INV, v1[4 * jb + 2 * rh + 1]) + bias2[jb].y;
                        if (relu) { a[jb] = fmaxf(a[jb], 0.f); b[jb] = fmaxf(b[jb], 0.f); }
                    }
                    const bool inside = ok[lh][rh];
                    if (!shuffle) {
                        if (p.residual != nullptr) {
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb) { a[jb] += res[lh][rh][jb].x; b[jb] += res[lh][rh][jb].y; }
                        }
                        if (p.y != nullptr && inside) {
                            const size_t off = pix[lh][rh] * p.y_pitch + p.y_coff + cb;
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb)
                                *reinterpret_cast<float2 *>(p.y + off + 8 * jb + 2 * tq) = make_float2(a[jb], b[jb]);
                        }
                        if (p.yh != nullptr) {
                            uint32_t h[4], l[4];
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb) split_h2(a[jb], b[jb], h[jb], l[jb]);
                            quad_transpose(h, tq);
                            quad_transpose(l, tq);
                            if (inside) {
                                const size_t offh = pix[lh][rh] * p.yh_pitch + p.y_coff + cb + 8 * tq;
                                *reinterpret_cast<uint4 *>(p.yh + offh) = make_uint4(h[0], h[1], h[2], h[3]);
                                *reinterpret_cast<uint4 *>(p.yh + offh + p.yh_lo_off) = make_uint4(l[0], l[1], l[2], l[3]);
                            }
                        }
                    } else {
                        // out[n, 2*oy+si, 2*ox+sj, cq] = conv[n, oy, ox, 4*cq + 2*si + sj]   (edsr.py:92-101)
                        const int oy = ty * TH + 2 * quarter + lh, ox = tx * TW + tr + 8 * rh;
                        const size_t W2 = 2 * (size_t)p.W;
                        const size_t row0 = ((size_t)n * (2 * p.H) + 2 * oy) * W2 + 2 * ox;
                        uint32_t fa[4], fb[4];
#pragma unroll
                        for (int jb = 0; jb < 4; ++jb) { fa[jb] = __float_as_uint(a[jb]); fb[jb] = __float_as_uint(b[jb]); }
                        quad_transpose(fa, tq);               // fa[s] = channel c0 + 2s, fb[s] = channel c0 + 2s + 1
                        quad_transpose(fb, tq);
                        if (inside) {
                            const int cq0 = (cb + 8 * tq) >> 2;
#pragma unroll
                            for (int si = 0; si < 2; ++si) {
                                const float x00 = __uint_as_float(fa[si]), x01 = __uint_as_float(fa[2 + si]);   // sj = 0: cq0, cq0+1
                                const float x10 = __uint_as_float(fb[si]), x11 = __uint_as_float(fb[2 + si]);   // sj = 1
                                if (p.y != nullptr) {
                                    const size_t o2 = (row0 + si * W2) * p.y_pitch + p.y_coff + cq0;
                                    *reinterpret_cast<float2 *>(p.y + o2) = make_float2(x00, x01);
                                    *reinterpret_cast<float2 *>(p.y + o2 + p.y_pitch) = make_float2(x10, x11);
                                }
                                if (p.yh != nullptr) {
                                    const size_t oh = (row0 + si * W2) * p.yh_pitch + p.y_coff + cq0;
                                    uint32_t h0, l0, h1, l1;
                                    split_h2(x00, x01, h0, l0);
                                    split_h2(x10, x11, h1, l1);
                                    *reinterpret_cast<uint32_t *>(p.yh + oh) = h0;
                                    *reinterpret_cast<uint32_t *>(p.yh + oh + p.yh_lo_off) = l0;
                                    *reinterpret_cast<uint32_t *>(p.yh + oh + p.yh_pitch) = h1;
                                    *reinterpret_cast<uint32_t *>(p.yh + oh + p.yh_pitch + p.yh_lo_off) = l1;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// 1x1 layers on split operands: tile = 128 pixels x all output channels; per 64-channel chunk of the input the
// stage holds the hi and the lo rows (2 x 16 KB); acc1 += hi*hi, acc2 += hi*lo + lo*hi with N = cout_pad.
// ---------------------------------------------------------------------------------------------
constexpr int K1_A_BYTES = 128 * 128;
constexpr int K1_EPI_WARPS = 8;
constexpr int K1_THREADS = 32 * (2 + K1_EPI_WARPS);
constexpr int K1_MAX_STAGES = 4;

struct Params1 {
    const float *bias;
    float *y;
    long long M;
    int Cout, y_pitch, y_coff;
    unsigned flags;
    int kchunks, npad, n_acc, tmem_cols, lo_ch;
};

__global__ void __launch_bounds__(K1_THREADS, 1)
conv1x1_f16x2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const Params1 p, const int n_stages, const int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int w_plane = p.kchunks * p.npad * 128;
    // [weights hi][weights lo][A ring: (hi 16 KB, lo 16 KB) per stage][barriers]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * w_plane + n_stages * 2 * K1_A_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * K1_MAX_STAGES + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t w_base = smem_u32(smem);
    const uint32_t a_base = w_base + 2 * w_plane;
    const uint32_t bar_base = smem_u32(bars);
    const uint32_t full0 = bar_base, empty0 = bar_base + 8u * K1_MAX_STAGES;
    const uint32_t tfull0 = bar_base + 8u * (2 * K1_MAX_STAGES), tempty0 = tfull0 + 16u;
    const uint32_t wbar = tfull0 + 32u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K1_MAX_STAGES; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8u * a, 1);
            mbar_init(tempty0 + 8u * a, K1_EPI_WARPS);
        }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, (uint32_t)(2 * w_plane));
            for (int kc = 0; kc < 2 * p.kchunks; ++kc)          // chunks 0..kchunks-1: hi, kchunks..: lo
                tma_load_2d(w_base + kc * p.npad * 128, &map_w, wbar, 0, kc * p.npad);
        }
        uint32_t stage = 0, phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (int kc = 0; kc < p.kchunks; ++kc) {
                if (lane == 0) {
                    mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                    mbar_expect_tx(full0 + 8u * stage, 2 * K1_A_BYTES);
                    tma_load_2d(a_base + stage * 2 * K1_A_BYTES, &map_x, full0 + 8u * stage, kc * 64, t * 128);
                    tma_load_2d(a_base + stage * 2 * K1_A_BYTES + K1_A_BYTES, &map_x, full0 + 8u * stage,
                                p.lo_ch + kc * 64, t * 128);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = idesc_f16((uint32_t)p.npad);
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0, acc_phase = 0;
        mbar_wait(wbar, 0);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d1 = tmem_base + acc * 2u * (uint32_t)p.npad, d2 = d1 + (uint32_t)p.npad;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(full0 + 8u * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ah = make_desc(a_base + stage * 2 * K1_A_BYTES);
                    const uint64_t al = make_desc(a_base + stage * 2 * K1_A_BYTES + K1_A_BYTES);
                    const uint64_t bh = make_desc(w_base + kc * p.npad * 128);
                    const uint64_t bl = make_desc(w_base + w_plane + kc * p.npad * 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t first = (kc | kk) != 0 ? 1u : 0u;
                        mma_f16(d1, ah + 2u * kk, bh + 2u * kk, idesc, first);
                        mma_f16(d2, ah + 2u * kk, bl + 2u * kk, idesc, first);
                        mma_f16(d2, al + 2u * kk, bh + 2u * kk, idesc, 1u);
                    }
                    mma_commit(empty0 + 8u * stage);
                    if (kc == p.kchunks - 1) mma_commit(tfull0 + 8u * acc);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
            if (p.n_acc == 2) {
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            } else {
                acc_phase ^= 1u;
            }
        }
    } else {
        const int e = warp - 2;
        const int quarter = warp & 3;
        const int half = e >> 2;
        const int ncol = p.npad >> 1;
        const int tq = lane & 3, tr = lane >> 2;
        const bool relu = (p.flags & L3C_CONV_RELU) != 0;
        uint32_t acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tfull0 + 8u * acc, acc_phase);
            tc_fence_after();
            for (int c0 = half * ncol; c0 < (half + 1) * ncol; c0 += 32) {
                if (c0 >= p.Cout) break;
#pragma unroll
                for (int lh = 0; lh < 2; ++lh) {
                    float v1[16], v2[16];
                    const uint32_t taddr = tmem_base + acc * 2u * (uint32_t)p.npad + c0 + ((uint32_t)(quarter * 32 + lh * 16) << 16);
                    tmem_ld_16x256b_x4(taddr, v1);
                    tmem_ld_16x256b_x4(taddr + (uint32_t)p.npad, v2);
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const long long pix = (long long)t * 128 + quarter * 32 + lh * 16 + tr + 8 * rh;
                        if (pix >= p.M) continue;
                        const size_t off = (size_t)pix * p.y_pitch + p.y_coff;
#pragma unroll
                        for (int jb = 0; jb < 4; ++jb) {
                            const int co = c0 + 8 * jb + 2 * tq;
                            if (co >= p.Cout) continue;
                            const float2 bb = __ldg(reinterpret_cast<const float2 *>(p.bias + co));
                            float a = __fmaf_rn(v2[4 * jb + 2 * rh], LO_INV, v1[4 * jb + 2 * rh]) + bb.x;
                            float b = __fmaf_rn(v2[4 * jb + 2 * rh + 1], LO_INV, v1[4 * jb + 2 * rh + 1]) + bb.y;
                            if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                            *reinterpret_cast<float2 *>(p.y + off + co) = make_float2(a, b);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
            if (p.n_acc == 2) {
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            } else {
                acc_phase ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// fp32 [n_px][C] -> split image [n_px][2*C]: for the activations that CUDA-core layers (5x5/s2, Cin = 3 / 5)
// produce in this mode and for callers that feed a tensor-core layer from outside the network
__global__ void split_kernel(const float4 *__restrict__ x, __half *__restrict__ out, long long n4, int C) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(x + i);
    const long long e = i * 4, px = e / C;
    const int c = (int)(e - px * C);
    uint32_t h0, l0, h1, l1;
    split_h2(v.x, v.y, h0, l0);
    split_h2(v.z, v.w, h1, l1);
    __half *o = out + px * 2 * C + c;
    *reinterpret_cast<uint2 *>(o) = make_uint2(h0, h1);
    *reinterpret_cast<uint2 *>(o + C) = make_uint2(l0, l1);
}

}  // namespace f16x2

extern "C" int l3c_split_f16x2(const float *x, long long n_px, int C, void *out, void *stream) {
    L3C_REQUIRE(x && out && n_px >= 0 && C >= 4 && C % 4 == 0, "l3c_split_f16x2: bad arguments (n_px=%lld C=%d)", n_px, C);
    const long long n4 = n_px * C / 4;
    if (n4 == 0) return L3C_OK;
    f16x2::split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(x), reinterpret_cast<__half *>(out), n4, C);
    L3C_LAUNCH_CHECK("split_f16x2_kernel");
    return L3C_OK;
}

// x_h: split operand image [N][H][W][x_pitch = 2*Cin] (hi | lo); w_h: split weight image (engine.PackedConv.get_f16x2)
int conv2d_f16x2(const l3c_conv_t &p, cudaStream_t st) {
    using namespace f16x2;
    const bool k3 = p.ksize == 3 && p.stride == 1 && p.Cin == 64 && p.x_pitch == 128;
    const bool k1 = p.ksize == 1 && p.stride == 1 && p.Cin % 64 == 0 && p.x_pitch == 2 * p.Cin &&
                    !(p.flags & L3C_CONV_PIXEL_SHUFFLE2);
    L3C_REQUIRE(k3 || k1, "l3c_conv2d[f16x2]: tensor-core path needs 3x3/s1/Cin=64 or 1x1/Cin%%64==0 on split operand "
                          "images (got k=%d s=%d Cin=%d pitch=%d)", p.ksize, p.stride, p.Cin, p.x_pitch);
    L3C_REQUIRE(p.x_h && p.w_h && (p.y || p.y_h), "l3c_conv2d[f16x2]: x_h / w_h and an output are required");
    L3C_REQUIRE(p.cout_pad % 64 == 0 && p.cout_pad >= p.Cout, "l3c_conv2d[f16x2]: cout_pad=%d", p.cout_pad);
    L3C_REQUIRE(!(p.flags & ~(L3C_CONV_RELU | L3C_CONV_PIXEL_SHUFFLE2)), "l3c_conv2d[f16x2]: unsupported flags %u", p.flags);
    EncodeTiledFn encode = get_encode_fn();
    L3C_REQUIRE(encode != nullptr, "l3c_conv2d: cuTensorMapEncodeTiled is not available from the driver");
    const int n_sm = stream_sm_count(st);
    static bool configured_dev[64] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        L3C_CUDA(cudaFuncSetAttribute(conv3x3_f16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        L3C_CUDA(cudaFuncSetAttribute(conv1x1_f16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    alignas(64) CUtensorMap map_x, map_w;
    if (k3) {
        L3C_REQUIRE(p.Cout % 64 == 0, "l3c_conv2d[f16x2]: 3x3 layers need Cout %% 64 == 0 (got %d)", p.Cout);
        const bool ps = (p.flags & L3C_CONV_PIXEL_SHUFFLE2) != 0;
        L3C_REQUIRE(!(ps && p.residual), "l3c_conv2d[f16x2]: residual with pixel shuffle is not built");
        L3C_REQUIRE(!p.y_h || (p.yh_pitch > 0 && p.yh_lo_off > 0 && p.yh_pitch % 8 == 0 && p.yh_lo_off % 8 == 0),
                    "l3c_conv2d[f16x2]: the split output image needs yh_pitch / yh_lo_off (multiples of 8)");
        L3C_REQUIRE(ps ? (p.y_pitch % 2 == 0 && p.y_coff % 2 == 0) : (p.y_pitch % 8 == 0 && p.y_coff % 8 == 0),
                    "l3c_conv2d[f16x2]: output pitch/offset alignment (pitch=%d coff=%d)", p.y_pitch, p.y_coff);
        const int d = p.dilation;
        const int a_rows = TH + 2 * d;
        const int a_bytes = a_rows * TW * 128;
        const int room = 227 * 1024 - 1024 - 512 - 2 * W_PLANE_BYTES;
        const int n_pipes = (2 * 2 * a_bytes <= room) ? 2 : 1;
        L3C_REQUIRE(2 * a_bytes <= room, "l3c_conv2d[f16x2]: halo of dilation %d does not fit in shared memory", d);
        {
            cuuint64_t dims[4] = {128, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
            cuuint64_t strides[3] = {256, (cuuint64_t)p.W * 256, (cuuint64_t)p.H * p.W * 256};
            cuuint32_t box[4] = {64, TW, (cuuint32_t)a_rows, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(p.x_h), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16x2]: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
        }
        {
            cuuint64_t dims[2] = {128, (cuuint64_t)9 * p.cout_pad};
            cuuint64_t strides[1] = {256};
            cuuint32_t box[2] = {64, 64};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.w_h), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16x2]: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
        }
        Params q;
        q.bias = p.bias; q.residual = p.residual; q.y = p.y; q.yh = reinterpret_cast<__half *>(p.y_h);
        q.N = p.N; q.H = p.H; q.W = p.W;
        q.Cout = p.Cout; q.y_pitch = p.y_pitch; q.y_coff = p.y_coff;
        q.yh_pitch = p.yh_pitch; q.yh_lo_off = p.yh_lo_off;
        q.dilation = d; q.flags = p.flags;
        q.tiles_x = ceil_div(p.W, TW);
        q.tiles_y = ceil_div(p.H, TH);
        const int cout_tiles = p.cout_pad / 64;
        const int ptiles = p.N * q.tiles_x * q.tiles_y;
        int per_ct = n_sm / cout_tiles;
        if (per_ct < 1) per_ct = 1;
        if (per_ct > ptiles) per_ct = ptiles;
        const int smem_bytes = 2 * W_PLANE_BYTES + n_pipes * 2 * a_bytes + 1024 + 512;
        conv3x3_f16x2_kernel<<<dim3(per_ct, cout_tiles), THREADS, smem_bytes, st>>>(map_x, map_w, q, n_pipes, a_bytes, ptiles);
        L3C_LAUNCH_CHECK("conv3x3_f16x2_kernel");
        return L3C_OK;
    }
    // ---- 1x1
    L3C_REQUIRE(p.cout_pad <= 256 && p.y && !p.y_h, "l3c_conv2d[f16x2]: 1x1 layers: Cout <= 256, fp32 output only");
    L3C_REQUIRE(p.Cout % 2 == 0 && p.y_pitch % 2 == 0 && p.y_coff % 2 == 0, "l3c_conv2d[f16x2]: 1x1 layers need even Cout/pitch/offset");
    L3C_REQUIRE(!p.residual, "l3c_conv2d[f16x2]: residual on a 1x1 layer is not built");
    const int kchunks = p.Cin / 64;
    const long long M = (long long)p.N * p.H * p.W;
    const int w_plane = kchunks * p.cout_pad * 128;
    int n_stages = (227 * 1024 - 1024 - 512 - 2 * w_plane) / (2 * K1_A_BYTES);
    if (n_stages > K1_MAX_STAGES) n_stages = K1_MAX_STAGES;
    L3C_REQUIRE(n_stages >= 1, "l3c_conv2d[f16x2]: weights of a %d -> %d 1x1 layer do not fit in shared memory", p.Cin, p.Cout);
    {
        cuuint64_t dims[2] = {(cuuint64_t)2 * p.Cin, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)p.x_pitch * 2};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.x_h), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16x2]: cuTensorMapEncodeTiled(x, 1x1) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)2 * kchunks * p.cout_pad};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)p.cout_pad};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.w_h), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16x2]: cuTensorMapEncodeTiled(w, 1x1) failed with %d", (int)r);
    }
    Params1 q;
    q.bias = p.bias; q.y = p.y;
    q.M = M; q.Cout = p.Cout; q.y_pitch = p.y_pitch; q.y_coff = p.y_coff; q.flags = p.flags;
    q.kchunks = kchunks; q.npad = p.cout_pad; q.lo_ch = p.Cin;
    q.n_acc = (4 * p.cout_pad <= 512) ? 2 : 1;
    q.tmem_cols = (q.n_acc * 2 * p.cout_pad <= 256) ? 256 : 512;
    const int n_tiles = (int)((M + 127) / 128);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    const int smem_bytes = 2 * w_plane + n_stages * 2 * K1_A_BYTES + 1024 + 512;
    conv1x1_f16x2_kernel<<<grid, K1_THREADS, smem_bytes, st>>>(map_x, map_w, q, n_stages, n_tiles);
    L3C_LAUNCH_CHECK("conv1x1_f16x2_kernel");
    return L3C_OK;
}

}  // namespace l3c
