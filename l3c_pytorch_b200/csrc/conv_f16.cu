// conv_f16.cu -- precision mode L3C_PREC_F16: the 64-input-channel 3x3 (optionally dilated) layers and the
// 1x1 layers of the L3C stack as implicit GEMMs on the Blackwell tensor cores with FP16 operands
// (tcgen05.mma kind::f16, M=128 pixels x N output channels x K=16) and fp32 accumulation in TMEM.
//
// Reference layers this replaces: every `default_conv(64, Cout, 3, rate=r)` of the stack
// (/root/reference/src/pytorch_ext.py:57-61; ResBlock edsr.py:63-89, body-final conv net.py:110,170,
// Head head.py:49-56, Upsampler edsr.py:92-101, atrous convs prob_clf.py:54-55) and the 1x1
// `lin` conv of the probability classifier (prob_clf.py:56,71-74) -- cuDNN fp32 in the reference.
//
// Why FP16 operands.  Round-to-nearest FP16 keeps the same 10 explicit mantissa bits as the TF32-RN
// operands of conv_tcgen05.cu (identical values for 6.1e-5 <= |x| <= 65504; saturating conversion above,
// absolute error < 3e-8 below), but
//   * one tcgen05.mma covers K = 16 instead of 8: half the MMA count and half the tensor-pipe time;
//   * an activation's operand image is 128 B per pixel (64 channels) = exactly ONE 128B-swizzle row, so a
//     halo box needs one TMA load per filter column (no K-halves) and half the L2->smem bytes;
//   * operand twins cost 128 instead of 256 B per pixel in HBM, and a layer whose output only feeds
//     tensor-core convs (first conv of a ResBlock, atrous convs, the finest decoder features) writes
//     ONLY the 2-byte image.
// With that the layer is HBM-bound on a B200: per pixel a 64->64 conv moves
// 256 B (operand in / operand out) to 768 B (operand + fp32 residual in, fp32 + operand out) against
// 73.7 kFLOP, i.e. 96-288 FLOP/B where the ridge is ~260.
//
// Structure of the 3x3 kernel (persistent, one CTA per SM, 384 threads):
//   * the 9 weight slabs of the CTA's Cout tile (9 x 64 rows x 128 B = 72 KB) stay in shared memory;
//   * per filter column dx ONE box of (8 + 2d) rows x 16 pixels x 64 channels is staged by TMA; the three
//     dy taps read it at row offsets dy*d*2048 B (multiples of the 1024 B swizzle atom);
//   * two independent (TMA producer warp, MMA issuer warp) pipes work on alternate tiles with their own
//     stages and TMEM accumulator pairs; the issuer code is warp-uniform and issues through elect.sync,
//     so descriptors live in uniform registers (2-3 SASS instructions per MMA instead of ~17);
//   * 8 epilogue warps drain TMEM (16x256b fragments), prefetch the fp32 residual BEFORE the accumulator
//     is complete, fuse bias / ReLU / residual / PixelShuffle(2), write fp32 float2 (full sectors) and the
//     FP16 image as 16-byte vectors after a 4x4 transpose inside each quad.
// Accumulation order per output element: (dx, dy, k-step) -- fixed, independent of batch / image size /
// tile position / pipe: encoder-side and decoder-side evaluations are bit-identical.
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "dmll_math.cuh"

namespace l3c {
namespace f16 {
using namespace tcx;

constexpr int TH = 8, TW = 16;                 // 128 output pixels = UMMA M
constexpr int W_TAP_BYTES = 64 * 128;          // one filter tap of one Cout tile: 64 rows x 64 fp16
constexpr int W_RES_BYTES = 9 * W_TAP_BYTES;   // 73728
constexpr int PIPES = 2;
constexpr int EPI_WARPS = 8;                   // two warps per TMEM lane quarter, 32 columns each
constexpr int THREADS = 32 * (2 * PIPES + EPI_WARPS);
constexpr int ACC_COLS = 64;
constexpr int TMEM_COLS = 256;                 // 2 pipes x 2 accumulators x 64 columns
constexpr int MAX_STAGES = 4;
constexpr int BARS_PER_PIPE = 2 * MAX_STAGES + 4;   // full[4], empty[4], tfull[2], tempty[2]

struct Params {
    const float *bias;      // padded to cout_pad
    const float *residual;  // fp32, layout of y, or null
    float *y;               // fp32 output or null
    __half *yh;             // fp16 operand image of the output or null
    int N, H, W;
    int Cout, y_pitch, y_coff;
    int dilation;
    unsigned flags;
    int tiles_x, tiles_y;
};

// two floats -> packed fp16 pair, round-to-nearest-even, saturating to +-65504
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // d.hi = first source
    return r;
}

// 4x4 transpose inside each quad of lanes: before, thread tq holds v[j] = element (column block j, pair tq);
// after, thread tq holds v[s] = element (column block tq, pair s) -> 8 consecutive channels = 16 bytes.
// Must be executed by all 32 lanes.
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

// KIND 0: 3x3 stride 1 (any dilation), weights resident.
// KIND 1: 5x5 stride 2 (the down-sampling conv of every encoder, net.py:101): the input box of a tap is read
//   with TMA element strides (2, 2), so the A operand of an 8x16 OUTPUT tile is again 128 dense rows of 128 B;
//   for a filter column kx the taps ky = 0,2,4 (and ky = 1,3) share one strided box of 10 rows read at row
//   offsets j*2048 B.  25 taps x 8 KB of weights do not fit beside the stages: every unit (kx, row parity)
//   brings its 3 (2) weight slabs along with its box.  10 units, 100 MMAs per tile.
constexpr int K5_A_BYTES = 10 * TW * 128;                  // 20480: strided halo box
constexpr int K5_STAGE_BYTES = K5_A_BYTES + 3 * W_TAP_BYTES;   // 45056: box + up to three weight slabs

template <int KIND>
__global__ void __launch_bounds__(THREADS, 1)
conv_f16_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                const Params p, const int n_stages, const int a_bytes, const int ptiles) {
    constexpr int W_RES = (KIND == 0) ? W_RES_BYTES : 0;    // resident weights
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // [weights 72 KB][pipe 0 stages][pipe 1 stages][barriers]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + W_RES + PIPES * n_stages * a_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + PIPES * BARS_PER_PIPE + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t w_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    const uint32_t wbar = bar_base + 8u * (PIPES * BARS_PER_PIPE);

    if (threadIdx.x == 0) {
        for (int k = 0; k < PIPES; ++k) {
            const uint32_t b0 = bar_base + 8u * (BARS_PER_PIPE * k);
            for (int s = 0; s < MAX_STAGES; ++s) {
                mbar_init(b0 + 8u * s, 1);                                 // full
                mbar_init(b0 + 8u * (MAX_STAGES + s), 1);                  // empty
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(b0 + 8u * (2 * MAX_STAGES + a), 1);              // tmem full
                mbar_init(b0 + 8u * (2 * MAX_STAGES + 2 + a), EPI_WARPS);  // tmem empty: one arrival per epilogue warp
            }
        }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == PIPES) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int d = p.dilation;
    const int ct = blockIdx.y;                        // this CTA's Cout tile (weights stay resident)

    if (warp < PIPES) {
        // ===================== TMA producer of pipe `warp` =====================
        const int k = warp;
        const uint32_t a_base = w_base + W_RES + k * n_stages * a_bytes;
        const uint32_t full0 = bar_base + 8u * (BARS_PER_PIPE * k), empty0 = full0 + 8u * MAX_STAGES;
        if (KIND == 0 && k == 0 && lane == 0) {
            mbar_expect_tx(wbar, W_RES_BYTES);
            for (int tap = 0; tap < 9; ++tap)         // rows [tap][cout_pad] of the weight image
                tma_load_2d(w_base + tap * W_TAP_BYTES, &map_w, wbar, 0, tap * (int)gridDim.y * 64 + ct * 64);
        }
        uint32_t stage = 0, phase = 0;
        for (int t = blockIdx.x + k * gridDim.x; t < ptiles; t += PIPES * gridDim.x) {
            int q = t;
            const int tx = q % p.tiles_x; q /= p.tiles_x;
            const int ty = q % p.tiles_y; q /= p.tiles_y;
            const int n = q;
            constexpr int UNITS = (KIND == 0) ? 3 : 10;
            for (int u = 0; u < UNITS; ++u) {
                if (lane == 0) {
                    mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                    const uint32_t dst = a_base + stage * a_bytes, bar = full0 + 8u * stage;
                    if (KIND == 0) {
                        mbar_expect_tx(bar, a_bytes);
                        tma_load_4d(dst, &map_x, bar, 0, tx * TW + (u - 1) * d, ty * TH - d, n);
                    } else {
                        const int kx = u >> 1, par = u & 1, ntaps = 3 - par;
                        mbar_expect_tx(bar, K5_A_BYTES + ntaps * W_TAP_BYTES);
                        // input pixel of output (oy, ox), tap (ky, kx): (2*oy + ky - 2, 2*ox + kx - 2)
                        tma_load_4d(dst, &map_x, bar, 0, 2 * tx * TW + kx - 2, 2 * ty * TH + par - 2, n);
                        for (int j = 0; j < ntaps; ++j)
                            tma_load_2d(dst + K5_A_BYTES + j * W_TAP_BYTES, &map_w, bar, 0,
                                        ((par + 2 * j) * 5 + kx) * (int)gridDim.y * 64 + ct * 64);
                    }
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp < 2 * PIPES) {
        // ===================== MMA issuer of pipe `warp - PIPES` (warp-uniform code) =====================
        const int k = warp - PIPES;
        const uint32_t a_base = w_base + W_RES + k * n_stages * a_bytes;
        const uint32_t full0 = bar_base + 8u * (BARS_PER_PIPE * k), empty0 = full0 + 8u * MAX_STAGES;
        const uint32_t tfull0 = full0 + 8u * (2 * MAX_STAGES), tempty0 = tfull0 + 16u;
        constexpr uint32_t IDESC = idesc_f16(64);
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0, acc_phase = 0;
        if (KIND == 0) mbar_wait(wbar, 0);
        for (int t = blockIdx.x + k * gridDim.x; t < ptiles; t += PIPES * gridDim.x) {
            mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1u);       // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (k * 2 + acc) * ACC_COLS;
            constexpr int UNITS = (KIND == 0) ? 3 : 10;
            for (int u = 0; u < UNITS; ++u) {
                mbar_wait(full0 + 8u * stage, phase);
                tc_fence_after();
                const uint32_t a0 = a_base + stage * a_bytes;
                if (elect_one()) {
                    if (KIND == 0) {
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint64_t da = make_desc(a0 + dy * d * (TW * 128));           // dy*d rows of 16 px
                            const uint64_t db = make_desc(w_base + (dy * 3 + u) * W_TAP_BYTES);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)                                     // +32 B = 16 fp16 along K
                                mma_f16(d_tmem, da + 2u * kk, db + 2u * kk, IDESC, (u | dy | kk) != 0 ? 1u : 0u);
                        }
                    } else {
                        const int ntaps = 3 - (u & 1);
                        for (int j = 0; j < ntaps; ++j) {
                            const uint64_t da = make_desc(a0 + j * (TW * 128));                // box row j = tap ky0 + 2j
                            const uint64_t db = make_desc(a0 + K5_A_BYTES + j * W_TAP_BYTES);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                mma_f16(d_tmem, da + 2u * kk, db + 2u * kk, IDESC, (u | j | kk) != 0 ? 1u : 0u);
                        }
                    }
                    mma_commit(empty0 + 8u * stage);              // frees the stage when the MMAs retire
                    if (u == UNITS - 1) mma_commit(tfull0 + 8u * acc);   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ===================== epilogue: tiles in sequence order, alternating pipes =====================
        const int e = warp - 2 * PIPES;
        const int quarter = warp & 3;                         // TMEM lane quarter = warp id % 4 (hardware rule)
        const int cbeg = (e >> 2) * 32;                       // this warp's 32 accumulator columns
        const int tq = lane & 3, tr = lane >> 2;              // column pair / row inside the fragment
        const bool relu = (p.flags & L3C_CONV_RELU) != 0;
        const bool shuffle = (p.flags & L3C_CONV_PIXEL_SHUFFLE2) != 0;
        const int cb = ct * 64 + cbeg;                        // first conv channel of this warp
        float2 bias2[4];                                      // bias of channels cb + 8j + 2*tq + {0,1}
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) bias2[jb] = __ldg(reinterpret_cast<const float2 *>(p.bias + cb + 8 * jb + 2 * tq));
        int j = 0;
        for (int t = blockIdx.x; t < ptiles; t += gridDim.x, ++j) {
            const int k = j % PIPES;
            const uint32_t acc = (uint32_t)(j / PIPES) & 1u;
            const uint32_t acc_phase = (uint32_t)(j / (2 * PIPES)) & 1u;
            const uint32_t tfull = bar_base + 8u * (BARS_PER_PIPE * k + 2 * MAX_STAGES + acc), tempty = tfull + 16u;
            int q = t;
            const int tx = q % p.tiles_x; q /= p.tiles_x;
            const int ty = q % p.tiles_y; q /= p.tiles_y;
            const int n = q;
            // fragment rows: lh -> tile row 2*quarter + lh, rh -> pixel tr + 8*rh of that row
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
            // the residual does not depend on the accumulator: fetch it while the MMAs are still running
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
                float v[16];
                tmem_ld_16x256b_x4(tmem_base + (k * 2 + acc) * ACC_COLS + cbeg + ((uint32_t)(quarter * 32 + lh * 16) << 16), v);
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    float a[4], b[4];
#pragma unroll
                    for (int jb = 0; jb < 4; ++jb) {
                        a[jb] = v[4 * jb + 2 * rh] + bias2[jb].x;
                        b[jb] = v[4 * jb + 2 * rh + 1] + bias2[jb].y;
                        if (relu) { a[jb] = fmaxf(a[jb], 0.f); b[jb] = fmaxf(b[jb], 0.f); }
                    }
                    const bool inside = ok[lh][rh];
                    if (!shuffle) {
                        const size_t off = pix[lh][rh] * p.y_pitch + p.y_coff + cb;
                        if (p.residual != nullptr) {
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb) { a[jb] += res[lh][rh][jb].x; b[jb] += res[lh][rh][jb].y; }
                        }
                        if (p.y != nullptr && inside) {
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb)
                                *reinterpret_cast<float2 *>(p.y + off + 8 * jb + 2 * tq) = make_float2(a[jb], b[jb]);
                        }
                        if (p.yh != nullptr) {
                            uint32_t h[4];
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb) h[jb] = pack_h2(a[jb], b[jb]);
                            quad_transpose(h, tq);            // -> channels cb + 8*tq + 0..7
                            if (inside) *reinterpret_cast<uint4 *>(p.yh + off + 8 * tq) = make_uint4(h[0], h[1], h[2], h[3]);
                        }
                    } else {
                        // out[n, 2*oy+si, 2*ox+sj, cq] = conv[n, oy, ox, 4*cq + 2*si + sj]   (edsr.py:92-101)
                        const int oy = ty * TH + 2 * quarter + lh, ox = tx * TW + tr + 8 * rh;
                        const size_t W2 = 2 * (size_t)p.W;
                        const size_t row0 = ((size_t)n * (2 * p.H) + 2 * oy) * W2 + 2 * ox;   // output pixel (si=0, sj=0)
                        if (p.y != nullptr) {
                            // same transpose on the fp32 values: thread tq then owns conv channels c0 + 0..7
                            // (c0 = cb + 8*tq) = output channels cq0, cq0 + 1 of the four sub-pixels, and a quad
                            // writes 32 contiguous bytes per output pixel
                            uint32_t fa[4], fb[4];
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb) { fa[jb] = __float_as_uint(a[jb]); fb[jb] = __float_as_uint(b[jb]); }
                            quad_transpose(fa, tq);           // fa[s] = channel c0 + 2s, fb[s] = channel c0 + 2s + 1
                            quad_transpose(fb, tq);
                            if (inside) {
                                const int cq0 = (cb + 8 * tq) >> 2;
#pragma unroll
                                for (int si = 0; si < 2; ++si) {
                                    const size_t o2 = (row0 + si * W2) * p.y_pitch + p.y_coff + cq0;
                                    *reinterpret_cast<float2 *>(p.y + o2) =
                                        make_float2(__uint_as_float(fa[si]), __uint_as_float(fa[2 + si]));        // sj = 0
                                    *reinterpret_cast<float2 *>(p.y + o2 + p.y_pitch) =
                                        make_float2(__uint_as_float(fb[si]), __uint_as_float(fb[2 + si]));        // sj = 1
                                }
                            }
                        }
                        if (p.yh != nullptr) {
                            uint32_t h[4];
#pragma unroll
                            for (int jb = 0; jb < 4; ++jb) h[jb] = pack_h2(a[jb], b[jb]);
                            quad_transpose(h, tq);            // h[s] = conv channels c0 + 2s + {0,1}, c0 = cb + 8*tq
                            if (inside) {
                                const int cq0 = (cb + 8 * tq) >> 2;          // output channels cq0, cq0 + 1
#pragma unroll
                                for (int si = 0; si < 2; ++si) {
                                    const size_t o2 = (row0 + si * W2) * p.y_pitch + p.y_coff + cq0;
                                    // sj = 0: low halves of h[si] (cq0) and h[2+si] (cq0+1); sj = 1: the high halves
                                    *reinterpret_cast<uint32_t *>(p.yh + o2) = __byte_perm(h[si], h[2 + si], 0x5410);
                                    *reinterpret_cast<uint32_t *>(p.yh + o2 + p.y_pitch) = __byte_perm(h[si], h[2 + si], 0x7632);
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
    if (warp == PIPES) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// 1x1 layers (Cin % 64 == 0, e.g. the 192 -> Kp `lin` conv of the probability classifier): a plain GEMM
// [pixels][Cin] x [Cin][Cout].  Tile = 128 consecutive pixels x ALL output channels (one MMA of
// N = cout_pad <= 256 per 16 input channels), so every activation row is read from HBM exactly once;
// the weights (Cin/64 chunks x cout_pad rows x 128 B) stay in shared memory.  HBM-bound: per pixel
// 2*Cin bytes in, 4*Cout bytes out.
// ---------------------------------------------------------------------------------------------
constexpr int K1_A_BYTES = 128 * 128;          // one K-chunk (64 channels) of a 128-pixel tile
constexpr int K1_EPI_WARPS = 8;
constexpr int K1_THREADS = 32 * (2 + K1_EPI_WARPS);
constexpr int K1_MAX_STAGES = 8;

struct Params1 {
    const float *bias;
    float *y;
    __half *yh;
    long long M;            // pixels
    int Cout, y_pitch, y_coff;
    unsigned flags;
    int kchunks, npad;      // Cin / 64, cout_pad
    int tmem_cols;
};

__global__ void __launch_bounds__(K1_THREADS, 1)
conv1x1_f16_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const Params1 p, const int n_stages, const int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int w_bytes = p.kchunks * p.npad * 128;
    // [weights][A ring][barriers: full[8], empty[8], tfull[2], tempty[2], weights]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + w_bytes + n_stages * K1_A_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * K1_MAX_STAGES + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t w_base = smem_u32(smem);
    const uint32_t a_base = w_base + w_bytes;
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(wbar, (uint32_t)w_bytes);
            for (int kc = 0; kc < p.kchunks; ++kc)
                tma_load_2d(w_base + kc * p.npad * 128, &map_w, wbar, 0, kc * p.npad);
        }
        uint32_t stage = 0, phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (int kc = 0; kc < p.kchunks; ++kc) {
                if (lane == 0) {
                    mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                    mbar_expect_tx(full0 + 8u * stage, K1_A_BYTES);
                    tma_load_2d(a_base + stage * K1_A_BYTES, &map_x, full0 + 8u * stage, kc * 64, t * 128);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform code) =====================
        const uint32_t idesc = idesc_f16((uint32_t)p.npad);
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0, acc_phase = 0;
        mbar_wait(wbar, 0);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.npad;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(full0 + 8u * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_desc(a_base + stage * K1_A_BYTES);
                    const uint64_t db = make_desc(w_base + kc * p.npad * 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        mma_f16(d_tmem, da + 2u * kk, db + 2u * kk, idesc, (kc | kk) != 0 ? 1u : 0u);
                    mma_commit(empty0 + 8u * stage);
                    if (kc == p.kchunks - 1) mma_commit(tfull0 + 8u * acc);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ===================== epilogue =====================
        const int e = warp - 2;
        const int quarter = warp & 3;
        const int half = e >> 2;
        const int ncol = p.npad >> 1;                         // columns of this warp: [half*ncol, (half+1)*ncol)
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
                    float v[16];
                    tmem_ld_16x256b_x4(tmem_base + acc * (uint32_t)p.npad + c0 + ((uint32_t)(quarter * 32 + lh * 16) << 16), v);
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const long long pix = (long long)t * 128 + quarter * 32 + lh * 16 + tr + 8 * rh;
                        if (pix >= p.M) continue;
                        const size_t off = (size_t)pix * p.y_pitch + p.y_coff;
#pragma unroll
                        for (int jb = 0; jb < 4; ++jb) {
                            const int co = c0 + 8 * jb + 2 * tq;
                            if (co >= p.Cout) continue;                       // Cout is even: co + 1 < Cout too
                            const float2 bb = __ldg(reinterpret_cast<const float2 *>(p.bias + co));
                            float a = v[4 * jb + 2 * rh] + bb.x, b = v[4 * jb + 2 * rh + 1] + bb.y;
                            if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                            if (p.y != nullptr) *reinterpret_cast<float2 *>(p.y + off + co) = make_float2(a, b);
                            if (p.yh != nullptr) *reinterpret_cast<uint32_t *>(p.yh + off + co) = pack_h2(a, b);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// The DMLL head fused into the 1x1 `lin` conv of the probability classifier (encode side): the GEMM of
// conv1x1_f16_kernel, but the epilogue never writes the Kp parameters of a pixel (480 / 600 B) to HBM -- each
// epilogue thread owns one pixel (TMEM lane), pulls the K logits / means / log-scales (and RGB coefficients) of
// one channel at a time out of the accumulator, and evaluates the two CDF bounds of the symbol being coded
// (dmll_math.cuh: the SAME functions the decoder's row builder uses on the parameters it computes with the
// plain kernel, so both sides see identical integers).  Output: 4 B per coded sub-pixel.
// Reference: criterion/logistic_mixture.py:248-275 + torchac_kernel.cu:26-76 after prob_clf.py:71-74.
// ---------------------------------------------------------------------------------------------
constexpr int LD_EPI_WARPS = 8;            // two groups of four (one warp per TMEM lane quarter), alternating tiles
constexpr int LD_THREADS = 32 * (2 + LD_EPI_WARPS);

struct ParamsLD {
    const float *bias;
    const uint8_t *sym;         // [N][C][HW] symbols being coded
    const float *targets;       // [L + 1] bin edges
    uint32_t *intervals;        // [N][C][HW] out
    long long M;                // pixels = N * HW
    int HW, L;
    int kchunks, npad, tmem_cols;
};

__device__ __forceinline__ void tmem_ld10_issue(uint32_t taddr, float (&v)[10]) {
    uint32_t r[10];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[8]), "=r"(r[9]) : "r"(taddr + 8u));
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = __uint_as_float(r[i]);
}
// tcgen05.wait::ld tied to the registers it makes valid (the compiler must not use them earlier)
__device__ __forceinline__ void tmem_ld10_wait(float (&v)[10]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                   "+f"(v[8]), "+f"(v[9])
                 :
                 : "memory");
}

template <int C, bool RGB>
__global__ void __launch_bounds__(LD_THREADS, 1)
lin_dmll_f16_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const ParamsLD p, const int n_stages, const int n_tiles) {
    constexpr int K = 10;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int w_bytes = p.kchunks * p.npad * 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + w_bytes + n_stages * K1_A_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * K1_MAX_STAGES + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t w_base = smem_u32(smem);
    const uint32_t a_base = w_base + w_bytes;
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
            mbar_init(tempty0 + 8u * a, 4);               // drained by ONE group of four epilogue warps
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(wbar, (uint32_t)w_bytes);
            for (int kc = 0; kc < p.kchunks; ++kc)
                tma_load_2d(w_base + kc * p.npad * 128, &map_w, wbar, 0, kc * p.npad);
        }
        uint32_t stage = 0, phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (int kc = 0; kc < p.kchunks; ++kc) {
                if (lane == 0) {
                    mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                    mbar_expect_tx(full0 + 8u * stage, K1_A_BYTES);
                    tma_load_2d(a_base + stage * K1_A_BYTES, &map_x, full0 + 8u * stage, kc * 64, t * 128);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform code) =====================
        const uint32_t idesc = idesc_f16((uint32_t)p.npad);
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0, acc_phase = 0;
        mbar_wait(wbar, 0);
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.npad;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(full0 + 8u * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_desc(a_base + stage * K1_A_BYTES);
                    const uint64_t db = make_desc(w_base + kc * p.npad * 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        mma_f16(d_tmem, da + 2u * kk, db + 2u * kk, idesc, (kc | kk) != 0 ? 1u : 0u);
                    mma_commit(empty0 + 8u * stage);
                    if (kc == p.kchunks - 1) mma_commit(tfull0 + 8u * acc);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ===================== epilogue: thread = pixel (TMEM lane 32*quarter + lane) =====================
        // The per-pixel DMLL arithmetic is long dependent chains (exp, division, sigmoids): four warps cannot
        // keep the SM's schedulers busy, so two groups of four take alternate tiles -- group g always drains
        // accumulator g (the issuer alternates accumulators per tile).
        const int quarter = warp & 3;
        const int grp = (warp - 2) >> 2;
        const float scale = (float)(65536 - p.L);            // 2^16 - (Lp - 1)
        const uint32_t acc = (uint32_t)grp;
        uint32_t acc_phase = 0;
        int seq = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++seq) {
            if ((seq & 1) != grp) continue;
            const long long pix = (long long)t * 128 + quarter * 32 + lane;
            const bool valid = pix < p.M;
            const int n = valid ? (int)(pix / p.HW) : 0;
            const int q = valid ? (int)(pix % p.HW) : 0;
            mbar_wait(tfull0 + 8u * acc, acc_phase);
            tc_fence_after();
            const uint32_t trow = tmem_base + acc * (uint32_t)p.npad + ((uint32_t)(quarter * 32) << 16);
            float xr = 0.f, xg = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float logit[K], mean[K], logs[K], co0[K], co1[K];
                tmem_ld10_issue(trow + 0 * C * K + c * K, logit);
                tmem_ld10_issue(trow + 1 * C * K + c * K, mean);
                tmem_ld10_issue(trow + 2 * C * K + c * K, logs);
                if (RGB && c == 1) tmem_ld10_issue(trow + 3 * C * K + 0 * K, co0);
                if (RGB && c == 2) {
                    tmem_ld10_issue(trow + 3 * C * K + 1 * K, co0);
                    tmem_ld10_issue(trow + 3 * C * K + 2 * K, co1);
                }
                tmem_ld10_wait(logit);
                tmem_ld10_wait(mean);
                tmem_ld10_wait(logs);
#pragma unroll
                for (int k = 0; k < K; ++k) {                // + bias (same add as the plain kernel's epilogue)
                    logit[k] += __ldg(p.bias + 0 * C * K + c * K + k);
                    mean[k] += __ldg(p.bias + 1 * C * K + c * K + k);
                    logs[k] += __ldg(p.bias + 2 * C * K + c * K + k);
                }
                if (RGB && c == 1) {
                    tmem_ld10_wait(co0);
#pragma unroll
                    for (int k = 0; k < K; ++k) { co0[k] += __ldg(p.bias + 3 * C * K + 0 * K + k); co1[k] = 0.f; }
                } else if (RGB && c == 2) {
                    tmem_ld10_wait(co0);
                    tmem_ld10_wait(co1);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        co0[k] += __ldg(p.bias + 3 * C * K + 1 * K + k);
                        co1[k] += __ldg(p.bias + 3 * C * K + 2 * K + k);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) co0[k] = co1[k] = 0.f;
                }
                ChanParams<K> cp;
                channel_params_core<K>(logit, mean, logs, co0, co1, c, RGB, xr, xg, cp);
                const size_t so = ((size_t)n * C + c) * p.HW + q;
                const int s = valid ? (int)p.sym[so] : 0;
                const uint32_t lo = mixture_cdf_u16<K>(cp.pi, cp.mu, cp.inv_s, __ldg(p.targets + s), scale, s);
                const uint32_t hi = (s == p.L - 1) ? 0x10000u
                                                   : mixture_cdf_u16<K>(cp.pi, cp.mu, cp.inv_s, __ldg(p.targets + s + 1),
                                                                        scale, s + 1);
                if (valid) p.intervals[so] = lo | ((hi - 1u) << 16);
                if (c == 0) xr = (float)s;
                if (c == 1) xg = (float)s;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
            acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace f16

// dynamic shared memory a conv CTA may take (bytes).  Default: everything an SM has (227 KB).  L3C_CONV_SMEM_KB
// lowers it so that small latency-bound CTAs (the range decoders: 16 KB per stream) can be resident BESIDE a
// conv CTA when no SM partition separates them (codec.lanes, L3C_SM_PARTITION=0).
static int conv_smem_cap() {
    static int cap = 0;
    if (cap == 0) {
        const char *e = getenv("L3C_CONV_SMEM_KB");
        int kb = e ? atoi(e) : 227;
        if (kb < 112) kb = 112;
        if (kb > 227) kb = 227;
        cap = kb * 1024;
    }
    return cap;
}

// operand image x_h: fp16 NHWC [N][H][W][x_pitch]; weight image w_h (engine.PackedConv.get_f16):
//   3x3: [tap 9][cout_pad][64] fp16      1x1: [Cin/64][cout_pad][64] fp16     (rows of 128 B)
int conv2d_f16(const l3c_conv_t &p, cudaStream_t st) {
    using namespace f16;
    const bool k3 = p.ksize == 3 && p.stride == 1 && p.Cin == 64 && p.x_pitch == 64;
    const bool k5 = p.ksize == 5 && p.stride == 2 && p.dilation == 1 && p.Cin == 64 && p.x_pitch == 64;
    const bool k1 = p.ksize == 1 && p.stride == 1 && p.Cin % 64 == 0 && p.x_pitch == p.Cin &&
                    !(p.flags & L3C_CONV_PIXEL_SHUFFLE2);
    L3C_REQUIRE(k3 || k5 || k1, "l3c_conv2d[f16]: tensor-core path needs 3x3/s1/Cin=64, 5x5/s2/Cin=64 or 1x1/Cin%%64==0, "
                                "dense input pitch (got k=%d s=%d Cin=%d pitch=%d)", p.ksize, p.stride, p.Cin, p.x_pitch);
    L3C_REQUIRE(p.x_h && p.w_h, "l3c_conv2d[f16]: x_h / w_h (fp16 operand images) are required");
    L3C_REQUIRE(p.y || p.y_h, "l3c_conv2d[f16]: no output");
    L3C_REQUIRE(p.cout_pad % 64 == 0 && p.cout_pad >= p.Cout, "l3c_conv2d[f16]: cout_pad=%d", p.cout_pad);
    L3C_REQUIRE(!(p.flags & ~(L3C_CONV_RELU | L3C_CONV_PIXEL_SHUFFLE2)), "l3c_conv2d[f16]: unsupported flags %u", p.flags);
    EncodeTiledFn encode = get_encode_fn();
    L3C_REQUIRE(encode != nullptr, "l3c_conv2d: cuTensorMapEncodeTiled is not available from the driver");
    const int n_sm = stream_sm_count(st);          // the stream may be confined to a group of SMs
    static bool configured_dev[64] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        L3C_CUDA(cudaFuncSetAttribute(conv_f16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        L3C_CUDA(cudaFuncSetAttribute(conv_f16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        L3C_CUDA(cudaFuncSetAttribute(conv1x1_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    alignas(64) CUtensorMap map_x, map_w;
    if (k3) {
        L3C_REQUIRE(p.Cout % 64 == 0, "l3c_conv2d[f16]: 3x3 layers need Cout %% 64 == 0 (got %d)", p.Cout);
        const bool ps = (p.flags & L3C_CONV_PIXEL_SHUFFLE2) != 0;
        L3C_REQUIRE(ps ? (p.y_pitch % 2 == 0 && p.y_coff % 2 == 0) : (p.y_pitch % 8 == 0 && p.y_coff % 8 == 0),
                    "l3c_conv2d[f16]: output pitch/offset alignment (pitch=%d coff=%d)", p.y_pitch, p.y_coff);
        L3C_REQUIRE(!(ps && p.residual), "l3c_conv2d[f16]: residual with pixel shuffle is not built");
        const int d = p.dilation;
        const int a_rows = TH + 2 * d;
        L3C_REQUIRE(a_rows <= 256, "l3c_conv2d[f16]: dilation %d too large", d);
        const int a_bytes = a_rows * TW * 128;
        int n_stages = (conv_smem_cap() - 1024 - 512 - W_RES_BYTES) / (PIPES * a_bytes);      // per pipe
        if (n_stages > 3) n_stages = 3;
        if (n_stages < 1 && PIPES * a_bytes <= 227 * 1024 - 1024 - 512 - W_RES_BYTES) n_stages = 1;   // above the cap
        L3C_REQUIRE(n_stages >= 1, "l3c_conv2d[f16]: halo of dilation %d does not fit in shared memory", d);
        {
            cuuint64_t dims[4] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
            cuuint64_t strides[3] = {128, (cuuint64_t)p.W * 128, (cuuint64_t)p.H * p.W * 128};
            cuuint32_t box[4] = {64, TW, (cuuint32_t)a_rows, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(p.x_h), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16]: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
        }
        {
            cuuint64_t dims[2] = {64, (cuuint64_t)9 * p.cout_pad};
            cuuint64_t strides[1] = {128};
            cuuint32_t box[2] = {64, 64};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.w_h), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16]: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
        }
        Params q;
        q.bias = p.bias; q.residual = p.residual; q.y = p.y; q.yh = reinterpret_cast<__half *>(p.y_h);
        q.N = p.N; q.H = p.H; q.W = p.W;
        q.Cout = p.Cout; q.y_pitch = p.y_pitch; q.y_coff = p.y_coff;
        q.dilation = d; q.flags = p.flags;
        q.tiles_x = ceil_div(p.W, TW);
        q.tiles_y = ceil_div(p.H, TH);
        const int cout_tiles = p.cout_pad / 64;
        const int ptiles = p.N * q.tiles_x * q.tiles_y;
        int per_ct = n_sm / cout_tiles;
        if (per_ct < 1) per_ct = 1;
        if (per_ct > ptiles) per_ct = ptiles;
        const int smem_bytes = W_RES_BYTES + PIPES * n_stages * a_bytes + 1024 + 512;
        conv_f16_kernel<0><<<dim3(per_ct, cout_tiles), THREADS, smem_bytes, st>>>(map_x, map_w, q, n_stages, a_bytes, ptiles);
        L3C_LAUNCH_CHECK("conv3x3_f16_kernel");
        return L3C_OK;
    }
    if (k5) {
        L3C_REQUIRE(p.Cout % 64 == 0, "l3c_conv2d[f16]: 5x5/s2 layers need Cout %% 64 == 0 (got %d)", p.Cout);
        L3C_REQUIRE(!(p.flags & L3C_CONV_PIXEL_SHUFFLE2) && p.y_pitch % 8 == 0 && p.y_coff % 8 == 0,
                    "l3c_conv2d[f16]: 5x5/s2 output pitch/offset alignment (pitch=%d coff=%d)", p.y_pitch, p.y_coff);
        const int Ho = (p.H + 4 - 5) / 2 + 1, Wo = (p.W + 4 - 5) / 2 + 1;
        {
            // element strides (2, 2): the box visits every second pixel / row; box sizes are given in
            // traversed elements (16 px -> 32, 10 rows -> 20)
            cuuint64_t dims[4] = {64, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
            cuuint64_t strides[3] = {128, (cuuint64_t)p.W * 128, (cuuint64_t)p.H * p.W * 128};
            cuuint32_t box[4] = {64, 2 * TW, 20, 1};
            cuuint32_t estr[4] = {1, 2, 2, 1};
            CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(p.x_h), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16]: cuTensorMapEncodeTiled(x, 5x5/s2) failed with %d", (int)r);
        }
        {
            cuuint64_t dims[2] = {64, (cuuint64_t)25 * p.cout_pad};
            cuuint64_t strides[1] = {128};
            cuuint32_t box[2] = {64, 64};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.w_h), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16]: cuTensorMapEncodeTiled(w, 5x5/s2) failed with %d", (int)r);
        }
        Params q;
        q.bias = p.bias; q.residual = p.residual; q.y = p.y; q.yh = reinterpret_cast<__half *>(p.y_h);
        q.N = p.N; q.H = Ho; q.W = Wo;                      // the epilogue works on OUTPUT pixels
        q.Cout = p.Cout; q.y_pitch = p.y_pitch; q.y_coff = p.y_coff;
        q.dilation = 1; q.flags = p.flags;
        q.tiles_x = ceil_div(Wo, TW);
        q.tiles_y = ceil_div(Ho, TH);
        const int cout_tiles = p.cout_pad / 64;
        const int ptiles = p.N * q.tiles_x * q.tiles_y;
        int per_ct = n_sm / cout_tiles;
        if (per_ct < 1) per_ct = 1;
        if (per_ct > ptiles) per_ct = ptiles;
        const int n_stages = (conv_smem_cap() - 1024 - 512) / (PIPES * K5_STAGE_BYTES) >= 2 ? 2 : 1;
        const int smem_bytes = PIPES * n_stages * K5_STAGE_BYTES + 1024 + 512;
        conv_f16_kernel<1><<<dim3(per_ct, cout_tiles), THREADS, smem_bytes, st>>>(map_x, map_w, q, n_stages, K5_STAGE_BYTES, ptiles);
        L3C_LAUNCH_CHECK("conv5x5s2_f16_kernel");
        return L3C_OK;
    }
    // ---- 1x1
    L3C_REQUIRE(p.cout_pad <= 256, "l3c_conv2d[f16]: 1x1 layers are built for Cout <= 256 (got %d)", p.Cout);
    L3C_REQUIRE(p.Cout % 2 == 0 && p.y_pitch % 2 == 0 && p.y_coff % 2 == 0, "l3c_conv2d[f16]: 1x1 layers need even Cout/pitch/offset");
    L3C_REQUIRE(!p.residual, "l3c_conv2d[f16]: residual on a 1x1 layer is not built");
    const int kchunks = p.Cin / 64;
    const long long M = (long long)p.N * p.H * p.W;
    const int w_bytes = kchunks * p.cout_pad * 128;
    int n_stages = (conv_smem_cap() - 1024 - 512 - w_bytes) / K1_A_BYTES;
    if (n_stages < 1 && w_bytes + K1_A_BYTES <= 227 * 1024 - 1024 - 512) n_stages = 1;
    if (n_stages > K1_MAX_STAGES) n_stages = K1_MAX_STAGES;
    L3C_REQUIRE(n_stages >= 2, "l3c_conv2d[f16]: weights of a %d -> %d 1x1 layer do not fit in shared memory", p.Cin, p.Cout);
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)p.x_pitch * 2};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.x_h), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16]: cuTensorMapEncodeTiled(x, 1x1) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)kchunks * p.cout_pad};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)p.cout_pad};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(p.w_h), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d[f16]: cuTensorMapEncodeTiled(w, 1x1) failed with %d", (int)r);
    }
    Params1 q;
    q.bias = p.bias; q.y = p.y; q.yh = reinterpret_cast<__half *>(p.y_h);
    q.M = M; q.Cout = p.Cout; q.y_pitch = p.y_pitch; q.y_coff = p.y_coff; q.flags = p.flags;
    q.kchunks = kchunks; q.npad = p.cout_pad;
    q.tmem_cols = 2 * p.cout_pad <= 128 ? 128 : (2 * p.cout_pad <= 256 ? 256 : 512);
    const int n_tiles = (int)((M + 127) / 128);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    const int smem_bytes = w_bytes + n_stages * K1_A_BYTES + 1024 + 512;
    conv1x1_f16_kernel<<<grid, K1_THREADS, smem_bytes, st>>>(map_x, map_w, q, n_stages, n_tiles);
    L3C_LAUNCH_CHECK("conv1x1_f16_kernel");
    return L3C_OK;
}

}  // namespace l3c

// 1x1 conv (Cin % 64 == 0 -> Kp = (rgb ? 4 : 3) * C * 10 parameters) fused with the per-symbol coding intervals
extern "C" int l3c_lin_dmll_intervals(const void *x_h, const void *w_h, const float *bias, const uint8_t *sym_dev,
                                      const float *targets_dev, int N, int HW, int Cin, int C, int K, int L, int rgb,
                                      uint32_t *intervals_dev, void *stream) {
    using namespace l3c;
    using namespace l3c::f16;
    L3C_REQUIRE(x_h && w_h && bias && sym_dev && targets_dev && intervals_dev, "l3c_lin_dmll_intervals: null pointer");
    L3C_REQUIRE(N >= 1 && HW >= 1 && Cin % 64 == 0 && Cin >= 64, "l3c_lin_dmll_intervals: N=%d HW=%d Cin=%d", N, HW, Cin);
    L3C_REQUIRE(K == 10 && L >= 2 && L <= 256, "l3c_lin_dmll_intervals: K=%d L=%d (only K=10 is built)", K, L);
    L3C_REQUIRE((rgb && C == 3) || (!rgb && C == 5), "l3c_lin_dmll_intervals: built for RGB (C=3) and C=5 bottlenecks, "
                                                     "got C=%d rgb=%d", C, rgb);
    const int Kp = (rgb ? 4 : 3) * C * K;
    const int npad = (Kp + 63) / 64 * 64;
    cudaStream_t st = (cudaStream_t)stream;
    EncodeTiledFn encode = get_encode_fn();
    L3C_REQUIRE(encode != nullptr, "l3c_lin_dmll_intervals: cuTensorMapEncodeTiled is not available from the driver");
    static bool configured_dev[64] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        L3C_CUDA(cudaFuncSetAttribute(lin_dmll_f16_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        L3C_CUDA(cudaFuncSetAttribute(lin_dmll_f16_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    const int kchunks = Cin / 64;
    const long long M = (long long)N * HW;
    const int w_bytes = kchunks * npad * 128;
    int n_stages = (conv_smem_cap() - 1024 - 512 - w_bytes) / K1_A_BYTES;
    if (n_stages < 1 && w_bytes + K1_A_BYTES <= 227 * 1024 - 1024 - 512) n_stages = 1;
    if (n_stages > K1_MAX_STAGES) n_stages = K1_MAX_STAGES;
    L3C_REQUIRE(n_stages >= 2, "l3c_lin_dmll_intervals: weights do not fit in shared memory (Cin=%d)", Cin);
    alignas(64) CUtensorMap map_x, map_w;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(x_h), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_lin_dmll_intervals: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)kchunks * npad};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)npad};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(w_h), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_lin_dmll_intervals: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
    }
    ParamsLD q;
    q.bias = bias; q.sym = sym_dev; q.targets = targets_dev; q.intervals = intervals_dev;
    q.M = M; q.HW = HW; q.L = L; q.kchunks = kchunks; q.npad = npad;
    q.tmem_cols = 2 * npad <= 128 ? 128 : (2 * npad <= 256 ? 256 : 512);
    const int n_tiles = (int)((M + 127) / 128);
    const int n_sm = stream_sm_count(st);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    const int smem_bytes = w_bytes + n_stages * K1_A_BYTES + 1024 + 512;
    if (rgb)
        lin_dmll_f16_kernel<3, true><<<grid, LD_THREADS, smem_bytes, st>>>(map_x, map_w, q, n_stages, n_tiles);
    else
        lin_dmll_f16_kernel<5, false><<<grid, LD_THREADS, smem_bytes, st>>>(map_x, map_w, q, n_stages, n_tiles);
    L3C_LAUNCH_CHECK("lin_dmll_f16_kernel");
    return L3C_OK;
}
