// conv_tcgen05.cu -- 3x3 (optionally dilated) 64-input-channel convolution as an implicit GEMM on the
// Blackwell 5th-generation tensor cores: TMA (cp.async.bulk.tensor) stages shifted NHWC windows and
// weight slabs into 128B-swizzled shared memory, ONE elected thread issues tcgen05.mma (kind::tf32,
// M=128 pixels x N=64 output channels x K=8), accumulators live in TMEM (two 64-column buffers so
// the epilogue of tile i overlaps the MMAs of tile i+1), a 4-warp epilogue drains TMEM with
// tcgen05.ld and fuses bias / ReLU / residual / channel-slice / PixelShuffle(2).
//
// Reference layers this replaces: every `default_conv(64, Cout, 3, rate=r)` of the L3C stack
// (/root/reference/src/pytorch_ext.py:57-61; ResBlock edsr.py:63-89, body-final conv net.py:110,170,
// Head head.py:49-56, Upsampler edsr.py:92-101, atrous convs prob_clf.py:54-55) -- cuDNN fp32 in the
// reference.
//
// Implicit GEMM mapping (per 8x16-pixel tile, per filter tap t=(ky,kx)):
//   A[m][k] = x[n, ty0 + m/16 + (ky-1)*d, tx0 + m%16 + (kx-1)*d, k]   (zero outside the image: TMA
//             out-of-bounds fill implements the conv padding), m = 0..127, k = 0..63
//   B[n][k] = w[n, k, ky, kx]                                          n = 0..63 (one Cout tile)
//   D[m][n] += sum_k A[m][k] * B[n][k]        -> 9 taps x 2 K-halves x 4 tcgen05.mma (K=8 tf32 each)
// Accumulation order per output element is fixed (tap, K-half, k-step), independent of batch,
// image size and tile position: encoder-side and decoder-side evaluations are bit-identical.
//
// Precision: L3C_PREC_TF32.  The tensor core reads only the upper 19 bits of each fp32 operand
// (truncation), which measurably biases the bit cost (+7e-4 bpsp at 512^2).  Operands are therefore
// pre-rounded to TF32 with round-to-nearest: weights once on the host side, activations by the
// producing layer's epilogue, which writes a second, rounded copy (`y_tf32`) next to the fp32 tensor
// that residual adds and the fp32 layers keep using.  Accumulation is fp32 in TMEM.
// (The faster FP16-operand variant of this path is conv_f16.cu, precision L3C_PREC_F16.)
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace l3c {

namespace tc {

constexpr int TH = 8, TW = 16;                 // 128 output pixels = UMMA M
constexpr int STAGES = 4;
constexpr int A_HALF = 128 * 128;              // 128 pixel rows x 128 B (32 fp32 channels)
constexpr int B_HALF = 64 * 128;               // 64 cout rows x 128 B
constexpr int STAGE_BYTES = 2 * A_HALF + 2 * B_HALF;   // 48 KB per filter tap
constexpr int THREADS = 192;                   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int ACC_COLS = 64;
constexpr int TMEM_COLS = 128;                 // two accumulators
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

// ---- PTX wrappers shared with conv_f16.cu: tc_ptx.cuh
using namespace tcx;

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


// instruction descriptor: D=f32, A=B=tf32, both K-major, N=64, M=128
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ float round_tf32(float x) {      // round-to-nearest TF32 image of x
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct Params {
    const float *bias;      // padded to cout_pad
    const float *residual;
    float *y;
    float *y_tf32;
    int N, H, W;
    int Cout, y_pitch, y_coff;
    int dilation;
    unsigned flags;
    int tiles_x, tiles_y, cout_tiles, total_tiles;
    int taps;        // 9 (3x3) or 1 (1x1)
    int kpairs;      // Cin / 64: stage iterations per tap, each stage = 2 K-halves of 32 channels
    int cout_rows;   // rows per (tap, K-half) slab of the weight image = cout_tiles * 64
};

// Epilogue of one 128-pixel x 64-channel accumulator (this thread = one pixel row of the tile):
// TMEM -> registers -> bias / ReLU / residual / TF32 twin -> global (NHWC, channel slice, or
// PixelShuffle(2) addressing).
template <int NCHUNK>
__device__ __forceinline__ void epilogue_tile(const Params &p, uint32_t taddr, int ct, int n, int oy, int ox,
                                              bool inside, bool relu, bool shuffle, bool round_y, int cbeg,
                                              const float *bias_regs /* NCHUNK*16 preloaded values or null */) {
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c0 = cbeg + ch * 16;
        float v[16];
        tmem_ld16(taddr + c0, v);
        const int co0 = ct * 64 + c0;
        if (inside && co0 < p.Cout) {
            if (bias_regs) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += bias_regs[ch * 16 + i];
            } else {
                const float4 *bp = reinterpret_cast<const float4 *>(p.bias + co0);   // padded, 64 B aligned
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b = __ldg(bp + (i >> 2));
                    v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                }
            }
            if (relu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (!shuffle) {
                const size_t off = (((size_t)n * p.H + oy) * p.W + ox) * p.y_pitch + p.y_coff + co0;
                const bool vec = (((p.y_pitch | p.y_coff) & 3) == 0) && (co0 + 16 <= p.Cout);
                if (vec) {
                    if (p.residual) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 r = __ldg(reinterpret_cast<const float4 *>(p.residual + off + i));
                            v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
                        }
                    }
                    if (p.y_tf32) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4 *>(p.y_tf32 + off + i) =
                                make_float4(round_tf32(v[i]), round_tf32(v[i + 1]), round_tf32(v[i + 2]),
                                            round_tf32(v[i + 3]));
                    }
                    if (round_y) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = round_tf32(v[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4 *>(p.y + off + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (co0 + i < p.Cout) {
                            float o = v[i];
                            if (p.residual) o += __ldg(p.residual + off + i);
                            if (p.y_tf32) p.y_tf32[off + i] = round_tf32(o);
                            p.y[off + i] = round_y ? round_tf32(o) : o;
                        }
                    }
                }
            } else {
                // out[n, 2*oy+i, 2*ox+j, cq] = conv[n, oy, ox, 4*cq + 2*i + j]
                const int H2 = p.H * 2, W2 = p.W * 2;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int co = co0 + i;
                    const int cq = co >> 2, si = (co >> 1) & 1, sj = co & 1;
                    const size_t o2 = (((size_t)n * H2 + 2 * oy + si) * W2 + 2 * ox + sj) * p.y_pitch + p.y_coff + cq;
                    if (p.y_tf32) p.y_tf32[o2] = round_tf32(v[i]);
                    p.y[o2] = round_y ? round_tf32(v[i]) : v[i];
                }
            }
        }
    }
}

// 4-warp epilogue loop of kernel v1: tile t of this CTA's sequence t0, t0+tstride, ...
__device__ __forceinline__ void epilogue_loop(const Params &p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                              int warp, int lane, int t0, int tstride, int t_end, int ct_fixed) {
    const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31 are this warp's
    const int m = quarter * 32 + lane;            // pixel row of the tile
    uint32_t acc = 0, acc_phase = 0;
    const bool relu = (p.flags & L3C_CONV_RELU) != 0;
    const bool shuffle = (p.flags & L3C_CONV_PIXEL_SHUFFLE2) != 0;
    const bool round_y = (p.flags & L3C_CONV_ROUND_TF32) != 0;
    for (int t = t0; t < t_end; t += tstride) {
        int q = t;
        int ct = ct_fixed;
        if (ct_fixed < 0) { ct = q % p.cout_tiles; q /= p.cout_tiles; }
        const int tx = q % p.tiles_x; q /= p.tiles_x;
        const int ty = q % p.tiles_y; q /= p.tiles_y;
        const int n = q;
        const int oy = ty * TH + (m >> 4);
        const int ox = tx * TW + (m & 15);
        const bool inside = (oy < p.H) && (ox < p.W);
        mbar_wait(tfull0 + 8u * acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * ACC_COLS + ((uint32_t)(quarter * 32) << 16);
        epilogue_tile<4>(p, taddr, ct, n, oy, ox, inside, relu, shuffle, round_y, 0, nullptr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
    }
}

__global__ void __launch_bounds__(THREADS, 1)
conv3x3_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                       const Params p) {
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle atoms need 1024 B alignment
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    // bars[0..3] full, [4..7] empty, [8..9] tmem_full, [10..11] tmem_empty
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (8 + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (10 + a); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);          // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int d = p.dilation;

    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t stage = 0, phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            int q = t;
            const int ct = q % p.cout_tiles; q /= p.cout_tiles;
            const int tx = q % p.tiles_x; q /= p.tiles_x;
            const int ty = q % p.tiles_y; q /= p.tiles_y;
            const int n = q;
            for (int it = 0; it < p.taps * p.kpairs; ++it) {
                const int tap = it / p.kpairs, kp = it % p.kpairs;
                if (lane == 0) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    const uint32_t a0 = smem_base + stage * STAGE_BYTES;
                    const int x0 = tx * TW + (p.taps == 9 ? (tap % 3 - 1) * d : 0);
                    const int y0 = ty * TH + (p.taps == 9 ? (tap / 3 - 1) * d : 0);
                    tma_load_4d(a0, &map_x, full_bar(stage), kp * 64, x0, y0, n);
                    tma_load_4d(a0 + A_HALF, &map_x, full_bar(stage), kp * 64 + 32, x0, y0, n);
                    // weight image rows: [tap][K-half][cout_rows]
                    const int wrow = (tap * 2 * p.kpairs + 2 * kp) * p.cout_rows + ct * 64;
                    tma_load_2d(a0 + 2 * A_HALF, &map_w, full_bar(stage), 0, wrow);
                    tma_load_2d(a0 + 2 * A_HALF + B_HALF, &map_w, full_bar(stage), 0, wrow + p.cout_rows);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0, acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            if (lane == 0) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);       // epilogue has drained this accumulator
                tc_fence_after();
            }
            __syncwarp();
            const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
            const int n_it = p.taps * p.kpairs;
            for (int it = 0; it < n_it; ++it) {
                if (lane == 0) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_base + stage * STAGE_BYTES;
#pragma unroll
                    for (int kh = 0; kh < 2; ++kh) {
                        const uint64_t da = make_desc(a0 + kh * A_HALF);
                        const uint64_t db = make_desc(a0 + 2 * A_HALF + kh * B_HALF);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // +32 B (8 tf32) along K inside the 128 B swizzle row: +2 in 16-byte units
                            mma_tf32(d_tmem, da + 2u * k, db + 2u * k, IDESC_TF32,
                                     (it | kh | k) != 0 ? 1u : 0u);
                        }
                    }
                    mma_commit(empty_bar(stage));                 // frees the smem slot when the MMAs retire
                    if (it == n_it - 1) mma_commit(tfull_bar(acc));   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        epilogue_loop(p, tmem_base, tfull_bar(0), tempty_bar(0), warp, lane, blockIdx.x, gridDim.x, p.total_tiles, -1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel v2 for the 3x3 layers: weights resident in shared memory, one halo load per filter COLUMN.
//
// v1 re-stages a shifted 128-pixel window (32 KB) plus a 16 KB weight slab for each of the 9 taps:
// 432 KB of L2 -> smem traffic per 128-pixel tile, which (at the ~51 B/clk/SM the TMA path sustains)
// caps it near 27 % of the TF32 pipe.  Here
//   * the 9 x 2 weight slabs of this CTA's Cout tile (144 KB) are loaded ONCE per CTA;
//   * for each dx in {0,1,2} and K-half, ONE box of (8 + 2d) rows x 16 pixels x 32 channels is staged
//     (x origin shifted by (dx-1)*d): the three dy taps of that column are the same buffer read at row
//     offsets dy*d*16 pixels = dy*d*2048 B -- multiples of the 1024 B swizzle atom, so plain UMMA
//     descriptors work without base-offset tricks;
//   -> 6 loads of 20 KB (d=1) per tile = 120 KB: 3.6x less staging traffic, 12 MMAs per load.
// Accumulation order per output element: (dx, K-half, dy, k-step) -- fixed, position independent.
// ---------------------------------------------------------------------------------------------
constexpr int W_RES_BYTES = 9 * 2 * B_HALF;       // 147456
constexpr int V2_PIPES = 2;                       // independent (TMA producer, MMA issuer) pairs
constexpr int V2_EPI_WARPS = 8;                   // two warps per TMEM lane quarter, 32 columns each
constexpr int V2_THREADS = 32 * (2 * V2_PIPES + V2_EPI_WARPS);
constexpr int V2_TMEM_COLS = 256;                 // 2 pipes x 2 accumulators x 64 columns

// ncu on kernel v1 (profiles/r01): the tensor pipe is busy 30 % of the time although L2 staging was
// cut 3.6x -- the limiter is the ISSUE of the MMAs: with N = 64 one tcgen05.mma is only 32 cycles of
// tensor work, while the single issuing thread needs ~25 SASS instructions (ELECT, five
// R2UR.BROADCAST, descriptor adds ...) ~ 100 cycles per MMA.  v2 therefore runs TWO issuer threads
// (on different warps / schedulers), each with its own producer warp, smem stages and pair of TMEM
// accumulators, working on alternate tiles; every output element is still accumulated by exactly one
// issuer in the fixed order (dx, K-half, dy, k-step), so results stay bit-reproducible.
__global__ void __launch_bounds__(V2_THREADS, 1)
conv3x3_tcgen05_v2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                          const Params p, const int n_stages, const int a_bytes, const int ptiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // [weights 144 KB][pipe 0 stages][pipe 1 stages][barriers]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + W_RES_BYTES + V2_PIPES * n_stages * a_bytes);
    // per pipe k (base 8k): full[0..1], empty[2..3], tfull[4..5], tempty[6..7]; bars[16] = weights
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 17);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t w_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    const uint32_t wbar = bar_base + 8u * 16;

    if (threadIdx.x == 0) {
        for (int k = 0; k < V2_PIPES; ++k) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(bar_base + 8u * (8 * k + s), 1);          // full
                mbar_init(bar_base + 8u * (8 * k + 2 + s), 1);      // empty
                mbar_init(bar_base + 8u * (8 * k + 4 + s), 1);      // tmem full
                mbar_init(bar_base + 8u * (8 * k + 6 + s), V2_EPI_WARPS);   // tmem empty: one arrival per epilogue warp
            }
        }
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == V2_PIPES) tmem_alloc(smem_u32(tmem_slot), V2_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int d = p.dilation;
    const int ct = blockIdx.y;                        // this CTA's Cout tile (weights stay resident)

    if (warp < V2_PIPES) {
        // ===================== TMA producer of pipe `warp` =====================
        const int k = warp;
        const uint32_t a_base = w_base + W_RES_BYTES + k * n_stages * a_bytes;
        const uint32_t full0 = bar_base + 8u * (8 * k), empty0 = bar_base + 8u * (8 * k + 2);
        if (k == 0 && lane == 0) {
            mbar_expect_tx(wbar, W_RES_BYTES);
            for (int slab = 0; slab < 18; ++slab)     // slab = tap*2 + K-half; rows [slab][cout_rows]
                tma_load_2d(w_base + slab * B_HALF, &map_w, wbar, 0, slab * p.cout_rows + ct * 64);
        }
        uint32_t stage = 0, phase = 0;
        for (int t = blockIdx.x + k * gridDim.x; t < ptiles; t += V2_PIPES * gridDim.x) {
            int q = t;
            const int tx = q % p.tiles_x; q /= p.tiles_x;
            const int ty = q % p.tiles_y; q /= p.tiles_y;
            const int n = q;
            for (int unit = 0; unit < 6; ++unit) {    // unit = dx*2 + K-half
                if (lane == 0) {
                    mbar_wait(empty0 + 8u * stage, phase ^ 1u);
                    mbar_expect_tx(full0 + 8u * stage, a_bytes);
                    tma_load_4d(a_base + stage * a_bytes, &map_x, full0 + 8u * stage, (unit & 1) * 32,
                                tx * TW + ((unit >> 1) - 1) * d, ty * TH - d, n);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp < 2 * V2_PIPES) {
        // ===================== MMA issuer of pipe `warp - V2_PIPES` =====================
        const int k = warp - V2_PIPES;
        const uint32_t a_base = w_base + W_RES_BYTES + k * n_stages * a_bytes;
        const uint32_t full0 = bar_base + 8u * (8 * k), empty0 = bar_base + 8u * (8 * k + 2);
        const uint32_t tfull0 = bar_base + 8u * (8 * k + 4), tempty0 = bar_base + 8u * (8 * k + 6);
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0, acc_phase = 0;
        mbar_wait(wbar, 0);
        for (int t = blockIdx.x + k * gridDim.x; t < ptiles; t += V2_PIPES * gridDim.x) {
            // warp-uniform code, MMAs issued through elect.sync: descriptors stay in uniform registers
            mbar_wait(tempty0 + 8u * acc, acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (k * 2 + acc) * ACC_COLS;
            for (int unit = 0; unit < 6; ++unit) {
                mbar_wait(full0 + 8u * stage, phase);
                tc_fence_after();
                const int dx = unit >> 1, kh = unit & 1;
                const uint32_t a0 = a_base + stage * a_bytes;
                if (elect_one()) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint64_t da = make_desc(a0 + dy * d * (TW * 128));           // dy*d rows of 16 px
                        const uint64_t db = make_desc(w_base + ((dy * 3 + dx) * 2 + kh) * B_HALF);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            mma_tf32(d_tmem, da + 2u * kk, db + 2u * kk, IDESC_TF32, (unit | dy | kk) != 0 ? 1u : 0u);
                    }
                    mma_commit(empty0 + 8u * stage);
                    if (unit == 5) mma_commit(tfull0 + 8u * acc);
                }
                __syncwarp();
                if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ===================== epilogue: tiles in sequence order, alternating pipes =====================
        // warp e of the 8 epilogue warps: TMEM lane quarter = warp id % 4 (hardware rule), column half
        // h = e / 4 (32 columns).  Accumulators are read with the 16x256b shape so that a quad of threads
        // owns 32 contiguous bytes of a pixel row: every global access below moves whole 32 B sectors
        // (the 32x32b shape made each STG.128 touch 32 half-used sectors).
        const int e = warp - 2 * V2_PIPES;
        const int quarter = warp & 3;
        const int cbeg = (e >> 2) * 32;
        const int tq = lane & 3, tr = lane >> 2;              // column pair / row inside the fragment
        const bool relu = (p.flags & L3C_CONV_RELU) != 0;
        const bool shuffle = (p.flags & L3C_CONV_PIXEL_SHUFFLE2) != 0;
        const bool round_y = (p.flags & L3C_CONV_ROUND_TF32) != 0;
        float2 bias2[4];                                      // bias of columns cbeg + 8j + 2*tq + {0,1}
#pragma unroll
        for (int jb = 0; jb < 4; ++jb)
            bias2[jb] = __ldg(reinterpret_cast<const float2 *>(p.bias + ct * 64 + cbeg + 8 * jb + 2 * tq));
        int j = 0;
        for (int t = blockIdx.x; t < ptiles; t += gridDim.x, ++j) {
            const int k = j % V2_PIPES;
            const uint32_t acc = (uint32_t)(j / V2_PIPES) & 1u;
            const uint32_t acc_phase = (uint32_t)(j / (2 * V2_PIPES)) & 1u;
            const uint32_t tfull = bar_base + 8u * (8 * k + 4 + acc), tempty = bar_base + 8u * (8 * k + 6 + acc);
            int q = t;
            const int tx = q % p.tiles_x; q /= p.tiles_x;
            const int ty = q % p.tiles_y; q /= p.tiles_y;
            const int n = q;
            mbar_wait(tfull, acc_phase);
            tc_fence_after();
#pragma unroll
            for (int lh = 0; lh < 2; ++lh) {
                // TMEM lanes 32*quarter + 16*lh + {tr, tr+8}  <->  tile rows (pixels) of the same index
                float v[16];
                tmem_ld_16x256b_x4(tmem_base + (k * 2 + acc) * ACC_COLS + cbeg +
                                       ((uint32_t)(quarter * 32 + lh * 16) << 16), v);
                const int oy = ty * TH + 2 * quarter + lh;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int ox = tx * TW + tr + 8 * rh;
                    if (oy >= p.H || ox >= p.W) continue;
                    const size_t pix = ((size_t)n * p.H + oy) * p.W + ox;
#pragma unroll
                    for (int jb = 0; jb < 4; ++jb) {
                        float a = v[4 * jb + 2 * rh] + bias2[jb].x;
                        float b = v[4 * jb + 2 * rh + 1] + bias2[jb].y;
                        if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                        const int co = ct * 64 + cbeg + 8 * jb + 2 * tq;
                        if (!shuffle) {
                            const size_t off = pix * p.y_pitch + p.y_coff + co;
                            if (p.residual) {
                                const float2 r = __ldg(reinterpret_cast<const float2 *>(p.residual + off));
                                a += r.x; b += r.y;
                            }
                            if (p.y_tf32) *reinterpret_cast<float2 *>(p.y_tf32 + off) = make_float2(round_tf32(a), round_tf32(b));
                            if (round_y) { a = round_tf32(a); b = round_tf32(b); }
                            *reinterpret_cast<float2 *>(p.y + off) = make_float2(a, b);
                        } else {
                            // out[n, 2*oy+i, 2*ox+jj, cq] = conv[n, oy, ox, 4*cq + 2*i + jj]; co is even:
                            // (co, co+1) -> same cq and i, jj = 0 / 1 -> two horizontally adjacent pixels
                            const int cq = co >> 2, si = (co >> 1) & 1;
                            const size_t o2 = (((size_t)n * (2 * p.H) + 2 * oy + si) * (2 * p.W) + 2 * ox) * p.y_pitch +
                                              p.y_coff + cq;
                            if (p.y_tf32) { p.y_tf32[o2] = round_tf32(a); p.y_tf32[o2 + p.y_pitch] = round_tf32(b); }
                            if (round_y) { a = round_tf32(a); b = round_tf32(b); }
                            p.y[o2] = a;
                            p.y[o2 + p.y_pitch] = b;
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
    if (warp == V2_PIPES) {
        tc_fence_after();
        tmem_dealloc(tmem_base, V2_TMEM_COLS);
    }
}

}  // namespace tc

namespace tcx {
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
}  // namespace tcx

int conv2d_tcgen05(const l3c_conv_t &p, cudaStream_t st) {
    using namespace tc;
    if (p.precision != L3C_PREC_TF32) {
        set_error("l3c_conv2d: precision mode %d is not built (available: fp32, tf32, f16)", p.precision);
        return L3C_EINVAL;
    }
    // The tensor-core kernel covers the 3x3 / 64-input-channel layers (> 92 % of the FLOPs of a round
    // trip).  The host mirror routes the other layers (5x5/s2 down-sampling, 1x1 convs, Cin != 64) to
    // L3C_PREC_FP32 explicitly (engine.conv2d); nothing falls back silently here.
    const bool k3 = p.ksize == 3 && p.stride == 1 && p.Cin == 64 && p.x_pitch == 64;
    const bool k1 = p.ksize == 1 && p.stride == 1 && p.Cin % 64 == 0 && p.x_pitch == p.Cin &&
                    !(p.flags & L3C_CONV_PIXEL_SHUFFLE2);
    const bool eligible = (k3 || k1) && p.cout_pad % 64 == 0 && p.cout_pad >= p.Cout &&
                          (k1 || ((p.Cout % 64 == 0) && (p.y_pitch % 4 == 0) && (p.y_coff % 4 == 0)));
    L3C_REQUIRE(eligible, "l3c_conv2d: tcgen05 path needs 3x3/Cin=64 or 1x1/Cin%%64==0, stride 1 "
                          "(got k=%d s=%d Cin=%d pitch=%d Cout=%d)", p.ksize, p.stride, p.Cin, p.x_pitch, p.Cout);
    const int taps = k3 ? 9 : 1;
    const int kpairs = p.Cin / 64;
    EncodeTiledFn encode = get_encode_fn();
    L3C_REQUIRE(encode != nullptr, "l3c_conv2d: cuTensorMapEncodeTiled is not available from the driver");

    // kernel v2 (weights resident, one halo load per filter column) for the 3x3 layers
    const int d = p.dilation;
    const int a_rows = TH + 2 * d;
    const int a_bytes = a_rows * TW * 128;
    int n_stages = (227 * 1024 - 1024 - 512 - W_RES_BYTES) / (V2_PIPES * a_bytes);      // per pipe
    if (n_stages > 2) n_stages = 2;
    const bool use_v2 = k3 && a_rows <= 256 && n_stages >= 1;

    alignas(64) CUtensorMap map_x, map_w;
    {
        cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)p.x_pitch * 4, (cuuint64_t)p.W * p.x_pitch * 4,
                                 (cuuint64_t)p.H * p.W * p.x_pitch * 4};
        cuuint32_t box[4] = {32, TW, (cuuint32_t)(use_v2 ? a_rows : TH), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(p.x), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    const int cout_tiles = p.cout_pad / 64;
    {
        // tensor-core weight image: [tap][K-half = Cin/32][cout_pad][32] fp32 (engine.PackedConv.get_tc)
        cuuint64_t dims[2] = {32, (cuuint64_t)taps * 2 * kpairs * p.cout_pad};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {32, 64};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p.w), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        L3C_REQUIRE(r == CUDA_SUCCESS, "l3c_conv2d: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
    }
    Params q;
    q.bias = p.bias;
    q.residual = p.residual;
    q.y = p.y;
    q.y_tf32 = p.y_tf32;
    q.N = p.N; q.H = p.H; q.W = p.W;
    q.Cout = p.Cout; q.y_pitch = p.y_pitch; q.y_coff = p.y_coff;
    q.dilation = p.dilation;
    q.flags = p.flags;
    q.tiles_x = ceil_div(p.W, TW);
    q.tiles_y = ceil_div(p.H, TH);
    q.cout_tiles = cout_tiles;
    q.taps = taps;
    q.kpairs = kpairs;
    q.cout_rows = p.cout_pad;
    q.total_tiles = p.N * q.tiles_x * q.tiles_y * cout_tiles;

    static bool configured_dev[64] = {};
    bool &configured = configured_dev[current_device_slot()];
    const int n_sm = stream_sm_count(st);          // the stream may be confined to a group of SMs
    if (!configured) {
        L3C_CUDA(cudaFuncSetAttribute(conv3x3_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        L3C_CUDA(cudaFuncSetAttribute(conv3x3_tcgen05_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
        configured = true;
    }
    if (use_v2) {
        const int ptiles = p.N * q.tiles_x * q.tiles_y;
        int per_ct = n_sm / cout_tiles;
        if (per_ct < 1) per_ct = 1;
        if (per_ct > ptiles) per_ct = ptiles;
        const int smem_bytes = W_RES_BYTES + V2_PIPES * n_stages * a_bytes + 1024 + 512;
        conv3x3_tcgen05_v2_kernel<<<dim3(per_ct, cout_tiles), V2_THREADS, smem_bytes, st>>>(map_x, map_w, q, n_stages,
                                                                                          a_bytes, ptiles);
        L3C_LAUNCH_CHECK("conv3x3_tcgen05_v2_kernel");
        return L3C_OK;
    }
    const int grid = q.total_tiles < n_sm ? q.total_tiles : n_sm;      // persistent: one CTA per SM
    conv3x3_tcgen05_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map_x, map_w, q);
    L3C_LAUNCH_CHECK("conv3x3_tcgen05_kernel");
    return L3C_OK;
}

}  // namespace l3c
