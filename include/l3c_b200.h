/*
 * l3c_b200.h -- C ABI of libl3c_b200.so: the sm_100a implementation of L3C's encode/decode hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every entry point returns 0 on success or a
 * negative L3C_E* code; l3c_last_error() gives the message of the last failure on the calling
 * thread.  "dev" pointers are CUDA device pointers on the current device, "host" pointers are
 * ordinary (ideally pinned) host memory.  `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream).  Kernels are asynchronous on `stream` unless a function says it
 * synchronises.
 *
 * Section A are the drop-in replacements for the FIVE exports of the reference's native module
 * (pybind module torchac_backend_{cpu,gpu}: /root/reference/src/torchac/torchac_backend/
 * torchac.cpp:433-443).  Sections B..E are the batched, device-resident building blocks the host
 * mirror (l3c_pytorch_b200.*) composes into the reference's Bitcoding / MultiscaleBlueprint API.
 *
 * Layout conventions
 *   activations ........ NHWC fp32, channel pitch given explicitly where slices are written
 *   conv weights ....... [KH][KW][Cin][CoutPad] fp32 (host mirror repacks the reference's OIHW)
 *   symbol planes ...... uint8  [N][C][H*W]    (one plane = one coded stream, row-major H,W;
 *                                               flatten order of coders.py:49-50)
 *   intervals .......... uint32 [N][C][H*W]    c_low | (c_high-1) << 16
 *   CDF tables ......... uint16 rows of `pitch` entries (entries 0..L-1 used; the reference's
 *                        (L+1)-th entry is never read, torchac.cpp:181,280)
 */
#ifndef L3C_B200_H_
#define L3C_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L3C_OK 0
#define L3C_EINVAL (-1)   /* bad argument (shape/alignment/range)      -> reference raises RuntimeError */
#define L3C_ECUDA (-2)    /* CUDA runtime error                                                          */
#define L3C_EOVERFLOW (-3)/* output buffer too small                                                     */
#define L3C_ENODEV (-4)   /* no sm_100 device                                                            */
#define L3C_EUNSUPPORTED (-5) /* optional facility not offered by this driver (caller may do without)      */

const char *l3c_last_error(void);
int l3c_version(void);

/* Launch log (no reference counterpart; evidence plumbing): every kernel this library launches is
 * counted under its kernel name.  Writes "kernel_name count\n" lines into buf (truncated to cap, may be
 * NULL) and returns the TOTAL number of launches since the last reset; reset != 0 clears the counters. */
long long l3c_launch_log(char *buf, size_t cap, int reset);

/* ------------------------------------------------------------------------------------------
 * A. Drop-in native exports (reference: torchac.cpp:433-443; python shim torchac.py:87-166)
 * ---------------------------------------------------------------------------------------- */

/* replaces `cuda_supported()` (torchac.cpp:439,441): 1 iff an sm_100 device is usable. */
int l3c_cuda_supported(void);

/* replaces `encode_cdf(cdf, sym) -> bytes` (torchac.cpp:262-269 -> encode() :152-227).
 * cdf_host: int16/uint16 [n_sym][Lp] (the 1HWLp tensor, flattened); sym_host: int16 [n_sym].
 * Synchronous.  *out_len receives the byte count; L3C_EOVERFLOW if it exceeds out_cap. */
int l3c_encode_cdf(const uint16_t *cdf_host, int64_t n_sym, int Lp, const int16_t *sym_host,
                   uint8_t *out_host, size_t out_cap, size_t *out_len);

/* replaces `decode_cdf(cdf, in) -> int16 tensor` (torchac.cpp:424-430 -> decode() :299-381).
 * n_sym comes from the cdf shape, exactly as in the reference; short input is zero-filled. */
int l3c_decode_cdf(const uint16_t *cdf_host, int64_t n_sym, int Lp, const uint8_t *in_host,
                   size_t in_len, int16_t *sym_out_host);

/* replaces `encode_logistic_mixture(targets, means, log_scales, logit_probs_softmax, sym)`
 * (torchac.cpp:231-258; CDF formula torchac_kernel.cu:20-76).  targets_dev f32[Lp];
 * means/log_scales/probs_dev f32 [K][n_sym] (1KHW contiguous, CUDA); sym_host int16 (CPU, as in
 * the reference).  No CDF table is materialised: only the two bounds of each coded symbol are
 * evaluated.  Synchronous. */
int l3c_encode_logistic_mixture(const float *targets_dev, const float *means_dev,
                                const float *log_scales_dev, const float *probs_dev, int K,
                                int64_t n_sym, int Lp, const int16_t *sym_host,
                                uint8_t *out_host, size_t out_cap, size_t *out_len);

/* replaces `decode_logistic_mixture(..., in) -> int16 tensor` (torchac.cpp:385-421). */
int l3c_decode_logistic_mixture(const float *targets_dev, const float *means_dev,
                                const float *log_scales_dev, const float *probs_dev, int K,
                                int64_t n_sym, int Lp, const uint8_t *in_host, size_t in_len,
                                int16_t *sym_out_host);

/* ------------------------------------------------------------------------------------------
 * B. Range coder, batched: one warp per stream, any number of streams per launch
 *    (reference loop: torchac.cpp:174-207 encode, :325-373 decode, one stream per call)
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    const uint32_t *intervals;  /* dev: n_sym packed intervals                                  */
    uint8_t *out;               /* dev: output slot, 4-byte aligned, out_cap bytes              */
    uint32_t n_sym;
    uint32_t out_cap;           /* multiple of 4                                                */
} l3c_enc_stream_t;

/* out_len_dev[i] = bytes produced by stream i (if > out_cap the slot content is truncated). */
int l3c_ac_encode_streams(const l3c_enc_stream_t *streams_dev, int n_streams,
                          uint32_t *out_len_dev, void *stream);

typedef struct {
    const uint16_t *table;      /* dev: CDF rows for symbols [0, n_sym) of this stream           */
    const uint8_t *in;          /* dev: code bytes at ANY byte address (decoded in place from the  */
                                /*      container); the buffer must be readable 4 bytes past the end */
    uint8_t *sym_out;           /* dev: n_sym decoded symbols                                    */
    uint32_t *state;            /* dev: 4 words of coder state (low, high, value, bitpos);        */
                                /*      lets one stream be decoded in several launches (chunks)   */
    int64_t row_pitch;          /* entries between consecutive rows; 0 = one shared row           */
    uint32_t n_sym;             /* total symbols of the stream                                    */
    uint32_t in_len;
} l3c_dec_stream_t;

/* Decodes symbols [first, first+count) of every stream (count clipped to n_sym).  first == 0
 * initialises the coder state from the stream head, otherwise it is resumed from `state`.
 * L = number of symbols of the alphabet (<= 256). */
int l3c_ac_decode_streams(const l3c_dec_stream_t *streams_dev, int n_streams, int L,
                          uint32_t first, uint32_t count, void *stream);

/* Copies the first len_dev[i] bytes of every encoder slot to blob_dev + dst_off_dev[i] (byte
 * offsets: the final .l3c container layout, so ONE device->host copy returns a whole batch). */
int l3c_pack_streams(const l3c_enc_stream_t *streams_dev, const uint32_t *len_dev,
                     const uint64_t *dst_off_dev, int n_streams, uint8_t *blob_dev, void *stream);

/* intervals_dev[i] = lut_dev[sym_dev[i]] -- coding intervals when all pixels share one CDF row
 * (uniform prior of the coarsest scale, bitcoding.py:171-186). lut entry = c_low | (c_high-1)<<16. */
int l3c_lut_intervals(const uint8_t *sym_dev, const uint32_t *lut_dev, int64_t n,
                      uint32_t *intervals_dev, void *stream);

/* ------------------------------------------------------------------------------------------
 * C. Discretised-logistic-mixture head
 *    (reference: criterion/logistic_mixture.py:134-275, torchac.py:174-213,
 *     torchac_kernel.cu:20-76, bitcoding.py:297-323)
 * ---------------------------------------------------------------------------------------- */

/* Shared CDF row of the uniform prior (bitcoding.py:297-323): L+1 uint16 entries written to
 * row_host (host-side helper, no device work). */
int l3c_uniform_cdf_row(int L, uint16_t *row_host);

/* Per-symbol coding intervals for all C channels of one scale of a batch.
 *   l_dev      f32 NHWC [N][HW][Kp], Kp = (rgb ? 4 : 3) * C * K, channel index p*C*K + c*K + k
 *   sym_dev    uint8 [N][C][HW]  symbols being coded (for RGB they are also the values that
 *              enter the mean coupling, logistic_mixture.py:262-272)
 *   targets_dev f32 [L+1] bin edges (coders_helpers.py:44-46)
 *   intervals_dev uint32 [N][C][HW] out */
int l3c_dmll_intervals(const float *l_dev, const uint8_t *sym_dev, const float *targets_dev,
                       int N, int HW, int C, int K, int L, int rgb, uint32_t *intervals_dev,
                       void *stream);

/* uint16 CDF rows (pitch = 32 for L <= 32, else 256 entries) of channel c -- or of all C channels
 * when c < 0 (non-RGB scales: channels are conditionally independent) -- for pixels
 * [pix0, pix0+npix) of every image: table_dev[((n*C + c)*HW + p) * pitch + l].  For rgb && c > 0
 * the already decoded channels are read from sym_dev. */
int l3c_dmll_build_table(const float *l_dev, const uint8_t *sym_dev, const float *targets_dev,
                         int N, int HW, int C, int K, int L, int rgb, int c, int pix0, int npix,
                         uint16_t *table_dev, int pitch, void *stream);

/* Tiled stream order (the codec's throughput mode): every H x W symbol plane is cut into tiles of th x tw
 * (smaller at the right / bottom edge), tiles in row-major tile order, symbols row-major inside a tile; each
 * tile is one coded stream.  l3c_dmll_build_table_tiled = l3c_dmll_build_table for ALL pixels with rows and
 * sym_dev indexed in tile order (l_dev stays raster NHWC).  l3c_reorder_tiles converts `planes` planes of
 * 1- or 4-byte elements between raster and tile order (to_tiles != 0: raster -> tiles). */
int l3c_dmll_build_table_tiled(const float *l_dev, const uint8_t *sym_dev, const float *targets_dev,
                               int N, int H, int W, int C, int K, int L, int rgb, int c, int th, int tw,
                               uint16_t *table_dev, int pitch, void *stream);
int l3c_reorder_tiles(const void *src_dev, void *dst_dev, int elem_bytes, int planes, int H, int W, int th,
                      int tw, int to_tiles, void *stream);

/* Negative log-likelihood (nats) summed per image: nll_dev f64 [N] (overwritten).
 * target value of symbol s is values_dev[s] (logistic_mixture.py:146-207). */
int l3c_dmll_nll(const float *l_dev, const uint8_t *sym_dev, const float *values_dev,
                 int N, int HW, int C, int K, int L, int rgb, float x_min, float x_max,
                 double *nll_dev, float *nll_map_dev /* f32 [N][C][HW] per-sub-pixel nats, or NULL */,
                 void *stream);

/* Reference-shaped per-channel mixture parameters (CDFOut, logistic_mixture.py:61-65,134-141,
 * 248-275): softmax(pi), mu (incl. RGB coupling from x_dec_dev f32 [N][C][HW]) and clamped
 * log sigma, each f32 [N][K][HW]. */
int l3c_dmll_channel_params(const float *l_dev, const float *x_dec_dev, int N, int HW, int C, int K,
                            int rgb, int c, float *pi_dev, float *mu_dev, float *log_scales_dev,
                            void *stream);

/* Sample every sub-pixel from its mixture (logistic_mixture.py:277-323; used by test.py --sample): Gumbel-max
 * choice of the component from u_sel_dev f32 [N][C][K][HW], inverse-CDF sample of its logistic from u_x_dev f32
 * [N][C][HW] (both uniform in [1e-5, 1 - 1e-5], supplied by the caller), RGB: means coupled through the
 * coefficients of the chosen components and values clamped to [0, 255].  x_dev f32 [N][C][HW] out. */
int l3c_dmll_sample(const float *l_dev, const float *u_sel_dev, const float *u_x_dev, int N, int HW, int C, int K,
                    int rgb, float *x_dev, void *stream);

/* ------------------------------------------------------------------------------------------
 * D. Convolution stack (reference: modules/net.py:89-184, edsr.py:52-119, head.py:26-59,
 *    prob_clf.py:29-74, pytorch_ext.py:57-61; all nn.Conv2d fp32)
 * ---------------------------------------------------------------------------------------- */

#define L3C_CONV_RELU 1u            /* y = max(y, 0) after bias                                  */
#define L3C_CONV_PIXEL_SHUFFLE2 2u  /* out[n,2h+i,2w+j,c] = y[n,h,w,4c+2i+j] (edsr.py:92-101)     */
#define L3C_CONV_ROUND_TF32 4u      /* (TF32 mode) round y itself to TF32 (cvt.rna): y only feeds tensor-core convs */

#define L3C_PREC_FP32 0             /* CUDA-core FFMA, fp32 throughout (bit-faithful ordering)    */
#define L3C_PREC_TF32 1             /* tcgen05 kind::tf32 on operands pre-rounded to TF32 (RN), fp32 accumulate in TMEM */
#define L3C_PREC_F16 2              /* tcgen05 kind::f16 on FP16 operand images (RN, saturating; the same 10 mantissa */
                                    /* bits as TF32-RN), fp32 accumulate in TMEM, fp32 residual stream               */
#define L3C_PREC_F16X2 3            /* strict tensor-core mode: every operand is the sum of two FP16 numbers, hi = fp16(x) */
                                    /* and lo = fp16((x - hi) * 2^11); a product is hi*hi + (hi*lo + lo*hi) / 2^11 = three   */
                                    /* tcgen05.mma into two TMEM accumulators (the error-compensated split SURVEY section 7 */
                                    /* asks for, with FP16 instead of TF32 pieces): ~2^-22 relative, i.e. fp32-class results  */
                                    /* at several times the CUDA-core rate.  Operand images are [..][2*C] FP16: hi planes in  */
                                    /* channels [0,C), lo planes in [C,2C); 3x3/s1/Cin=64 and 1x1 layers                      */

typedef struct {
    const float *x;        /* dev NHWC [N][H][W][x_pitch] fp32, channels [0,Cin) read (FP32 / TF32 kernels)     */
    const float *w;        /* dev [KH][KW][Cin][cout_pad] (FP32) or the TF32 operand image (TF32)                */
    const float *bias;     /* dev [cout_pad]                                                       */
    const float *residual; /* dev NHWC like the output (same pitch/offset), or NULL: y += res      */
    float *y;              /* dev NHWC [N][Ho][Wo][y_pitch], channels [y_coff, y_coff+Cout) written; may be NULL  */
                           /* in F16 mode when only y_h is wanted                                                */
    float *y_tf32;         /* optional second output, same layout as y: the result rounded to TF32     */
                           /* (round-to-nearest) = operand image for a following TF32 tensor-core conv; NULL ok */
    void *y_h;             /* optional output, same element pitch/offset as y: the result as FP16 (RN, saturating) */
                           /* = operand image for a following F16 tensor-core conv; written by every kernel; NULL ok */
    const void *x_h;       /* F16 tensor-core kernels: FP16 operand image of the input, NHWC [N][H][W][x_pitch]    */
    const void *w_h;       /* F16 tensor-core kernels: FP16 weight image, 3x3: [9 taps][cout_pad][64],             */
                           /* 1x1: [Cin/64][cout_pad][64]                                                         */
    int N, H, W, Cin, x_pitch;
    int Cout, cout_pad, y_pitch, y_coff;
    int ksize, stride, dilation;   /* padding = ksize/2 if dilation==1 else dilation               */
    unsigned flags;
    int precision;
    int yh_pitch;          /* element pitch of y_h when it differs from y_pitch (0 = same): split images are 2x as wide */
    int yh_lo_off;         /* > 0: y_h is a SPLIT image (L3C_PREC_F16X2): lo plane yh_lo_off elements after the hi plane */
} l3c_conv_t;

int l3c_conv2d(const l3c_conv_t *p, void *stream);

/* The DMLL head fused into the 1x1 `lin` conv of the probability classifier (encode side; reference:
 * prob_clf.py:71-74 followed by logistic_mixture.py:248-275 and torchac_kernel.cu:26-76): the Kp parameters of
 * a pixel never reach HBM -- the conv's epilogue evaluates the two CDF bounds of every coded symbol directly
 * from the accumulator.  x_h: FP16 operand image NHWC [N][HW][Cin] (the atrous concat); w_h: FP16 weight image
 * [Cin/64][cout_pad][64]; bias f32 [cout_pad]; sym_dev uint8 [N][C][HW]; targets_dev f32 [L+1];
 * intervals_dev uint32 [N][C][HW] out, bit-identical to l3c_conv2d (F16) + l3c_dmll_intervals. */
int l3c_lin_dmll_intervals(const void *x_h, const void *w_h, const float *bias, const uint8_t *sym_dev,
                           const float *targets_dev, int N, int HW, int Cin, int C, int K, int L, int rgb,
                           uint32_t *intervals_dev, void *stream);

/* sub_rgb_mean + first MeanShift of RGBHead as one per-pixel 3x3 affine pair
 * (multiscale_network.py:181-183, head.py:31-33): img uint8 [N][3][HW] planes ->
 * x_sub f32 NHWC [N][HW][3] (= A1 img + b1)  and  t f32 NHWC [N][HW][4] (= A2 x_sub + b2, ch 3 = 0).
 * A1,b1,A2,b2: dev f32 [9],[3],[9],[3] (the two 1x1 conv parameter sets). Either output may be NULL. */
int l3c_rgb_prep(const uint8_t *img_dev, const float *A1, const float *b1, const float *A2,
                 const float *b2, int N, int HW, float *xsub_dev, float *t_dev, void *stream);

/* F16 path of the RGB head (head.py:26-59): the two MeanShift affines of l3c_rgb_prep followed by an im2col of the
 * 3x3 neighbourhood: out_h_dev FP16 [N][HW][64], element (ky*3+kx)*3 + c = normalised channel c at tap (ky,kx)
 * (0 outside the image), elements 27..63 zero -- the 3 -> 64 conv then runs as a K = 64 GEMM on the tensor cores. */
int l3c_rgb_im2col_f16(const uint8_t *img_dev, const float *A1, const float *b1, const float *A2, const float *b2,
                       int N, int H, int W, void *out_h_dev, void *stream);

/* L3C_PREC_F16X2: fp32 [n_px][C] -> split operand image FP16 [n_px][2*C] (hi = fp16(x) in channels [0,C),
 * lo = fp16((x - hi) * 2^11) in [C,2C)); C % 4 == 0.  For the outputs of the CUDA-core layers of that mode (no
 * reference counterpart: the reference's convolutions are cuDNN fp32, pytorch_models/modules/edsr.py:33-38). */
int l3c_split_f16x2(const float *x_dev, long long n_px, int C, void *out_h_dev, void *stream);

/* to_q 1x1 conv + hard quantiser (net.py:116-148, quantizer.py:62-90):
 * sym = argmin_l (q - level_l)^2 (first minimum), bn_q = levels[sym].
 * f_dev NHWC [N][HW][Cf]; w_dev [Cf][C]; bias [C]; levels [L];
 * sym_dev uint8 [N][C][HW] planes; bnq_dev f32 NHWC [N][HW][C]. */
int l3c_quantize_head(const float *f_dev, const float *w_dev, const float *bias_dev,
                      const float *levels_dev, int N, int HW, int Cf, int C, int L,
                      uint8_t *sym_dev, float *bnq_dev, void *stream);

/* bn = values[sym] for symbol planes -> NHWC f32 [N][HW][C] minus optional per-channel shift
 * (decode side of quantizer.py:44-47 / SURVEY finding 1, and the RGB baselines' x - rgb_mean). */
int l3c_symbols_to_values(const uint8_t *sym_dev, const float *values_dev, const float *shift_dev,
                          int N, int HW, int C, int L, float *out_dev, void *stream);

/* Pillow-compatible bicubic x0.5 on uint8 planes [N][3][H][W] -> [N][3][H/2][W/2]
 * (dataloaders/images_loader.py:277-293 via PIL.Image.resize(BICUBIC); RGB baselines only). */
int l3c_bicubic_half_u8(const uint8_t *in_dev, int N, int H, int W, uint8_t *out_dev, void *stream);

/* RGB scale of a decode, channel-pipelined, as ONE host call (the runtime loop of codec.py moved into the
 * library: ~400 launches and event operations per decode): chunk j of channel c-1 decoded -> rows of chunk j of
 * channel c built on bld_streams[c] -> channel c's decoders resume on dec_streams[c] (state carried in the
 * descriptors).  desc_dev_per_channel[c]: device array of N stream descriptors of channel c.  Ordered after
 * `cur_stream`'s queue; `cur_stream` waits for all of it.  Reference: the strictly serial R -> G -> B loop of
 * bitcoding.py:199-237. */
int l3c_decode_rgb_pipelined(const float *l_dev, uint8_t *sym_dev, const float *targets_dev, int N, int HW, int K,
                             int L, uint16_t *table_dev, int pitch,
                             const l3c_dec_stream_t *const *desc_dev_per_channel, int chunk_px, void *cur_stream,
                             void *const *bld_streams, void *const *dec_streams);

/* ------------------------------------------------------------------------------------------
 * E. Execution resources
 * ---------------------------------------------------------------------------------------- */

/* Streams confined to two disjoint groups of SMs of the current device (driver green contexts):
 * `n_a` streams (high priority) whose kernels only run on a group of about `sm_a` SMs (the driver
 * rounds up to its granularity, 8 on sm_90+), `n_b` streams on the remaining SMs (all but the last of
 * them high priority, the last one default priority).  Used by the
 * pipelined RGB decode (reference: the strictly serial R -> G -> B loop of bitcoding.py:199-237) so
 * that the latency-bound decoder warps do not share SMs with the CDF-row builders.  The streams
 * belong to the primary context (same memory, events interoperate) and live until process exit;
 * repeated calls with the same `sm_a` return the same streams.  *sm_a_out / *sm_b_out (optional)
 * receive the actual group sizes.  L3C_EUNSUPPORTED if the driver cannot partition: the caller then
 * uses ordinary streams (slower, same results). */
int l3c_partition_streams(int sm_a, int n_a, void **streams_a, int n_b, void **streams_b,
                          int *sm_a_out, int *sm_b_out);
/* the same with explicit counts of high- and default-priority streams on the remaining SMs */
int l3c_partition_streams2(int sm_a, int n_a, void **streams_a, int n_b_high, void **streams_b_high,
                           int n_b_low, void **streams_b_low, int *sm_a_out, int *sm_b_out);

#ifdef __cplusplus
}
#endif
#endif /* L3C_B200_H_ */
