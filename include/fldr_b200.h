/*
 * fldr_b200.h - C ABI of the B200-native (sm_100a) replacement for fLDR-VFI's two CuPy-JIT ops.
 *
 * Drop-in boundary.  Each entry point cites the reference interface it
 * replaces (paths relative to the reference checkout):
 *
 *   fldr_splat_fwd   <- softSplat.py:320-352  FunctionSoftsplat(tenInput, tenFlow, tenMetric, strType)
 *                       softSplat.py:222-259  _FunctionSoftsplat.forward + kernel_Softsplat_updateOutput (12-52)
 *   fldr_splat_bwd   <- softSplat.py:262-317  _FunctionSoftsplat.backward + kernel_Softsplat_updateGradInput (54-98)
 *                       + kernel_Softsplat_updateGradFlow (100-158), plus the autograd of the torch
 *                       glue at 320-352 (metric / normaliser gradients)
 *   fldr_corr81_fwd  <- OpticalFlow/correlation.py:296-348  _FunctionCorrelation.forward
 *                       (kernel_Correlation_rearrange 17-42 x2 + kernel_Correlation_updateOutput 44-112)
 *   fldr_corr81_bwd  <- OpticalFlow/correlation.py:353-409  _FunctionCorrelation.backward
 *                       (kernel_Correlation_updateGradFirst 114-176, updateGradSecond 178-242)
 *   fldr_bwarp_fwd, fldr_warp_metric_fwd  <- fLDRnet.py:546-581 DCTVFInet.bwarp and the metric at 442-446
 *                       (the step just before the image splat)
 *   fldr_occ_blend_fwd  <- fLDRnet.py:510-524 occlusion softmax + six-way blend (the last step; 8f rank 2)
 *   fldr_pca_features_fwd  <- pca_comp.py:473-528 to_pca_diff, called at fLDRnet.py:146 (the first device step; 8f rank 4)
 *
 * Conventions
 *   - plain C: raw device pointers, explicit element strides, explicit sizes, a CUDA stream handle.
 *   - all tensors are fp32 and live on the current CUDA device; nothing is allocated or freed here,
 *     scratch comes from the caller (`ws`, size from the matching *_workspace_bytes call).
 *   - every call is asynchronous on `stream`; no host synchronisation.
 *   - return value: FLDR_OK (0) or a negative fldr_status.  No exceptions cross this boundary.
 *   - strides are in ELEMENTS, for logical shape [N, C, H, W] (NCHW); outputs are written
 *     NCHW-contiguous, exactly as the reference allocates them (softSplat.py:234, correlation.py:305).
 */
#ifndef FLDR_B200_H
#define FLDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLDR_B200_ABI_VERSION 1

/* cudaStream_t without pulling CUDA headers into the consumer (cgo / JNI / ctypes friendly). */
typedef struct CUstream_st* fldr_stream_t;

typedef enum fldr_status {
    FLDR_OK = 0,
    FLDR_ERR_INVALID_ARGUMENT = -1,   /* null pointer, non-positive size, unknown mode, bad stride */
    FLDR_ERR_WORKSPACE_TOO_SMALL = -2,
    FLDR_ERR_CUDA = -3,               /* a CUDA runtime call failed; see fldr_last_cuda_error() */
    FLDR_ERR_UNSUPPORTED = -4,        /* e.g. metric missing for 'linear' (softSplat.py:328) */
    FLDR_ERR_NO_DEVICE = -5
} fldr_status;

/* strType of FunctionSoftsplat (softSplat.py:322) plus the raw summation splat of _FunctionSoftsplat. */
typedef enum fldr_splat_mode {
    FLDR_SPLAT_SUMMATION = 0,   /* y = (S - 0.5) * 2                      (softSplat.py:349 is unconditional) */
    FLDR_SPLAT_AVERAGE = 1,     /* A = [x, 1]                             (325) */
    FLDR_SPLAT_LINEAR = 2,      /* A = [x * z, z]        metric required  (328) */
    FLDR_SPLAT_SOFTMAX = 3,     /* A = [(x+1)/2 * e^z, e^z]; metric NULL -> [(x+1)/2, 1]  (334-338) */
    FLDR_SPLAT_RAW = 4          /* y = S : _FunctionSoftsplat.apply(input, flow), no pre/post-processing */
} fldr_splat_mode;

int fldr_abi_version(void);
const char* fldr_status_string(int status);
/* cudaError_t of the most recent failing CUDA call made by this library on the calling thread (0 if none). */
int fldr_last_cuda_error(void);
/*
 * Tuning / diagnostic switches (process-wide; initial values come from FLDR_<NAME> environment variables):
 *   "splat_tma"      1 (default): the scatter pass stages its inputs with TMA box loads when the views allow it
 *                    (16-byte aligned rows, unit pixel stride); 0 = always the plain-load scatter kernel
 *   "splat_fused_max" frames with at most this many accumulator float4s (N * ceil((C+1)/4) * H * W, default 40000)
 *                    run zero + scatter + normalise as ONE cooperative launch; 0 disables
 *   "corr_th"        tile height of the correlation forward kernel: 0 automatic, 8 or 16 forced
 *   "splat_pf_rows"  accumulator rows prefetched into L2 ahead of the scatter (0 = default: 4 with splat_snake 0, none with
 *                    splat_snake 1; negative = off)
 *   "corr_bwd_rows"  1 (default): correlation backward with three output rows per thread and gradOut streamed through a TMA
 *                    ring when the views allow it and C <= 32; 2 = for every C; 0 = always the 4-row tile kernel
 *   "splat_snake"    1 (default): the three passes of the whole-frame splat run in alternating row order (zero fill front to back
 *                    with L2-allocating stores, scatter back to front, normalise front to back) so that each starts on the
 *                    accumulator lines its predecessor left in L2; 0 = cudaMemsetAsync + every pass front to back
 * Results are identical (within the summation-order tolerance) for every setting.
 */
int fldr_set_option(const char* name, int value);
int fldr_get_option(const char* name);

/* ---------------------------------------------------------------- splat ---------------------------------- */

/* Scratch needed by fldr_splat_fwd for this shape (accumulator in pixel-interleaved layout). */
size_t fldr_splat_fwd_workspace_bytes(int mode, int N, int C, int H, int W);

/*
 * Forward splat with the mode's pre/post-processing fused in.
 *   in      [N,C,H,W]  strides in_strides[4]
 *   flow    [N,2,H,W]  strides flow_strides[4]      (channel 0 = x displacement, 1 = y; softSplat.py:23-24)
 *   metric  [N,1,H,W]  strides metric_strides[4], or NULL (allowed for SOFTMAX / AVERAGE / SUMMATION / RAW)
 *   out     [N,C,H,W]  contiguous.  Holes (normaliser == 0) come out as -1 (softSplat.py:346-349).
 *   norm    [N,1,H,W]  contiguous or NULL: raw normaliser S[:, C] before the 0 -> 1 fix-up, saved for
 *                      fldr_splat_bwd.  Ignored (may be NULL) for SUMMATION / RAW.
 * Pixels whose target coordinate is not finite are skipped (the reference device-asserts, 25-26).
 */
int fldr_splat_fwd(int mode,
                   const float* in, const int64_t* in_strides,
                   const float* flow, const int64_t* flow_strides,
                   const float* metric, const int64_t* metric_strides,
                   float* out, float* norm,
                   int N, int C, int H, int W,
                   void* ws, size_t ws_bytes, fldr_stream_t stream);

/*
 * Debug aid for the reference's device assert on non-finite flow (softSplat.py:25-26, which kills the CUDA context).  Here such
 * a pixel is skipped; when a device word is registered with this call (per device; NULL unregisters), every forward splat
 * that meets a non-finite target coordinate stores 1 into it.  The caller zeroes and reads the word (a synchronisation, so
 * the Python wrapper only does it when FLDR_B200_CHECK_FLOW=1).
 */
int fldr_splat_set_nonfinite_flag(unsigned int* device_flag);

size_t fldr_splat_bwd_workspace_bytes(int mode, int N, int C, int H, int W);

/*
 * Backward of fldr_splat_fwd.  Any of grad_in / grad_flow / grad_metric may be NULL (gradient not
 * requested: softSplat.py:276-277 needs_input_grad).
 *   out, norm   the forward results (contiguous); unused (may be NULL) for SUMMATION / RAW
 *   grad_out    [N,C,H,W]  strides grad_out_strides[4]
 *   grad_in     [N,C,H,W]  contiguous     grad_flow [N,2,H,W] contiguous     grad_metric [N,1,H,W] contiguous
 */
int fldr_splat_bwd(int mode,
                   const float* in, const int64_t* in_strides,
                   const float* flow, const int64_t* flow_strides,
                   const float* metric, const int64_t* metric_strides,
                   const float* out, const float* norm,
                   const float* grad_out, const int64_t* grad_out_strides,
                   float* grad_in, float* grad_flow, float* grad_metric,
                   int N, int C, int H, int W,
                   void* ws, size_t ws_bytes, fldr_stream_t stream);

/* ------------------------------------------------------------ correlation -------------------------------- */

size_t fldr_corr81_fwd_workspace_bytes(int B, int C, int H, int W);

/*
 * 81-channel cost volume, displacements dy,dx in [-4,4], zero padding, mean over channels:
 *   out[b,(dy+4)*9+(dx+4),y,x] = (1/C) sum_c first[b,c,y,x] * second[b,c,y+dy,x+dx]
 *   first, second [B,C,H,W] with strides; out [B,81,H,W] contiguous.
 */
int fldr_corr81_fwd(const float* first, const int64_t* first_strides,
                    const float* second, const int64_t* second_strides,
                    float* out, int B, int C, int H, int W,
                    void* ws, size_t ws_bytes, fldr_stream_t stream);

/*
 * Next row (PWCNet.py:146-160): the same cost volume with PWC-Net's leaky_relu(., negative_slope)
 * applied in the store epilogue (negative_slope = 1 disables it) and written with `out_sample_stride` elements
 * between samples (>= 81*H*W; 0 means dense), so it can land directly in channels 0..80 of the decoder's
 * concatenation buffer [B, 81 + ..., H, W] (PWCNet.py:160) - no separate activation pass, no copy by torch.cat.
 */
int fldr_corr81_fwd_act(const float* first, const int64_t* first_strides,
                        const float* second, const int64_t* second_strides,
                        float* out, int64_t out_sample_stride, float negative_slope,
                        int B, int C, int H, int W, fldr_stream_t stream);

size_t fldr_corr81_bwd_workspace_bytes(int B, int C, int H, int W);

/* grad_first / grad_second [B,C,H,W] contiguous, either may be NULL (correlation.py:358-361). */
int fldr_corr81_bwd(const float* first, const int64_t* first_strides,
                    const float* second, const int64_t* second_strides,
                    const float* grad_out, const int64_t* grad_out_strides,
                    float* grad_first, float* grad_second,
                    int B, int C, int H, int W,
                    void* ws, size_t ws_bytes, fldr_stream_t stream);

/* ------------------------------------------------- backward warp + splat metric (next row) ---- */

/*
 * DCTVFInet.bwarp(x, flo, withmask) of fLDRnet.py:546-581: out[n,c,y,x] = bilinear sample of x[n,c] at
 *   ix = ((2*(x+u)/max(W-1,1) - 1 + 1) * W - 1) / 2   (the reference's normalisation + grid_sample's
 *   align_corners=False default, zero padding), iy likewise, times the mask
 *   [sum of in-frame bilinear weights >= 0.999] when with_mask != 0 (569-578).
 *   x [N,C,H,W] with strides (non-negative H/W strides), flow [N,2,H,W] with strides, out [N,C,H,W] contiguous.
 * convention 0: as above.  convention 1: PWC-Net's Backward (OpticalFlow/PWCNet.py:116-143), the warp of the second
 *   feature map in front of every correlation: g = linspace(-1,1,W)[x] + u / ((W-1)/2), same
 *   un-normalisation, mask = [in-frame weight > 0.999] always applied (with_mask is ignored).
 */
int fldr_bwarp_fwd(const float* x, const int64_t* x_strides,
                   const float* flow, const int64_t* flow_strides,
                   float* out, int N, int C, int H, int W, int with_mask, int convention,
                   fldr_stream_t stream);

/*
 * Backward of fldr_bwarp_fwd (what autograd derives for fLDRnet.py:556-578 / PWCNet.py:134-143; the mask and floor()
 * carry no gradient).  grad_out [N,C,H,W] with strides; grad_x [N,C,H,W] contiguous (zero-filled here, then
 * accumulated) and grad_flow [N,2,H,W] contiguous; either may be NULL.
 */
int fldr_bwarp_bwd(const float* x, const int64_t* x_strides,
                   const float* flow, const int64_t* flow_strides,
                   const float* grad_out, const int64_t* grad_out_strides,
                   float* grad_x, float* grad_flow,
                   int N, int C, int H, int W, int with_mask, int convention, fldr_stream_t stream);

/*
 * The splat metric of fLDRnet.py:442-446 in one pass, without materialising the warped image:
 *   out[n,0,y,x] = (1/C) * sum_c alpha * | ref[n,c,y,x] - bwarp(src, flow)[n,c,y,x] |
 *   ref, src [N,C,H,W] with strides; out [N,1,H,W] contiguous; alpha = z_alpha[i] rounded to fp32.
 */
int fldr_warp_metric_fwd(const float* ref, const int64_t* ref_strides,
                         const float* src, const int64_t* src_strides,
                         const float* flow, const int64_t* flow_strides,
                         float alpha, float* out, int N, int C, int H, int W, int with_mask,
                         fldr_stream_t stream);

/* ------------------------------------- occlusion softmax + image synthesis (next row) ---------- */

/*
 * fLDRnet.py:510-524 in one pass, in float64 like the reference (its temperature is a float64 Parameter):
 *   occ = softmax(logits[:, 0:6] / T, dim=1);  a_k = (k even ? 1-t : t) * occ_k
 *   out = (a0*img0 + a1*img1 + a2*img2 + a3*img3 + a4*img4 + a5*img5) / (a0 + ... + a5)
 *   logits      [N,>=6,H,W] float32 with strides (channels 0..5 are read)
 *   images      host array of 6 device pointers, each [N,C,H,W] float32: warped_img0, warped_img1, im0_tot, im1_tot,
 *               x0, x1 (the order of lines 518-521); image_strides = host array of 6 x 4 element strides
 *   t_value     device, float32, one per sample, t_stride elements apart;  temperature  device, one float64 (T_param)
 *   out         [N,C,H,W] float64 contiguous;  occ0  [N,1,H,W] float64 contiguous or NULL (occ[:, 0:1], line 512)
 */
int fldr_occ_blend_fwd(const float* logits, const int64_t* logits_strides,
                       const float* const* images, const int64_t* image_strides,
                       const float* t_value, int64_t t_stride, const double* temperature,
                       double* out, double* occ0, int N, int C, int H, int W, fldr_stream_t stream);

/* ------------------------------------- block-PCA feature extraction (next row) -------------------------- */

/* Scratch of fldr_pca_features_fwd: the min / max words (+ the float64 intermediate when the output is float32). */
size_t fldr_pca_features_workspace_bytes(int chan, int H, int W, int ncomp, int out_is_f32);

/*
 * pca_comp.py:473-528 `to_pca_diff(im, params, args, mean, EV, mean_vec)` with params.wiS = 8, the first device step of every
 * fLDRnet forward (fLDRnet.py:146), in float64 like the reference:
 *   t[c*ncomp + k, yb, xb] = sum_j (im[c, 8 yb + j / 8, 8 xb + j % 8] - mean[j]) * EV[k][j]     ( / mean_vec[k] if given )
 *   out = ((t - min t) / (max t - min t)) * 2 - 1              (min / max over the whole tensor, pca_comp.py:521-526)
 *   im        [chan, H, W] float32, strides im_strides[3] in elements (unit pixel stride); H, W multiples of 8
 *             (anything else returns FLDR_ERR_INVALID_ARGUMENT where the reference raises, 486-487)
 *   mean      [64] float64;  ev [ncomp <= 16][64] float64, ev_row_stride elements between rows (EV.permute(1,0) of 507 is
 *             folded into the indexing);  mean_vec [ncomp] float64 or NULL (args.mean_vector_norm off)
 *   out       [chan * ncomp, H/8, W/8] contiguous, float64 (out_is_f32 = 0: what the reference returns) or float32
 *             (out_is_f32 = 1: the `.float()` the caller applies at fLDRnet.py:146, fused)
 */
int fldr_pca_features_fwd(const float* im, const int64_t* im_strides, const double* mean, const double* ev,
                          int64_t ev_row_stride, const double* mean_vec, void* out, int out_is_f32,
                          int chan, int H, int W, int ncomp, void* ws, size_t ws_bytes, fldr_stream_t stream);

/* ------------------------------------- input pyramid (next row, SURVEY 8f-4, second half) --------------------------- */

/*
 * main.py:855-856 (test) / 562-563 (train): level i > 0 of `input_gpu` is
 *   F.interpolate(frames, scale_factor = scales[0] / scales[i], mode = 'bicubic', align_corners = args.align_cornerse)
 * of the full-resolution (padded) frames, computed by the reference on the CPU and copied to the device level by level.
 * Here the frames are on the device and every level is written by one call (one launch for the shipped presets: factors
 * 1/2, 1/4, ... 1/32 with align_corners = 0; any other factor list / align_corners = 1 takes one generic launch per level).
 * Arithmetic follows ATen's upsample_bicubic2d (A = -0.75, taps clamped to the frame, horizontal sums first).
 *   frames         [planes, H, W] float32 (planes = B * C * T: interpolation is per plane), unit pixel stride,
 *                  row_stride / plane_stride in elements
 *   scale_factors  host array of n_levels doubles (scales[0] / scales[i]); level i has floor(H * f_i) x floor(W * f_i) pixels
 *   out_levels     host array of n_levels device pointers, level i contiguous [planes, floor(H f_i), floor(W f_i)] float32
 * No workspace.  Empty pyramids (n_levels = 0) and zero planes succeed without a launch.
 */
int fldr_bicubic_pyramid_fwd(const float* frames, int64_t plane_stride, int64_t row_stride, int planes, int H, int W,
                             int n_levels, const double* scale_factors, int align_corners, float* const* out_levels,
                             fldr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FLDR_B200_H */
