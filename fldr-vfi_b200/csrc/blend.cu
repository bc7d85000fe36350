// Occlusion softmax + six-way image synthesis of fLDRnet, sm_100a.
//
// Replaces fLDRnet.py:510-524 - softmax over six logits divided by a float64 temperature, six weight maps scaled by
// t / (1-t), three pairs of weighted images, a divisor and a divide: about thirty torch kernels over float64
// [N,3,H,W] temporaries (226 MB each at 4K) - by ONE elementwise kernel.  The reference computes all of this in float64
// (T_param is a double Parameter, fLDRnet.py:357, and promotes everything downstream); so does the kernel, term by
// term and in the reference's association order, with __d*_rn intrinsics so nothing is contracted:
//   e_k = logit_k / T;  occ_k = exp(e_k - max e) / sum_j exp(e_j - max e)                                  (511)
//   a_k = (k even ? (double)(1.f - t) : (double)t) * occ_k
//   divisor = (((a0 + a1) + a2) + a3) + (a4 + a5)                                                          (517, 522)
//   out = ((a0*w0 + a1*w1) + (a2*i0 + a3*i1)) + (a4*x0 + a5*x1);  out /= divisor                           (518-524)
// Two places trade the reference's twelve float64 divisions per pixel for cheaper, equally accurate forms (a float64
// division is ~35 instructions; with all of them exact the kernel was instruction-bound at 409 us for the 4K frame):
// logit / T is skipped when T == 1 (exact; the shipped checkpoint) and the softmax normalisation multiplies by one
// exact reciprocal of the sum (<= 1 ulp per weight, far inside the 2e-14 parity bound; tests/test_gpu_blend.py).
// Inputs float32 (logits, six images), output float64 as in the reference.  Algorithmic bytes per pixel:
// 6*4 + 6*C*4 + C*8 (= 120 B at C = 3): HBM-bound.
#include "common.cuh"

namespace fldr {

struct Img6 {
    const float* p[6];
    long long sn[6], sc[6], sh[6], sw[6];
};

template <int CT>      // compile-time channel count, 0 = run-time loop
__global__ void __launch_bounds__(128) occ_blend_kernel(View4 logits, Img6 im, const float* __restrict__ t_value,
                                                        long long t_stride, const double* __restrict__ temperature,
                                                        double* __restrict__ out, double* __restrict__ occ0, int C_,
                                                        int H, int W, int y_base) {
    const int C = CT ? CT : C_;
    const int x = blockIdx.x * 128 + threadIdx.x, y = y_base + blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const long long HW = (long long)H * W, idx = (long long)y * W + x;
    const double T = __ldg(temperature);
    const float t = __ldg(t_value + n * t_stride);
    const float* lp = logits.p + n * logits.sn + y * logits.sh + x * logits.sw;
    double e[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) e[k] = (double)__ldcs(lp + k * logits.sc);
    if (T != 1.0) {                                   // uniform branch
#pragma unroll
        for (int k = 0; k < 6; ++k) e[k] = __ddiv_rn(e[k], T);
    }
    double m = e[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) m = fmax(m, e[k]);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        e[k] = exp(__dsub_rn(e[k], m));
        s = k == 0 ? e[0] : __dadd_rn(s, e[k]);
    }
    const double omt = (double)__fsub_rn(1.0f, t), td = (double)t;
    const double rs = __ddiv_rn(1.0, s);
    double a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = __dmul_rn((k & 1) ? td : omt, __dmul_rn(e[k], rs));
    if (occ0) __stcs(occ0 + n * HW + idx, __ddiv_rn(e[0], s));
    const double divisor = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), a[2]), a[3]), __dadd_rn(a[4], a[5]));
    const float* ip[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) ip[k] = im.p[k] + n * im.sn[k] + y * im.sh[k] + x * im.sw[k];
    double* op = out + (long long)n * C * HW + idx;
#pragma unroll
    for (int c = 0; c < (CT ? CT : 1); ++c) {
        for (int cc = CT ? c : 0; cc < (CT ? c + 1 : C); ++cc) {       // unrolled when CT != 0, a plain loop otherwise
            double v[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = (double)__ldcs(ip[k] + cc * im.sc[k]);
            double o = __dadd_rn(__dmul_rn(a[0], v[0]), __dmul_rn(a[1], v[1]));
            o = __dadd_rn(o, __dadd_rn(__dmul_rn(a[2], v[2]), __dmul_rn(a[3], v[3])));
            o = __dadd_rn(o, __dadd_rn(__dmul_rn(a[4], v[4]), __dmul_rn(a[5], v[5])));
            __stcs(op + cc * HW, __ddiv_rn(o, divisor));
        }
    }
}

}  // namespace fldr

using namespace fldr;

extern "C" int fldr_occ_blend_fwd(const float* logits, const int64_t* logits_strides, const float* const* images,
                                  const int64_t* image_strides, const float* t_value, int64_t t_stride,
                                  const double* temperature, double* out, double* occ0, int N, int C, int H, int W,
                                  fldr_stream_t stream) {
    if (!logits || !logits_strides || !images || !image_strides || !t_value || !temperature || !out)
        return FLDR_ERR_INVALID_ARGUMENT;
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return FLDR_ERR_INVALID_ARGUMENT;
    if ((long long)H * W >= (1ll << 31) || N > 65535) return FLDR_ERR_UNSUPPORTED;
    Img6 im;
    for (int k = 0; k < 6; ++k) {
        if (!images[k]) return FLDR_ERR_INVALID_ARGUMENT;
        im.p[k] = images[k];
        im.sn[k] = image_strides[4 * k]; im.sc[k] = image_strides[4 * k + 1];
        im.sh[k] = image_strides[4 * k + 2]; im.sw[k] = image_strides[4 * k + 3];
    }
    const View4 vl = make_view(logits, logits_strides);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    for (int y0 = 0; y0 < H; y0 += 65535) {
        const int rows = H - y0 < 65535 ? H - y0 : 65535;
        dim3 grid((unsigned)((W + 127) / 128), (unsigned)rows, (unsigned)N);
        if (C == 3) occ_blend_kernel<3><<<grid, 128, 0, s>>>(vl, im, t_value, t_stride, temperature, out, occ0, C, H, W, y0);
        else if (C == 1) occ_blend_kernel<1><<<grid, 128, 0, s>>>(vl, im, t_value, t_stride, temperature, out, occ0, C, H, W, y0);
        else occ_blend_kernel<0><<<grid, 128, 0, s>>>(vl, im, t_value, t_stride, temperature, out, occ0, C, H, W, y0);
        const int st = check_launch();
        if (st != FLDR_OK) return st;
    }
    return FLDR_OK;
}
