// Thin PyTorch C++ extension over the C ABI of include/fldr_b200.h (north_star: "a thin PyTorch C++ extension over a C-ABI").
// It holds no kernels and no arithmetic: it turns tensors into the raw pointers / strides / stream the library takes,
// allocates outputs with at::empty, keeps one grow-only scratch buffer per (device, stream), and returns tensors.  The
// reference-facing argument checks (exception types of softSplat.py / correlation.py) stay in the Python mirror; the
// autograd Functions there call these entry points.  The same library also serves ctypes / cgo / JNI consumers.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/fldr_b200.h"

namespace {

void check(int status) {
    if (status != FLDR_OK)
        throw std::runtime_error(std::string(fldr_status_string(status)) + " (status " + std::to_string(status) + ", cudaError " +
                                 std::to_string(fldr_last_cuda_error()) + ")");
}

struct Strides4 {
    int64_t s[4];
    explicit Strides4(const at::Tensor& t) { for (int i = 0; i < 4; ++i) s[i] = t.stride(i); }
};

// grow-only scratch per (device, stream): kernels of successive calls on a stream are ordered, so they can share it
at::Tensor workspace(size_t nbytes, const at::Tensor& like, cudaStream_t stream) {
    static std::mutex mu;
    static std::map<std::pair<int, void*>, at::Tensor> cache;
    if (nbytes < 16) nbytes = 16;
    auto opts = at::TensorOptions().dtype(at::kByte).device(like.device());
    if (at::cuda::currentStreamCaptureStatusMayInitCtx() != at::cuda::CaptureStatus::None) return at::empty({(int64_t)nbytes}, opts);
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair((int)like.get_device(), (void*)stream);
    auto it = cache.find(key);
    if (it == cache.end() || (size_t)it->second.numel() < nbytes) {
        cache[key] = at::empty({(int64_t)nbytes}, opts);
        return cache[key];
    }
    return it->second;
}

const float* fptr(const at::Tensor& t) { return t.data_ptr<float>(); }
float* fptr_mut(at::Tensor& t) { return t.data_ptr<float>(); }

}  // namespace

// fldr_splat_fwd: returns (out, norm or undefined)
std::tuple<at::Tensor, at::Tensor> splat_fwd(int64_t mode, const at::Tensor& in, const at::Tensor& flow, const c10::optional<at::Tensor>& metric,
                                             bool want_norm) {
    const c10::cuda::CUDAGuard guard(in.device());
    const int N = in.size(0), C = in.size(1), H = in.size(2), W = in.size(3);
    at::Tensor out = at::empty({N, C, H, W}, in.options());
    const bool has_norm = mode == FLDR_SPLAT_AVERAGE || mode == FLDR_SPLAT_LINEAR || mode == FLDR_SPLAT_SOFTMAX;
    at::Tensor norm;
    if (want_norm && has_norm) norm = at::empty({N, 1, H, W}, in.options());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(in.get_device()).stream();
    const size_t ws_bytes = fldr_splat_fwd_workspace_bytes((int)mode, N, C, H, W);
    at::Tensor ws = workspace(ws_bytes, in, stream);
    Strides4 si(in), sf(flow);
    at::Tensor m;
    if (metric.has_value()) m = metric->expand({N, 1, H, W});
    int64_t sm[4] = {0, 0, 0, 0};
    if (m.defined()) for (int i = 0; i < 4; ++i) sm[i] = m.stride(i);
    check(fldr_splat_fwd((int)mode, fptr(in), si.s, fptr(flow), sf.s, m.defined() ? fptr(m) : nullptr, m.defined() ? sm : nullptr,
                         fptr_mut(out), norm.defined() ? fptr_mut(norm) : nullptr, N, C, H, W, ws.data_ptr(), (size_t)ws.numel(),
                         reinterpret_cast<fldr_stream_t>(stream)));
    return std::make_tuple(out, norm);
}

// fldr_splat_bwd: returns (grad_in, grad_flow, grad_metric), undefined where not requested
std::tuple<at::Tensor, at::Tensor, at::Tensor> splat_bwd(int64_t mode, const at::Tensor& in, const at::Tensor& flow,
                                                          const c10::optional<at::Tensor>& metric, const c10::optional<at::Tensor>& out,
                                                          const c10::optional<at::Tensor>& norm, const at::Tensor& grad_out, bool need_in,
                                                          bool need_flow, bool need_metric) {
    const c10::cuda::CUDAGuard guard(in.device());
    const int N = in.size(0), C = in.size(1), H = in.size(2), W = in.size(3);
    at::Tensor gin, gfl, gme;
    if (need_in) gin = at::empty({N, C, H, W}, in.options());
    if (need_flow) gfl = at::empty({N, 2, H, W}, in.options());
    if (need_metric) gme = at::empty({N, 1, H, W}, in.options());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(in.get_device()).stream();
    Strides4 si(in), sf(flow), sg(grad_out);
    at::Tensor m;
    if (metric.has_value()) m = metric->expand({N, 1, H, W});
    int64_t sm[4] = {0, 0, 0, 0};
    if (m.defined()) for (int i = 0; i < 4; ++i) sm[i] = m.stride(i);
    const size_t ws_bytes = fldr_splat_bwd_workspace_bytes((int)mode, N, C, H, W);
    at::Tensor ws = workspace(ws_bytes, in, stream);
    check(fldr_splat_bwd((int)mode, fptr(in), si.s, fptr(flow), sf.s, m.defined() ? fptr(m) : nullptr, m.defined() ? sm : nullptr,
                         out.has_value() ? fptr(*out) : nullptr, norm.has_value() ? fptr(*norm) : nullptr, fptr(grad_out), sg.s,
                         gin.defined() ? fptr_mut(gin) : nullptr, gfl.defined() ? fptr_mut(gfl) : nullptr,
                         gme.defined() ? fptr_mut(gme) : nullptr, N, C, H, W, ws.data_ptr(), (size_t)ws.numel(),
                         reinterpret_cast<fldr_stream_t>(stream)));
    return std::make_tuple(gin, gfl, gme);
}

at::Tensor corr81_fwd(const at::Tensor& first, const at::Tensor& second) {
    const c10::cuda::CUDAGuard guard(first.device());
    const int B = first.size(0), C = first.size(1), H = first.size(2), W = first.size(3);
    at::Tensor out = at::empty({B, 81, H, W}, first.options());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(first.get_device()).stream();
    Strides4 s1(first), s2(second);
    check(fldr_corr81_fwd(fptr(first), s1.s, fptr(second), s2.s, fptr_mut(out), B, C, H, W, nullptr, 0, reinterpret_cast<fldr_stream_t>(stream)));
    return out;
}

std::tuple<at::Tensor, at::Tensor> corr81_bwd(const at::Tensor& first, const at::Tensor& second, const at::Tensor& grad_out, bool need_first,
                                              bool need_second) {
    const c10::cuda::CUDAGuard guard(first.device());
    const int B = first.size(0), C = first.size(1), H = first.size(2), W = first.size(3);
    at::Tensor g1, g2;
    if (need_first) g1 = at::empty_like(first, first.options(), at::MemoryFormat::Contiguous);
    if (need_second) g2 = at::empty_like(first, first.options(), at::MemoryFormat::Contiguous);
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(first.get_device()).stream();
    Strides4 s1(first), s2(second), sg(grad_out);
    check(fldr_corr81_bwd(fptr(first), s1.s, fptr(second), s2.s, fptr(grad_out), sg.s, g1.defined() ? fptr_mut(g1) : nullptr,
                          g2.defined() ? fptr_mut(g2) : nullptr, B, C, H, W, nullptr, 0, reinterpret_cast<fldr_stream_t>(stream)));
    return std::make_tuple(g1, g2);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "thin torch binding of libfldr_b200.so (the C ABI of include/fldr_b200.h)";
    m.def("abi_version", []() { return fldr_abi_version(); });
    m.def("splat_fwd", &splat_fwd);
    m.def("splat_bwd", &splat_bwd);
    m.def("corr81_fwd", &corr81_fwd);
    m.def("corr81_bwd", &corr81_bwd);
}
