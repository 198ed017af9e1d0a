/*
 * gf_torch_ops.cpp -- the C ABI of include/golemflavor_b200.h registered as `torch.ops.golemflavor.*`
 * (BASELINE north star: the fr.py / llh.py callables "dispatch through a thin torch C++/CUDA extension (C-ABI)").
 *
 * Nothing is computed here: every operator checks its tensors (CUDA, float64, contiguous), allocates the outputs
 * with torch, takes torch's current CUDA stream and calls ONE entry point of libgolemflavor_b200.so.  The flattened
 * model travels as a CPU uint8 tensor holding the bytes of `gf_model` (built once per closure by model.flatten).
 * Compiled by golemflavor_b200/build.py into golemflavor_b200/lib/libgolemflavor_b200_torch.so and loaded with
 * torch.ops.load_library; the ctypes binding (_lib.py) remains for the host-buffer calls and for non-torch callers.
 */
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/torch.h>

#include <tuple>

#include "../../include/golemflavor_b200.h"

namespace {

using at::Tensor;

const gf_model* model_of(const Tensor& blob) {
    TORCH_CHECK(blob.device().is_cpu() && blob.scalar_type() == at::kByte && blob.is_contiguous() &&
                    blob.numel() == (int64_t)sizeof(gf_model),
                "golemflavor: model must be a contiguous CPU uint8 tensor of ", sizeof(gf_model), " bytes (gf_model)");
    return reinterpret_cast<const gf_model*>(blob.data_ptr<uint8_t>());
}

void check_theta(const Tensor& theta, const gf_model* m) {
    TORCH_CHECK(theta.is_cuda() && theta.scalar_type() == at::kDouble && theta.dim() == 2 && theta.is_contiguous(),
                "golemflavor: theta must be a contiguous CUDA float64 tensor [N, ndim]");
    TORCH_CHECK(theta.size(1) == m->ndim, "golemflavor: theta has ", theta.size(1), " columns, the model has ndim = ", m->ndim);
}

void check_rc(int rc) {
    if (rc == GF_OK) return;
    /* same mapping as _lib.check: bad arguments -> ValueError (fr.py:198-202), runtime -> RuntimeError */
    if (rc == GF_ERR_ARG) TORCH_CHECK_VALUE(false, gf_last_error());
    TORCH_CHECK(false, gf_last_error());
}

void* stream_of(const Tensor& t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

/* llh.ln_prob (llh.py:121-130): lnprob [N], optional fr [N, 3] and status [N] (empty tensors when not requested) */
std::tuple<Tensor, Tensor, Tensor> lnprob(const Tensor& theta, const Tensor& model, bool want_fr, bool want_status) {
    const gf_model* m = model_of(model);
    check_theta(theta, m);
    const c10::cuda::CUDAGuard guard(theta.device());
    const int64_t n = theta.size(0);
    Tensor out = at::empty({n}, theta.options());
    Tensor fr = want_fr ? at::empty({n, 3}, theta.options()) : at::empty({0}, theta.options());
    Tensor st = want_status ? at::empty({n}, theta.options().dtype(at::kByte)) : at::empty({0}, theta.options().dtype(at::kByte));
    check_rc(gf_lnprob(m, theta.data_ptr<double>(), n, m->ndim, 1, out.data_ptr<double>(), want_fr ? fr.data_ptr<double>() : nullptr,
                       want_status ? st.data_ptr<uint8_t>() : nullptr, stream_of(theta)));
    return {out, fr, st};
}

/* llh.lnprior (llh.py:65-91) */
Tensor lnprior(const Tensor& theta, const Tensor& model) {
    const gf_model* m = model_of(model);
    check_theta(theta, m);
    const c10::cuda::CUDAGuard guard(theta.device());
    Tensor out = at::empty({theta.size(0)}, theta.options());
    check_rc(gf_lnprior(m, theta.data_ptr<double>(), theta.size(0), m->ndim, 1, out.data_ptr<double>(), stream_of(theta)));
    return out;
}

/* fr.flux_averaged_BSMu (fr.py:403-458): fr [N, 3], status [N] */
std::tuple<Tensor, Tensor> flux_averaged_fr(const Tensor& theta, const Tensor& model) {
    const gf_model* m = model_of(model);
    check_theta(theta, m);
    const c10::cuda::CUDAGuard guard(theta.device());
    const int64_t n = theta.size(0);
    Tensor fr = at::empty({n, 3}, theta.options());
    Tensor st = at::empty({n}, theta.options().dtype(at::kByte));
    check_rc(gf_flux_averaged_fr(m, theta.data_ptr<double>(), n, m->ndim, 1, fr.data_ptr<double>(), st.data_ptr<uint8_t>(), stream_of(theta)));
    return {fr, st};
}

Tensor check_f64(const Tensor& t, int64_t last, const char* what) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kDouble && t.is_contiguous() && t.dim() >= 1 && t.size(-1) == last,
                "golemflavor: ", what, " must be a contiguous CUDA float64 tensor [..., ", last, "]");
    return t;
}

/* fr.angles_to_u (fr.py:116-162): angles [N, 4] -> U [N, 3, 3] complex128 */
Tensor angles_to_u(const Tensor& angles) {
    check_f64(angles, 4, "angles");
    const c10::cuda::CUDAGuard guard(angles.device());
    const int64_t n = angles.numel() / 4;
    Tensor u = at::empty({n, 3, 3, 2}, angles.options());
    check_rc(gf_angles_to_u(angles.data_ptr<double>(), n, u.data_ptr<double>(), stream_of(angles)));
    return at::view_as_complex(u);
}

/* fr.angles_to_fr (fr.py:82-113): [N, 2] -> [N, 3] */
Tensor angles_to_fr(const Tensor& src_angles) {
    check_f64(src_angles, 2, "src_angles");
    const c10::cuda::CUDAGuard guard(src_angles.device());
    const int64_t n = src_angles.numel() / 2;
    Tensor fr = at::empty({n, 3}, src_angles.options());
    check_rc(gf_angles_to_fr(src_angles.data_ptr<double>(), n, fr.data_ptr<double>(), stream_of(src_angles)));
    return fr;
}

/* fr.u_to_fr (fr.py:502-536): source [3] or [N, 3], U [N, 3, 3] complex128 -> [N, 3] */
Tensor u_to_fr(const Tensor& source, const Tensor& u) {
    TORCH_CHECK(u.is_cuda() && u.scalar_type() == at::kComplexDouble && u.is_contiguous() && u.dim() == 3 && u.size(1) == 3 && u.size(2) == 3,
                "golemflavor: matrix must be a contiguous CUDA complex128 tensor [N, 3, 3]");
    check_f64(source, 3, "source_fr");
    const int64_t n = u.size(0), ns = source.numel() / 3;
    TORCH_CHECK(ns == 1 || ns == n, "golemflavor: source_fr has ", ns, " rows for ", n, " matrices");
    const c10::cuda::CUDAGuard guard(u.device());
    Tensor ur = at::view_as_real(u);
    Tensor fr = at::empty({n, 3}, source.options());
    check_rc(gf_u_to_fr(source.data_ptr<double>(), ns == 1 ? 0 : 3, ur.data_ptr<double>(), n, fr.data_ptr<double>(), stream_of(u)));
    return fr;
}

/* llh.multi_gaussian (llh.py:32-54): fr [N, 3] -> [N] */
Tensor multi_gaussian(const Tensor& fr, double bf0, double bf1, double bf2, double smearing, double offset, bool emulate_underflow) {
    check_f64(fr, 3, "fr");
    const c10::cuda::CUDAGuard guard(fr.device());
    const int64_t n = fr.numel() / 3;
    Tensor out = at::empty({n}, fr.options());
    const double bf[3] = {bf0, bf1, bf2};
    check_rc(gf_multi_gaussian(fr.data_ptr<double>(), n, bf, smearing, offset, emulate_underflow ? 1 : 0, out.data_ptr<double>(), stream_of(fr)));
    return out;
}

/* Monte-Carlo scan (mc_unitary.py / mc_x.py / mc_texture.py + plot.py:364-370): ADDS into hist [(nb+1)^3] and kept [1] (int64) */
void scan_hist(const Tensor& model, int64_t seed, int64_t first_index, int64_t count, int64_t nb, Tensor hist, Tensor kept) {
    const gf_model* m = model_of(model);
    TORCH_CHECK(hist.is_cuda() && hist.scalar_type() == at::kLong && hist.is_contiguous() && hist.numel() == (nb + 1) * (nb + 1) * (nb + 1),
                "golemflavor: hist must be a contiguous CUDA int64 tensor of (nb+1)^3 cells");
    TORCH_CHECK(kept.is_cuda() && kept.scalar_type() == at::kLong && kept.numel() == 1, "golemflavor: kept must be a CUDA int64 tensor of one element");
    const c10::cuda::CUDAGuard guard(hist.device());
    gf_scan_config cfg = {};
    cfg.seed = (uint64_t)seed;
    cfg.first_index = (uint64_t)first_index;
    cfg.count = (uint64_t)count;
    cfg.nb = (int32_t)nb;
    check_rc(gf_scan_hist(m, &cfg, reinterpret_cast<unsigned long long*>(hist.data_ptr<int64_t>()),
                          reinterpret_cast<unsigned long long*>(kept.data_ptr<int64_t>()), stream_of(hist)));
}

int64_t abi_version() { return gf_abi_version(); }

}  // namespace

TORCH_LIBRARY(golemflavor, lib) {
    lib.def("abi_version() -> int", &abi_version);
    lib.def("lnprob(Tensor theta, Tensor model, bool want_fr=False, bool want_status=False) -> (Tensor, Tensor, Tensor)");
    lib.def("lnprior(Tensor theta, Tensor model) -> Tensor");
    lib.def("flux_averaged_fr(Tensor theta, Tensor model) -> (Tensor, Tensor)");
    lib.def("angles_to_u(Tensor angles) -> Tensor");
    lib.def("angles_to_fr(Tensor src_angles) -> Tensor");
    lib.def("u_to_fr(Tensor source_fr, Tensor matrix) -> Tensor");
    lib.def("multi_gaussian(Tensor fr, float bf0, float bf1, float bf2, float smearing, float offset=-320., bool emulate_underflow=True) -> Tensor");
    lib.def("scan_hist(Tensor model, int seed, int first_index, int count, int nb, Tensor(a!) hist, Tensor(b!) kept) -> ()");
}

/* The model blob is a CPU tensor while the data are CUDA tensors: the operators are registered for every backend
 * (CompositeExplicitAutograd) and validate devices themselves. */
TORCH_LIBRARY_IMPL(golemflavor, CompositeExplicitAutograd, lib) {
    lib.impl("lnprob", &lnprob);
    lib.impl("lnprior", &lnprior);
    lib.impl("flux_averaged_fr", &flux_averaged_fr);
    lib.impl("angles_to_u", &angles_to_u);
    lib.impl("angles_to_fr", &angles_to_fr);
    lib.impl("u_to_fr", &u_to_fr);
    lib.impl("multi_gaussian", &multi_gaussian);
    lib.impl("scan_hist", &scan_hist);
}
