// qmcpack_b200/csrc/api.cu -- the C ABI (include/qmcb.h) over the host classes.  No torch types, no CPU fallback:
// every entry point fails with a message when CUDA is unavailable.
#include "internal.h"
#include "spline.cuh"
#include <cstring>
#include <vector>

using namespace qmcb;

namespace
{
thread_local std::string g_err;
template<typename F>
int guarded(F&& f)
{
  try
  {
    f();
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return 1;
  }
  catch (...)
  {
    g_err = "unknown error";
    return 2;
  }
}
void need(const void* p, const char* what)
{
  if (!p)
    throw std::runtime_error(std::string("null argument: ") + what);
}
void require_device()
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    throw std::runtime_error("libqmcb: no CUDA device available (there is no CPU fallback)");
}

// host-side staging for the standalone SPOSet entry points
template<typename T>
struct Stage
{
  DevBuf<T> r, inv, phi, rg;
  DevBuf<int> ref;
};

template<typename T>
void spline_eval_host(qmcb_spline* h, int mode, int nw, const double* r_host, const void* invrow_host, size_t ld_inv,
                      int n_rows, const int* ref_host, std::vector<T>* phi_out, std::vector<T>* rg_out)
{
  SplineSPOBase& S = *h->impl;
  const int vt     = S.vt_per_orb();
  Stage<T> g;
  std::vector<T> rh(3 * (size_t)nw);
  for (size_t i = 0; i < rh.size(); ++i)
    rh[i] = (T)r_host[i];
  g.r.alloc(rh.size(), false);
  QMCB_CUDA(cudaMemcpy(g.r.p, rh.data(), rh.size() * sizeof(T), cudaMemcpyHostToDevice));
  if (invrow_host)
  {
    g.inv.alloc((size_t)n_rows * ld_inv * vt, false);
    QMCB_CUDA(cudaMemcpy(g.inv.p, invrow_host, g.inv.bytes(), cudaMemcpyHostToDevice));
    g.rg.alloc((size_t)nw * S.rg_parts() * 4 * vt);
  }
  if (ref_host)
  {
    g.ref.alloc(nw, false);
    QMCB_CUDA(cudaMemcpy(g.ref.p, ref_host, nw * sizeof(int), cudaMemcpyHostToDevice));
  }
  const size_t nphi = (size_t)(mode == MODE_VGL ? 5 : 1) * nw * S.n_orb * vt;
  if (phi_out)
    g.phi.alloc(nphi);
  S.evaluate_dev(mode, nw, g.r.p, g.inv.p, ld_inv, g.ref.p, g.phi.p, g.rg.p, nullptr);
  QMCB_CUDA(cudaDeviceSynchronize());
  if (phi_out)
  {
    phi_out->resize(nphi);
    QMCB_CUDA(cudaMemcpy(phi_out->data(), g.phi.p, nphi * sizeof(T), cudaMemcpyDeviceToHost));
  }
  if (rg_out && invrow_host)
  {
    // add the per-(tile, warp) partial dots in index order
    const int np = S.rg_parts(), nred = 4 * vt;
    std::vector<T> parts((size_t)nw * np * nred);
    QMCB_CUDA(cudaMemcpy(parts.data(), g.rg.p, parts.size() * sizeof(T), cudaMemcpyDeviceToHost));
    rg_out->assign((size_t)nw * nred, T(0));
    for (int iw = 0; iw < nw; ++iw)
      for (int q = 0; q < np; ++q)
        for (int e = 0; e < nred; ++e)
          (*rg_out)[(size_t)iw * nred + e] += parts[((size_t)iw * np + q) * nred + e];
  }
}

// complex-aware division a/b on interleaved storage
template<typename T>
void cdiv(const T* a, const T* b, T* out, int vt)
{
  if (vt == 1)
    out[0] = a[0] / b[0];
  else
  {
    const T d = b[0] * b[0] + b[1] * b[1];
    out[0]    = (a[0] * b[0] + a[1] * b[1]) / d;
    out[1]    = (a[1] * b[0] - a[0] * b[1]) / d;
  }
}

template<typename T>
void spline_vgl_host(qmcb_spline* h, int nw, const double* r, void* psi, void* dpsi, void* d2psi)
{
  SplineSPOBase& S = *h->impl;
  const int vt = S.vt_per_orb(), n = S.n_orb;
  std::vector<T> phi;
  spline_eval_host<T>(h, MODE_VGL, nw, r, nullptr, 0, 0, nullptr, &phi, nullptr);
  const size_t fs = (size_t)nw * n * vt;
  T* p  = static_cast<T*>(psi);
  T* dp = static_cast<T*>(dpsi);
  T* d2 = static_cast<T*>(d2psi);
  for (int iw = 0; iw < nw; ++iw)
    for (int j = 0; j < n; ++j)
      for (int c = 0; c < vt; ++c)
      {
        const size_t src = ((size_t)iw * n + j) * vt + c;
        if (p)
          p[src] = phi[src];
        if (dp)
          for (int d = 0; d < 3; ++d)
            dp[(((size_t)iw * n + j) * 3 + d) * vt + c] = phi[(1 + d) * fs + src];
        if (d2)
          d2[src] = phi[4 * fs + src];
      }
}

template<typename T>
void spline_ratio_host(qmcb_spline* h, int nw, const double* r, const void* invrow, size_t ld, void* phi_out,
                       void* ratios, void* grads)
{
  SplineSPOBase& S = *h->impl;
  const int vt     = S.vt_per_orb();
  std::vector<T> phi, rg;
  spline_eval_host<T>(h, MODE_VGL, nw, r, invrow, ld, nw, nullptr, phi_out ? &phi : nullptr, &rg);
  if (phi_out)
    std::memcpy(phi_out, phi.data(), phi.size() * sizeof(T));
  T* ra = static_cast<T*>(ratios);
  T* gr = static_cast<T*>(grads);
  for (int iw = 0; iw < nw; ++iw)
  {
    const T* q = &rg[(size_t)iw * 4 * vt];
    for (int c = 0; c < vt; ++c)
      ra[(size_t)iw * vt + c] = q[c];
    if (gr)
      for (int d = 0; d < 3; ++d)
        cdiv<T>(q + (1 + d) * vt, q, gr + ((size_t)iw * 3 + d) * vt, vt);
  }
}

// the reference's contract for an offload SPOSet (SPOSet.h:346-352 with isOMPoffload() == true): positions from the host,
// inverse rows and phi_vgl_v in device memory, ratios and gradients back on the host
template<typename T>
void spline_ratio_offload(qmcb_spline* h, int nw, const double* r, const void* invrow_dev, size_t ld, void* phi_dev,
                          void* ratios, void* grads, cudaStream_t st)
{
  SplineSPOBase& S = *h->impl;
  const int vt = S.vt_per_orb(), np = S.rg_parts(), nred = 4 * vt;
  Stage<T> g;
  std::vector<T> rh(3 * (size_t)nw);
  for (size_t i = 0; i < rh.size(); ++i)
    rh[i] = (T)r[i];
  g.r.alloc(rh.size(), false);
  g.rg.alloc((size_t)nw * np * nred);
  QMCB_CUDA(cudaMemcpyAsync(g.r.p, rh.data(), rh.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  S.evaluate_dev(MODE_VGL, nw, g.r.p, invrow_dev, ld, nullptr, phi_dev, g.rg.p, st);
  std::vector<T> parts((size_t)nw * np * nred);
  QMCB_CUDA(cudaMemcpyAsync(parts.data(), g.rg.p, parts.size() * sizeof(T), cudaMemcpyDeviceToHost, st));
  QMCB_CUDA(cudaStreamSynchronize(st));
  T* ra = static_cast<T*>(ratios);
  T* gr = static_cast<T*>(grads);
  std::vector<T> q(nred);
  for (int iw = 0; iw < nw; ++iw)
  {
    // the per-(tile, warp) partial dots in index order (the reference's host code adds its per-team partials the same
    // way, SplineR2R.cpp:566-581)
    std::fill(q.begin(), q.end(), T(0));
    for (int t = 0; t < np; ++t)
      for (int e = 0; e < nred; ++e)
        q[e] += parts[((size_t)iw * np + t) * nred + e];
    for (int c = 0; c < vt; ++c)
      ra[(size_t)iw * vt + c] = q[c];
    if (gr)
      for (int d = 0; d < 3; ++d)
        cdiv<T>(q.data() + (1 + d) * vt, q.data(), gr + ((size_t)iw * 3 + d) * vt, vt);
  }
}
} // namespace

extern "C"
{
const char* qmcb_last_error(void) { return g_err.c_str(); }

int qmcb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int qmcb_init(int device)
{
  return guarded([&] {
    require_device();
    QMCB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QMCB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw std::runtime_error(std::string("libqmcb is built for sm_100a only; device is ") + prop.name + " (sm_" +
                               std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    QMCB_CUDA(cudaFree(0));
  });
}

size_t qmcb_aligned_size(int precision, size_t n)
{
  return precision == QMCB_MIXED ? aligned_size<float>(n) : aligned_size<double>(n);
}
unsigned long long qmcb_kernel_launch_count(void) { return g_launch_count.load(); }

// ---- SPOSet
int qmcb_spline_create(qmcb_spline** h, int precision, int kind, const int grid[3], int n_orb, int n_spl, size_t npad,
                       const void* coefs_host, const double G[9], const int halfG[3], const double* kcart)
{
  return guarded([&] {
    need(h, "handle");
    need(grid, "grid");
    need(coefs_host, "coefs_host");
    need(G, "G");
    require_device();
    auto* s = new qmcb_spline;
    s->impl.reset(make_spline(precision, kind, grid, n_orb, n_spl, npad, coefs_host, G, halfG, kcart));
    QMCB_CUDA(cudaGetDevice(&s->impl->device));
    *h = s;
  });
}
int qmcb_spline_destroy(qmcb_spline* h)
{
  return guarded([&] { delete h; });
}
size_t qmcb_spline_table_bytes(const qmcb_spline* h) { return h ? h->impl->table_bytes() : 0; }

int qmcb_spline_mw_evaluate_value(qmcb_spline* h, int nw, const double* r_host, void* psi_host)
{
  return guarded([&] {
    need(h, "spline");
    QMCB_CUDA(cudaSetDevice(h->impl->device));
    need(r_host, "r_host");
    need(psi_host, "psi_host");
    if (h->impl->precision == QMCB_MIXED)
    {
      std::vector<float> phi;
      spline_eval_host<float>(h, MODE_V, nw, r_host, nullptr, 0, 0, nullptr, &phi, nullptr);
      std::memcpy(psi_host, phi.data(), phi.size() * sizeof(float));
    }
    else
    {
      std::vector<double> phi;
      spline_eval_host<double>(h, MODE_V, nw, r_host, nullptr, 0, 0, nullptr, &phi, nullptr);
      std::memcpy(psi_host, phi.data(), phi.size() * sizeof(double));
    }
  });
}
int qmcb_spline_mw_evaluate_vgl(qmcb_spline* h, int nw, const double* r_host, void* psi_host, void* dpsi_host,
                                void* d2psi_host)
{
  return guarded([&] {
    need(h, "spline");
    QMCB_CUDA(cudaSetDevice(h->impl->device));
    need(r_host, "r_host");
    if (h->impl->precision == QMCB_MIXED)
      spline_vgl_host<float>(h, nw, r_host, psi_host, dpsi_host, d2psi_host);
    else
      spline_vgl_host<double>(h, nw, r_host, psi_host, dpsi_host, d2psi_host);
  });
}
int qmcb_spline_mw_evaluate_vgl_ratio_grads(qmcb_spline* h, int nw, const double* r_host, const void* invrow_host,
                                            size_t ld_inv, void* phi_vgl_host, void* ratios_host, void* grads_host)
{
  return guarded([&] {
    need(h, "spline");
    QMCB_CUDA(cudaSetDevice(h->impl->device));
    need(r_host, "r_host");
    need(invrow_host, "invrow_host");
    need(ratios_host, "ratios_host");
    if (h->impl->precision == QMCB_MIXED)
      spline_ratio_host<float>(h, nw, r_host, invrow_host, ld_inv, phi_vgl_host, ratios_host, grads_host);
    else
      spline_ratio_host<double>(h, nw, r_host, invrow_host, ld_inv, phi_vgl_host, ratios_host, grads_host);
  });
}
int qmcb_spline_mw_evaluate_det_ratios(qmcb_spline* h, int nvp, const double* r_vp_host, const int* ref_walker_host,
                                       int n_ref, const void* invrow_host, size_t ld_inv, void* ratios_host)
{
  return guarded([&] {
    need(h, "spline");
    QMCB_CUDA(cudaSetDevice(h->impl->device));
    need(r_vp_host, "r_vp_host");
    need(ref_walker_host, "ref_walker_host");
    need(invrow_host, "invrow_host");
    need(ratios_host, "ratios_host");
    for (int i = 0; i < nvp; ++i)
      if (ref_walker_host[i] < 0 || ref_walker_host[i] >= n_ref)
        throw std::runtime_error("ref_walker index out of range");
    const int vt = h->impl->vt_per_orb();
    if (h->impl->precision == QMCB_MIXED)
    {
      std::vector<float> rg;
      spline_eval_host<float>(h, MODE_V, nvp, r_vp_host, invrow_host, ld_inv, n_ref, ref_walker_host, nullptr, &rg);
      float* out = static_cast<float*>(ratios_host);
      for (int i = 0; i < nvp; ++i)
        for (int c = 0; c < vt; ++c)
          out[(size_t)i * vt + c] = rg[(size_t)i * 4 * vt + c];
    }
    else
    {
      std::vector<double> rg;
      spline_eval_host<double>(h, MODE_V, nvp, r_vp_host, invrow_host, ld_inv, n_ref, ref_walker_host, nullptr, &rg);
      double* out = static_cast<double*>(ratios_host);
      for (int i = 0; i < nvp; ++i)
        for (int c = 0; c < vt; ++c)
          out[(size_t)i * vt + c] = rg[(size_t)i * 4 * vt + c];
    }
  });
}
int qmcb_spline_mw_vgl_ratio_grads_dev(qmcb_spline* h, int nw, const void* r_dev, const void* invrow_dev, size_t ld_inv,
                                       void* phi_vgl_dev, void* ratio_grad_dev, void* stream)
{
  return guarded([&] {
    need(h, "spline");
    QMCB_CUDA(cudaSetDevice(h->impl->device));
    need(r_dev, "r_dev");
    h->impl->evaluate_dev(MODE_VGL, nw, r_dev, invrow_dev, ld_inv, nullptr, phi_vgl_dev, ratio_grad_dev,
                          static_cast<cudaStream_t>(stream));
  });
}

int qmcb_spline_rg_parts(const qmcb_spline* h) { return h ? h->impl->rg_parts() : 0; }

int qmcb_spline_mw_evaluate_vgl_ratio_grads_offload(qmcb_spline* h, int nw, const double* r_host, const void* invrow_dev,
                                                    size_t ld_inv, void* phi_vgl_dev, void* ratios_host, void* grads_host,
                                                    void* stream)
{
  return guarded([&] {
    need(h, "spline");
    QMCB_CUDA(cudaSetDevice(h->impl->device));
    need(r_host, "r_host");
    need(invrow_dev, "invrow_dev");
    need(ratios_host, "ratios_host");
    if (nw <= 0)
      return;
    if (h->impl->precision == QMCB_MIXED)
      spline_ratio_offload<float>(h, nw, r_host, invrow_dev, ld_inv, phi_vgl_dev, ratios_host, grads_host,
                                  static_cast<cudaStream_t>(stream));
    else
      spline_ratio_offload<double>(h, nw, r_host, invrow_dev, ld_inv, phi_vgl_dev, ratios_host, grads_host,
                                   static_cast<cudaStream_t>(stream));
  });
}

// ---- crowd
int qmcb_crowd_create(qmcb_crowd** c, const qmcb_system* sys, int nw)
{
  return guarded([&] {
    need(c, "handle");
    need(sys, "system");
    require_device();
    auto* p = new qmcb_crowd;
    p->impl.reset(make_crowd(sys, nw));
    QMCB_CUDA(cudaGetDevice(&p->impl->device));
    *c = p;
  });
}
int qmcb_crowd_destroy(qmcb_crowd* c)
{
  return guarded([&] { delete c; });
}
#define CROWD_CALL(expr)                        \
  return guarded([&] {                          \
    need(c, "crowd");                           \
    QMCB_CUDA(cudaSetDevice(c->impl->device));  \
    c->impl->expr;                              \
  })
int qmcb_crowd_sync(qmcb_crowd* c) { CROWD_CALL(sync()); }
size_t qmcb_crowd_device_bytes(const qmcb_crowd* c) { return c ? c->impl->device_bytes() : 0; }
int qmcb_crowd_is_complex(const qmcb_crowd* c) { return (c && c->impl->is_complex()) ? 1 : 0; }
int qmcb_crowd_set_positions(qmcb_crowd* c, const double* R) { CROWD_CALL(set_positions(R)); }
int qmcb_crowd_get_positions(qmcb_crowd* c, double* R) { CROWD_CALL(get_positions(R)); }
int qmcb_twf_mw_recompute(qmcb_crowd* c) { CROWD_CALL(twf_recompute()); }
int qmcb_twf_mw_eval_grad(qmcb_crowd* c, int iat, double* g) { CROWD_CALL(twf_eval_grad(iat, g)); }
int qmcb_ps_mw_make_move(qmcb_crowd* c, int iat, const double* d) { CROWD_CALL(ps_make_move(iat, d)); }
int qmcb_twf_mw_calc_ratio_grad(qmcb_crowd* c, int iat, double* r, double* g) { CROWD_CALL(twf_calc_ratio_grad(iat, r, g)); }
int qmcb_twf_mw_accept_reject(qmcb_crowd* c, int iat, const uint8_t* a, int safe) { CROWD_CALL(twf_accept_reject(iat, a, safe)); }
int qmcb_twf_mw_complete_updates(qmcb_crowd* c) { CROWD_CALL(twf_complete_updates()); }
int qmcb_twf_mw_calc_ratio(qmcb_crowd* c, int iat, double* ratios) { CROWD_CALL(twf_calc_ratio(iat, ratios)); }
int qmcb_twf_mw_evaluate_ratios(qmcb_crowd* c, int nvp, const int* walker, const int* ref, const double* r_vp, int ct, double* ratios)
{
  CROWD_CALL(twf_evaluate_ratios(nvp, walker, ref, r_vp, ct, ratios));
}
int qmcb_twf_mw_evaluate_gl(qmcb_crowd* c, double* G, double* L, double* lp, double* ke) { CROWD_CALL(twf_evaluate_gl(G, L, lp, ke)); }
int qmcb_det_mw_eval_grad(qmcb_crowd* c, int spin, int row, void* g) { CROWD_CALL(det_eval_grad(spin, row, g)); }
int qmcb_det_mw_get_inv_row(qmcb_crowd* c, int spin, int row, const void** dev, size_t* ld, void* host)
{
  CROWD_CALL(det_get_inv_row(spin, row, dev, ld, host));
}
int qmcb_det_mw_ratio_grad(qmcb_crowd* c, int spin, int row, void* r, void* g) { CROWD_CALL(det_ratio_grad(spin, row, r, g, false)); }
int qmcb_det_mw_ratio_grad_from_phi(qmcb_crowd* c, int spin, int row, void* r, void* g) { CROWD_CALL(det_ratio_grad(spin, row, r, g, true)); }
int qmcb_det_mw_accept_reject(qmcb_crowd* c, int spin, int row, const uint8_t* a) { CROWD_CALL(det_accept_reject(spin, row, a)); }
int qmcb_det_mw_complete_updates(qmcb_crowd* c, int spin, void* inv, double* ld) { CROWD_CALL(det_complete_updates(spin, inv, ld)); }
int qmcb_det_mw_recompute_from_matrices(qmcb_crowd* c, int spin, const void* m, const void* dm, const void* d2m)
{
  CROWD_CALL(det_recompute_from_matrices(spin, m, dm, d2m));
}
int qmcb_det_set_phi_vgl(qmcb_crowd* c, int spin, const void* phi) { CROWD_CALL(det_set_phi_vgl(spin, phi)); }
int qmcb_det_delay_count(qmcb_crowd* c, int spin) { return c ? c->impl->det_delay_count(spin) : -1; }
int qmcb_det_time_update_inv_mat(qmcb_crowd* c, int spin, int dc, int reps, double* us) { CROWD_CALL(det_time_update_inv_mat(spin, dc, reps, us)); }
int qmcb_det_time_inverse(qmcb_crowd* c, int spin, int method, int reps, double* us) { CROWD_CALL(det_time_inverse(spin, method, reps, us)); }
int qmcb_dtaa_get_temp_rows(qmcb_crowd* c, void* rows) { CROWD_CALL(dtaa_get_temp_rows(rows)); }
int qmcb_j2_mw_ratio_grad(qmcb_crowd* c, int iat, double* r, void* g) { CROWD_CALL(j2_ratio_grad(iat, r, g)); }
int qmcb_j2_mw_accept_reject(qmcb_crowd* c, int iat, const uint8_t* a) { CROWD_CALL(j2_accept_reject(iat, a)); }
int qmcb_j2_get_state(qmcb_crowd* c, int iw, double* U, double* dU, double* d2U) { CROWD_CALL(j2_get_state(iw, U, dU, d2U)); }
int qmcb_vmc_init(qmcb_crowd* c, const qmcb_vmc_params* p) { CROWD_CALL(vmc_init(p)); }
int qmcb_vmc_sweep(qmcb_crowd* c, int nsteps, uint8_t* log) { CROWD_CALL(vmc_sweep(nsteps, log)); }
int qmcb_vmc_sweep_async(qmcb_crowd* c) { CROWD_CALL(vmc_sweep_async()); }
int qmcb_vmc_counts(qmcb_crowd* c, long long* a, long long* r) { CROWD_CALL(vmc_counts(a, r)); }
int qmcb_vmc_sweep_kernel(qmcb_crowd* c) { return (c && c->impl) ? c->impl->vmc_sweep_kernel() : 0; }
int qmcb_crowd_host_kernel(qmcb_crowd* c) { return (c && c->impl) ? c->impl->host_kernel() : 0; }
int qmcb_vmc_profile_sweep(qmcb_crowd* c, double* out9) { CROWD_CALL(vmc_profile_sweep(out9)); }
int qmcb_dmc_get_rr(qmcb_crowd* c, double* a, double* p) { CROWD_CALL(dmc_get_rr(a, p)); }
size_t qmcb_crowd_walker_bytes(const qmcb_crowd* c) { return c ? c->impl->walker_bytes() : 0; }
int qmcb_crowd_pack_walker(qmcb_crowd* c, int iw, void* buf) { CROWD_CALL(pack_walker(iw, buf)); }
int qmcb_crowd_unpack_walker(qmcb_crowd* c, int iw, const void* buf) { CROWD_CALL(unpack_walker(iw, buf)); }
int qmcb_crowd_copy_walker(qmcb_crowd* c, int src, int dst) { CROWD_CALL(copy_walker(src, dst)); }
int qmcb_crowd_set_num_walkers(qmcb_crowd* c, int n) { CROWD_CALL(set_num_walkers(n)); }
int qmcb_crowd_num_walkers(const qmcb_crowd* c) { return c ? c->impl->num_walkers() : 0; }
int qmcb_crowd_capacity(const qmcb_crowd* c) { return c ? c->impl->capacity() : 0; }
void* qmcb_crowd_stream(qmcb_crowd* c) { return c ? (void*)c->impl->stream() : nullptr; }
} // extern "C"
