// =====================================================================================
// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" surface over the CPU restatement (qmc_oracle.hpp / qmc_oracle_driver.hpp) so that
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can call it
// through ctypes.  Built twice by oracle/Makefile:
//   oracle/liboracle.so          the restatement ("port")
//   oracle/_ref/libqmcref.so     same driver, but spline VGH / delayed update / inversion are the
//                                reference's own headers compiled from /root/reference (-DQMC_ORACLE_USE_REFERENCE)
// Precision codes: 0 = double, 1 = float.
// =====================================================================================
#ifdef QMC_ORACLE_USE_REFERENCE
#include "ref_kernels.hpp"
#endif
#include "qmc_oracle_driver.hpp"
#include <chrono>
#include <cstdio>
#include <string>

using namespace orc;

namespace
{
std::string g_err;
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
}
} // namespace

extern "C"
{
const char* orc_last_error() { return g_err.c_str(); }
int orc_is_reference_build()
{
#ifdef QMC_ORACLE_USE_REFERENCE
  return 1;
#else
  return 0;
#endif
}
int orc_max_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n);
#endif
}

// ---------------------------------------------------------------- small pieces (golden vectors)
double orc_symtrace(double h00, double h01, double h02, double h11, double h12, double h22, const double* gg)
{
  return SymTrace(h00, h01, h02, h11, h12, h22, gg);
}
void orc_prefactors_d(double t, double* a, double* da, double* d2a) { compute_prefactors(a, da, d2a, t); }
void orc_prefactors_f(float t, float* a, float* da, float* d2a) { compute_prefactors(a, da, d2a, t); }
void orc_prefactors_v_d(double t, double* a) { compute_prefactors(a, t); }
void orc_prefactors_v_f(float t, float* a) { compute_prefactors(a, t); }
void orc_spline_bound_d(double x, int nmax, int* ind, double* dx) { getSplineBound(x, nmax, *ind, *dx); }
#ifdef QMC_ORACLE_USE_REFERENCE
extern "C" int orc_ref_einspline_create_3d_d(const int* M, const double* data, double* coefs); // ref_einspline.c
void orc_create_periodic_coefs_3d(const int* M, const double* data, double* coefs)
{
  orc_ref_einspline_create_3d_d(M, data, coefs);
}
#else
void orc_create_periodic_coefs_3d(const int* M, const double* data, double* coefs)
{
  create_periodic_coefs_3d(M, data, coefs);
}
#endif
long orc_aligned_size(int prec, long n) { return prec ? (long)aligned_size<float>(n) : (long)aligned_size<double>(n); }

// ---------------------------------------------------------------- raw table evaluation (unit coordinates)
// which: 0 = V, 1 = VGL, 2 = VGH.  out layouts: V [npad]; VGL v[npad] g[3][npad] l[3][npad]; VGH v[npad] g[3][npad] h[6][npad]
#ifdef QMC_ORACLE_USE_REFERENCE
#define ORC_EVAL_BODY                                                                                                  \
  auto m = make_ref_spline(t);                                                                                         \
  if (which == 0)                                                                                                      \
    spline2::evaluate_v_impl(&m, x, y, z, v, 0, ns);                                                                   \
  else if (which == 1)                                                                                                 \
    spline2::evaluate_vgl_impl(&m, x, y, z, v, g, h, (size_t)npad, 0, ns);                                             \
  else                                                                                                                 \
    spline2::evaluate_vgh_impl(&m, x, y, z, v, g, h, (size_t)npad, 0, ns);
#else
#define ORC_EVAL_BODY                                                                                                  \
  if (which == 0)                                                                                                      \
    evaluate_v(t, x, y, z, v);                                                                                         \
  else if (which == 1)                                                                                                 \
    evaluate_vgl(t, x, y, z, v, g, h, (size_t)npad);                                                                   \
  else                                                                                                                 \
    evaluate_vgh(t, x, y, z, v, g, h, (size_t)npad);
#endif
#define ORC_SPLINE_EVAL(SUF, ST)                                                                                       \
  int orc_spline_eval_##SUF(const ST* coefs, const int* grid, int ns, long npad, int which, ST x, ST y, ST z, ST* v,   \
                            ST* g, ST* h)                                                                              \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      SplineTable<ST> t;                                                                                               \
      t.set(coefs, grid, ns, (size_t)npad);                                                                            \
      ORC_EVAL_BODY                                                                                                    \
    });                                                                                                                \
  }
ORC_SPLINE_EVAL(d, double)
ORC_SPLINE_EVAL(f, float)

// ---------------------------------------------------------------- SPOSet-level batched calls
// kind 0: SplineR2R, positions Cartesian double [nw][3] (cast to RT = ST's precision class inside).
// invrow: [nw][ld_inv] VT or NULL (then ratios/grads are not produced).  psi [nw][n], dpsi [nw][n][3], d2psi [nw][n].
#define ORC_R2R(SUF, ST)                                                                                               \
  int orc_r2r_mw_evaluate_vgl_##SUF(const ST* coefs, const int* grid, int ns, long npad, const double* G,              \
                                    const int* halfG, int norb, int nw, const double* r, ST* psi, ST* dpsi,            \
                                    ST* d2psi)                                                                         \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      SplineTable<ST> t;                                                                                               \
      t.set(coefs, grid, ns, (size_t)npad);                                                                            \
      LatticeG lat;                                                                                                    \
      lat.set(G, halfG);                                                                                               \
      SplineR2R<ST, ST> spo;                                                                                           \
      spo.init(t, lat, norb);                                                                                          \
      for (int iw = 0; iw < nw; ++iw)                                                                                  \
      {                                                                                                                \
        ST pos[3] = {(ST)r[3 * iw], (ST)r[3 * iw + 1], (ST)r[3 * iw + 2]};                                             \
        spo.evaluateVGL(pos, psi + (size_t)iw * norb, dpsi + (size_t)iw * norb * 3, d2psi + (size_t)iw * norb);        \
      }                                                                                                                \
    });                                                                                                                \
  }                                                                                                                    \
  int orc_r2r_mw_evaluate_value_##SUF(const ST* coefs, const int* grid, int ns, long npad, const double* G,            \
                                      const int* halfG, int norb, int nw, const double* r, ST* psi)                    \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      SplineTable<ST> t;                                                                                               \
      t.set(coefs, grid, ns, (size_t)npad);                                                                            \
      LatticeG lat;                                                                                                    \
      lat.set(G, halfG);                                                                                               \
      SplineR2R<ST, ST> spo;                                                                                           \
      spo.init(t, lat, norb);                                                                                          \
      for (int iw = 0; iw < nw; ++iw)                                                                                  \
      {                                                                                                                \
        ST pos[3] = {(ST)r[3 * iw], (ST)r[3 * iw + 1], (ST)r[3 * iw + 2]};                                             \
        spo.evaluateValue(pos, psi + (size_t)iw * norb);                                                               \
      }                                                                                                                \
    });                                                                                                                \
  }                                                                                                                    \
  int orc_r2r_mw_vgl_ratio_grads_##SUF(const ST* coefs, const int* grid, int ns, long npad, const double* G,           \
                                       const int* halfG, int norb, int nw, const double* r, const ST* invrow,          \
                                       long ld_inv, ST* phi_vgl, ST* ratios, ST* grads)                                \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      SplineTable<ST> t;                                                                                               \
      t.set(coefs, grid, ns, (size_t)npad);                                                                            \
      LatticeG lat;                                                                                                    \
      lat.set(G, halfG);                                                                                               \
      SplineR2R<ST, ST> spo;                                                                                           \
      spo.init(t, lat, norb);                                                                                          \
      std::vector<ST> pos(3 * (size_t)nw);                                                                             \
      std::vector<const ST*> rows(nw);                                                                                 \
      for (int iw = 0; iw < nw; ++iw)                                                                                  \
      {                                                                                                                \
        for (int d = 0; d < 3; ++d)                                                                                    \
          pos[3 * iw + d] = (ST)r[3 * iw + d];                                                                         \
        rows[iw] = invrow + (size_t)iw * ld_inv;                                                                       \
      }                                                                                                                \
      mw_evaluateVGLandDetRatioGrads(spo, nw, pos.data(), rows.data(), phi_vgl, ratios, grads);                        \
    });                                                                                                                \
  }
ORC_R2R(d, double)
ORC_R2R(f, float)

// complex orbitals: outputs interleaved (re,im)
#define ORC_C2C(SUF, ST)                                                                                               \
  int orc_c2c_mw_evaluate_vgl_##SUF(const ST* coefs, const int* grid, int ns, long npad, const double* G,              \
                                    const double* kcart, int norb, int nw, const double* r, ST* psi, ST* dpsi,         \
                                    ST* d2psi, int value_only)                                                         \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      using CT = std::complex<ST>;                                                                                     \
      SplineTable<ST> t;                                                                                               \
      t.set(coefs, grid, ns, (size_t)npad);                                                                            \
      LatticeG lat;                                                                                                    \
      lat.set(G, nullptr);                                                                                             \
      SplineC2C<ST, CT> spo;                                                                                           \
      spo.init(t, lat, norb, kcart);                                                                                   \
      for (int iw = 0; iw < nw; ++iw)                                                                                  \
      {                                                                                                                \
        ST pos[3] = {(ST)r[3 * iw], (ST)r[3 * iw + 1], (ST)r[3 * iw + 2]};                                             \
        if (value_only)                                                                                                \
          spo.evaluateValue(pos, reinterpret_cast<CT*>(psi) + (size_t)iw * norb);                                      \
        else                                                                                                           \
          spo.evaluateVGL(pos, reinterpret_cast<CT*>(psi) + (size_t)iw * norb,                                         \
                          reinterpret_cast<CT*>(dpsi) + (size_t)iw * norb * 3,                                         \
                          reinterpret_cast<CT*>(d2psi) + (size_t)iw * norb);                                           \
      }                                                                                                                \
    });                                                                                                                \
  }                                                                                                                    \
  int orc_c2c_mw_vgl_ratio_grads_##SUF(const ST* coefs, const int* grid, int ns, long npad, const double* G,           \
                                       const double* kcart, int norb, int nw, const double* r, const ST* invrow,       \
                                       long ld_inv, ST* phi_vgl, ST* ratios, ST* grads)                                \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      using CT = std::complex<ST>;                                                                                     \
      SplineTable<ST> t;                                                                                               \
      t.set(coefs, grid, ns, (size_t)npad);                                                                            \
      LatticeG lat;                                                                                                    \
      lat.set(G, nullptr);                                                                                             \
      SplineC2C<ST, CT> spo;                                                                                           \
      spo.init(t, lat, norb, kcart);                                                                                   \
      std::vector<ST> pos(3 * (size_t)nw);                                                                             \
      std::vector<const CT*> rows(nw);                                                                                 \
      for (int iw = 0; iw < nw; ++iw)                                                                                  \
      {                                                                                                                \
        for (int d = 0; d < 3; ++d)                                                                                    \
          pos[3 * iw + d] = (ST)r[3 * iw + d];                                                                         \
        rows[iw] = reinterpret_cast<const CT*>(invrow) + (size_t)iw * ld_inv;                                          \
      }                                                                                                                \
      mw_evaluateVGLandDetRatioGrads(spo, nw, pos.data(), rows.data(), reinterpret_cast<CT*>(phi_vgl),                 \
                                     reinterpret_cast<CT*>(ratios), reinterpret_cast<CT*>(grads));                     \
    });                                                                                                                \
  }
ORC_C2C(d, double)
ORC_C2C(f, float)

// ---------------------------------------------------------------- dense inverse
// a: psiM [n][a_cols] row-major; inv: [n][lda] receives (psiM^-1)^T; logdet[2] = (re, im)
int orc_invert_transpose_d(const double* a, int n, int a_cols, double* inv, int lda, double* logdet)
{
  return guarded([&] {
    std::complex<double> ld;
    kernel_invert_transpose(a, n, a_cols, inv, lda, ld);
    logdet[0] = ld.real();
    logdet[1] = ld.imag();
  });
}
int orc_invert_transpose_f(const float* a, int n, int a_cols, float* inv, int lda, double* logdet)
{
  return guarded([&] {
    std::complex<double> ld;
    kernel_invert_transpose(a, n, a_cols, inv, lda, ld);
    logdet[0] = ld.real();
    logdet[1] = ld.imag();
  });
}

// complex variants: buffers are std::complex (interleaved re, im), FullPrecValueType = complex<double>
int orc_invert_transpose_z(const std::complex<double>* a, int n, int a_cols, std::complex<double>* inv, int lda,
                           double* logdet)
{
  return guarded([&] {
    std::complex<double> ld;
    kernel_invert_transpose(a, n, a_cols, inv, lda, ld);
    logdet[0] = ld.real();
    logdet[1] = ld.imag();
  });
}
int orc_invert_transpose_c(const std::complex<float>* a, int n, int a_cols, std::complex<float>* inv, int lda, double* logdet)
{
  return guarded([&] {
    std::complex<double> ld;
    kernel_invert_transpose(a, n, a_cols, inv, lda, ld);
    logdet[0] = ld.real();
    logdet[1] = ld.imag();
  });
}
} // extern "C"
template<typename T>
static inline T make_value(double re, double) { return (T)re; }
template<>
inline std::complex<double> make_value<std::complex<double>>(double re, double im) { return {re, im}; }
template<>
inline std::complex<float> make_value<std::complex<float>>(double re, double im) { return {(float)re, (float)im}; }
extern "C" {

// ---------------------------------------------------------------- delayed-update engine (one walker)
#define ORC_DU(SUF, T)                                                                                                 \
  void* orc_du_create_##SUF(int n, int k)                                                                              \
  {                                                                                                                    \
    auto* e = new Engine<T>();                                                                                  \
    e->resize(n, k);                                                                                                   \
    return e;                                                                                                          \
  }                                                                                                                    \
  void orc_du_destroy_##SUF(void* h) { delete static_cast<Engine<T>*>(h); }                                     \
  void orc_du_get_inv_row_##SUF(void* h, const T* Ainv, int lda, int row, T* invRow)                                   \
  {                                                                                                                    \
    static_cast<Engine<T>*>(h)->getInvRow(Ainv, lda, row, invRow);                                              \
  }                                                                                                                    \
  void orc_du_accept_row_##SUF(void* h, T* Ainv, int lda, int row, const T* psiV, double ratio, double ratio_im)       \
  {                                                                                                                    \
    static_cast<Engine<T>*>(h)->acceptRow(Ainv, lda, row, psiV, make_value<T>(ratio, ratio_im));                \
  }                                                                                                                    \
  void orc_du_pseudo_accept_row_##SUF(void* h, T* Ainv, int lda, int row)                                              \
  {                                                                                                                    \
    static_cast<Engine<T>*>(h)->pseudoAcceptRow(Ainv, lda, row);                                                \
  }                                                                                                                    \
  void orc_du_update_inv_mat_##SUF(void* h, T* Ainv, int lda) { static_cast<Engine<T>*>(h)->updateInvMat(Ainv, lda); } \
  int orc_du_delay_count_##SUF(void* h) { return static_cast<Engine<T>*>(h)->delay_count; }                     \

ORC_DU(d, double)
ORC_DU(f, float)
ORC_DU(z, std::complex<double>)
ORC_DU(c, std::complex<float>)

// ---------------------------------------------------------------- Jastrow functor + distance rows
#define ORC_J(SUF, RT)                                                                                                 \
  int orc_functor_eval_##SUF(const double* params, int np, double rcut, double cusp, int nr, const RT* r, RT* u,       \
                             RT* du, RT* d2u)                                                                          \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      BsplineFunctor<RT> f;                                                                                            \
      f.set(params, np, rcut, cusp);                                                                                   \
      for (int i = 0; i < nr; ++i)                                                                                     \
        u[i] = f.evaluate(r[i], du[i], d2u[i]);                                                                        \
    });                                                                                                                \
  }                                                                                                                    \
  int orc_functor_coefs_##SUF(const double* params, int np, double rcut, double cusp, RT* coefs, RT* delta_r_inv)      \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      BsplineFunctor<RT> f;                                                                                            \
      f.set(params, np, rcut, cusp);                                                                                   \
      std::copy(f.coefs.begin(), f.coefs.end(), coefs);                                                                \
      *delta_r_inv = f.DeltaRInv;                                                                                      \
    });                                                                                                                \
  }                                                                                                                    \
  /* rows: rsoa [3][npad]; out [4][npad] */                                                                            \
  int orc_dist_row_##SUF(const double* lattice, const RT* pos, const RT* rsoa, long npad, int nsrc, int flip_ind,      \
                         RT* out)                                                                                      \
  {                                                                                                                    \
    return guarded([&] {                                                                                               \
      MinImage<RT> mi;                                                                                                 \
      mi.set(lattice);                                                                                                 \
      mi.row(pos, rsoa, (size_t)npad, nsrc, flip_ind, out);                                                            \
    });                                                                                                                \
  }
ORC_J(d, double)
ORC_J(f, float)

// ---------------------------------------------------------------- RNG
void orc_rng_uniform(uint32_t seed, int n, double* out)
{
  StdRandom r(seed);
  for (int i = 0; i < n; ++i)
    out[i] = r();
}
void orc_rng_raw(uint32_t seed, long n, uint32_t* out)
{
  std::mt19937 e(seed);
  for (long i = 0; i < n; ++i)
    out[i] = e();
}
void orc_rng_gauss_d(uint32_t seed, int n, double* out)
{
  StdRandom r(seed);
  assignGaussRand(out, (unsigned)n, r);
}
void orc_rng_gauss_f(uint32_t seed, int n, float* out)
{
  StdRandom r(seed);
  assignGaussRand(out, (unsigned)n, r);
}

// stateful generator (one per crowd, like ContextForSteps::get_random_gen())
void* orc_rng_create(uint32_t seed) { return new StdRandom(seed); }
void orc_rng_destroy(void* h) { delete static_cast<StdRandom*>(h); }
double orc_rng_next(void* h) { return (*static_cast<StdRandom*>(h))(); }
void orc_rng_next_n(void* h, int n, double* out)
{
  auto& r = *static_cast<StdRandom*>(h);
  for (int i = 0; i < n; ++i)
    out[i] = r();
}
void orc_rng_gauss_next_d(void* h, int n, double* out) { assignGaussRand(out, (unsigned)n, *static_cast<StdRandom*>(h)); }
void orc_rng_gauss_next_f(void* h, int n, float* out) { assignGaussRand(out, (unsigned)n, *static_cast<StdRandom*>(h)); }

// ---------------------------------------------------------------- VMC harness
struct VMCHandle
{
  int precision;
  VMC<double, double, double>* d = nullptr;
  VMC<float, float, float>* f   = nullptr;
  VMC<double, double, std::complex<double>>* zd = nullptr; // complex orbitals (SplineC2C), full precision
  VMC<float, float, std::complex<float>>* cf    = nullptr; // complex orbitals, mixed precision
};

void* orc_vmc_create(const VMCParams* p)
{
  VMCHandle* h = nullptr;
  int rc       = guarded([&] {
    h            = new VMCHandle;
    h->precision = p->precision;
    if (p->complex_orbitals)
    {
      if (p->precision == 0)
        h->zd = new VMC<double, double, std::complex<double>>(*p);
      else
        h->cf = new VMC<float, float, std::complex<float>>(*p);
    }
    else if (p->precision == 0)
      h->d = new VMC<double, double, double>(*p);
    else
      h->f = new VMC<float, float, float>(*p);
  });
  return rc ? nullptr : h;
}
void orc_vmc_destroy(void* hv)
{
  auto* h = static_cast<VMCHandle*>(hv);
  if (!h)
    return;
  delete h->d;
  delete h->f;
  delete h->zd;
  delete h->cf;
  delete h;
}
#define VMC_DISPATCH(h, expr)                                                                                          \
  do                                                                                                                   \
  {                                                                                                                    \
    if ((h)->d)                                                                                                        \
    {                                                                                                                  \
      auto& v = *(h)->d;                                                                                               \
      expr;                                                                                                            \
    }                                                                                                                  \
    else if ((h)->f)                                                                                                   \
    {                                                                                                                  \
      auto& v = *(h)->f;                                                                                               \
      expr;                                                                                                            \
    }                                                                                                                  \
    else if ((h)->zd)                                                                                                  \
    {                                                                                                                  \
      auto& v = *(h)->zd;                                                                                              \
      expr;                                                                                                            \
    }                                                                                                                  \
    else                                                                                                               \
    {                                                                                                                  \
      auto& v = *(h)->cf;                                                                                              \
      expr;                                                                                                            \
    }                                                                                                                  \
  } while (0)

int orc_vmc_set_positions(void* hv, const double* R)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] { VMC_DISPATCH(h, v.setPositions(R)); });
}
// TrialWaveFunction evalGrad / makeMove / calcRatioGrad of walker iw for a prescribed displacement (no accept);
// ratio [2], grad_old [3][2], grad_new [3][2] as (re, im) doubles
int orc_vmc_probe_move(void* hv, int iw, int iat, const double* displ, double* ratio, double* grad_old, double* grad_new)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    std::complex<double> r, go[3], gn[3];
    VMC_DISPATCH(h, v.probeMove(iw, iat, displ, r, go, gn));
    ratio[0] = r.real(), ratio[1] = r.imag();
    for (int d = 0; d < 3; ++d)
    {
      grad_old[2 * d] = go[d].real(), grad_old[2 * d + 1] = go[d].imag();
      grad_new[2 * d] = gn[d].real(), grad_new[2 * d + 1] = gn[d].imag();
    }
  });
}
// TrialWaveFunction::mw_evaluateRatios for one walker: nk virtual positions of electron `ref`; out [nk][2] (re, im)
int orc_vmc_evaluate_ratios(void* hv, int iw, int ref, int nk, const double* r_vp, int ct, double* out)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    std::vector<std::complex<double>> r(nk);
    VMC_DISPATCH(h, v.evaluateRatios(iw, ref, nk, r_vp, ct, r.data()));
    for (int i = 0; i < nk; ++i)
      out[2 * i] = r[i].real(), out[2 * i + 1] = r[i].imag();
  });
}
int orc_vmc_recompute(void* hv)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] { VMC_DISPATCH(h, v.recompute()); });
}
// runs nsteps sweeps; returns seconds spent in the sweep loop through *seconds; accept_log (optional) [nsteps][N][nw]
int orc_vmc_sweep(void* hv, int nsteps, uint8_t* accept_log, double* seconds)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    auto t0 = std::chrono::steady_clock::now();
    VMC_DISPATCH(h, v.sweep(nsteps, accept_log != nullptr);
                 if (accept_log) std::copy(v.accept_log.begin(), v.accept_log.end(), accept_log));
    auto t1 = std::chrono::steady_clock::now();
    if (seconds)
      *seconds = std::chrono::duration<double>(t1 - t0).count();
  });
}
// teacher-forced variant: accept flags imposed ([nsteps][N][nw]); ratio_log (optional) receives every move's ratio
int orc_vmc_sweep_forced(void* hv, int nsteps, const uint8_t* forced, double* ratio_log)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] { VMC_DISPATCH(h, v.sweep(nsteps, false, forced, ratio_log)); });
}
int orc_vmc_get_positions(void* hv, double* R)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    VMC_DISPATCH(h, for (int iw = 0; iw < v.nw; ++iw) for (size_t i = 0; i < 3 * (size_t)v.N; ++i)
                        R[(size_t)iw * 3 * v.N + i] = v.walkers[iw].R[i]);
  });
}
// per walker: logpsi, kinetic energy, n_accept, n_reject; optional G [nw][N][3], L [nw][N]
int orc_vmc_evaluate_gl(void* hv, double* logpsi, double* ke, double* G, double* L)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    VMC_DISPATCH(h, const size_t cs = v.is_cplx ? 2 : 1; for (int iw = 0; iw < v.nw; ++iw)
                        v.evaluateGL(iw, G ? G + (size_t)iw * 3 * v.N * cs : nullptr, L ? L + (size_t)iw * v.N * cs : nullptr,
                                     logpsi ? logpsi + iw : nullptr, ke ? ke + iw : nullptr));
  });
}
int orc_vmc_get_counts(void* hv, long* n_accept, long* n_reject)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    VMC_DISPATCH(h, for (int iw = 0; iw < v.nw; ++iw) {
      n_accept[iw] = v.walkers[iw].n_accept;
      n_reject[iw] = v.walkers[iw].n_reject;
    });
  });
}
} // extern "C"
static inline void put_value(double* out, size_t i, double v) { out[i] = v; }
static inline void put_value(double* out, size_t i, float v) { out[i] = v; }
template<typename R>
static inline void put_value(double* out, size_t i, const std::complex<R>& v)
{
  out[2 * i]     = v.real();
  out[2 * i + 1] = v.imag();
}
extern "C" {
// psiMinv of spin s for walker iw, as doubles [n][n] (padding stripped; complex: [n][n][2]); logdet (re)
int orc_vmc_get_psiminv(void* hv, int iw, int s, double* out, double* logdet)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    VMC_DISPATCH(h, auto& d = v.walkers[iw].det[s]; for (int i = 0; i < d.n; ++i) for (int j = 0; j < d.n; ++j)
                                                         put_value(out, (size_t)i * d.n + j, d.psiMinv[(size_t)i * d.lda + j]);
                 if (logdet) *logdet = d.log_value.real());
  });
}
int orc_vmc_set_num_walkers(void* hv, int n)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] { VMC_DISPATCH(h, v.setNumWalkers(n)); });
}
int orc_vmc_copy_walker(void* hv, int src, int dst)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] { VMC_DISPATCH(h, v.copyWalker(src, dst)); });
}
// DMC: per-walker rr_accepted / rr_proposed of the last sweep
int orc_vmc_get_rr(void* hv, double* rr_accepted, double* rr_proposed)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    VMC_DISPATCH(h, for (int iw = 0; iw < v.nw; ++iw) {
      rr_accepted[iw] = v.walkers[iw].rr_accepted;
      rr_proposed[iw] = v.walkers[iw].rr_proposed;
    });
  });
}
int orc_vmc_get_j2(void* hv, int iw, double* Uat, double* dUat /*[3][N]*/, double* d2Uat)
{
  auto* h = static_cast<VMCHandle*>(hv);
  return guarded([&] {
    VMC_DISPATCH(h, auto& s = v.walkers[iw].j2; for (int i = 0; i < v.N; ++i) {
      Uat[i]   = s.Uat[i];
      d2Uat[i] = s.d2Uat[i];
      for (int d = 0; d < 3; ++d)
        dUat[(size_t)d * v.N + i] = s.dUat[d * v.j2.npad + i];
    });
  });
}
} // extern "C"
