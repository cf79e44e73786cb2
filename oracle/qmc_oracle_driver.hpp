// =====================================================================================
// oracle/qmc_oracle_driver.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see qmc_oracle.hpp).
//
// Restatement of the batched VMC particle-by-particle sweep and of the per-walker wavefunction
// state it drives:
//   ref: QMCDrivers/VMC/VMCBatched.cpp:58-226 (advanceWalkers; hot loop :107-176)
//   ref: QMCWaveFunctions/TrialWaveFunction.cpp:568-592 (mw_evalGrad), :685-741 (mw_calcRatioGrad),
//        :790-833 (mw_accept_rejectMove), :833-869 (mw_completeUpdates), :869-930 (mw_evaluateGL)
//   ref: Fermion/DiracDeterminantBatched.cpp:184-208 (mw_evalGrad), :308-354 (mw_ratioGrad), :464-521
//        (mw_accept_rejectMove), :545-591 (mw_completeUpdates), :594-604 (computeGL), :1122-1198 (mw_recompute)
//   ref: Particle/ParticleSet.cpp:397-428 (mw_makeMove), :717-758 (mw_accept_rejectMove)
// Organisation = the reference's batched CPU path: one crowd per OpenMP thread, walkers looped inside
// every mw_* call (QMCWaveFunctions/SPOSet.cpp:115-126, WaveFunctionComponent.cpp:145-190).
//
// When QMC_ORACLE_USE_REFERENCE is defined (oracle/_ref build), the spline VGH evaluation, the delayed
// update engine and the matrix inversion are the reference's own compiled code
// (spline2::evaluate_vgh_impl, DelayedUpdate<T>, DiracMatrix<T>); see ref_kernels.hpp.
// =====================================================================================
#pragma once
#include "qmc_oracle.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc
{
struct VMCParams
{
  int precision;   // 0 = full (double), 1 = mixed (float)
  int n_up, n_dn;  // electrons per spin group; N = n_up + n_dn
  double lattice[9];
  // spline tables (real R2R), one per spin (may alias); ST-typed [M0+3][M1+3][M2+3][npad]
  const void* coefs[2];
  int grid[3];
  int npad;
  // J2: functor parameter tables uu and ud (10 params each for NiO); n_j2 = 0 disables
  int n_j2;
  const double* j2_uu;
  const double* j2_ud;
  double j2_rcut;
  // J1: nions = 0 disables
  int nions;
  const double* ion_pos; // [nions][3]
  const int* ion_grp;    // [nions]
  int n_ion_groups;
  int n_j1;              // params per group
  const double* j1_params; // [n_ion_groups][n_j1]
  const double* j1_rcut;   // [n_ion_groups]
  // driver
  int nw, ncrowds;
  const uint32_t* seeds; // [ncrowds]
  double tau;
  int use_drift;
  int delay_rank;
  int batched_engine; // 1: DelayedUpdateBatched semantics (crowd-wide delay_count + pseudo-accept); 0: per-walker DelayedUpdate
  // complex orbitals (SplineC2C): tables hold 2*n real components per spin, kpts[s] = [n][3] Cartesian twist vectors
  int complex_orbitals;
  const double* kpts[2];
  // 1: the move loop of DMCBatched::advanceWalkers (DMC/DMCBatched.cpp:142-262) instead of VMCBatched's
  int dmc;
  // optional (0 = the spline lives on `lattice`): the PRIMITIVE cell of the spline table when the simulation cell is a
  // tiling of it (BsplineSet::PrimLattice, test_TrialWaveFunction_diamondC_2x1x1.cpp)
  int has_spline_lattice;
  double spline_lattice[9];
};

#ifndef QMC_ORACLE_USE_REFERENCE
template<typename ST>
inline void kernel_vgh(const SplineTable<ST>& s, ST x, ST y, ST z, ST* v, ST* g, ST* h, size_t stride)
{
  evaluate_vgh(s, x, y, z, v, g, h, stride);
}
template<typename T>
using Engine = DelayedUpdate<T>;
template<typename VT>
inline void kernel_invert_transpose(const VT* amat, int n, int a_cols, VT* inv, int lda, std::complex<double>& logdet)
{
  invert_transpose(amat, n, a_cols, inv, lda, logdet);
}
#endif

template<typename RT, typename ST, typename VT>
struct VMC
{
  using J2 = TwoBodyJastrow<RT>;
  using J1 = OneBodyJastrow<RT>;
  // complex build of the reference: ValueType = complex<RT>, PsiValue = complex<double> (Configuration.h:QMCTraits)
  static constexpr bool is_cplx = !std::is_same<VT, RT>::value;
  using PsiV                    = typename FullPrec<VT>::type;
  static RT real_of(const VT& v)
  {
    // convertToReal (type_traits/ConvertToReal.h:37-44): the real part
    if constexpr (is_cplx)
      return v.real();
    else
      return v;
  }

  struct Det
  {
    int n = 0, lda = 0, first = 0;
    std::vector<VT> psiMinv, dpsiM, d2psiM; // [n][lda], [n][n][3], [n][n]
    std::vector<VT> invRow;
    int invRow_id = -1;
    Engine<VT> eng;
    std::complex<double> log_value;
    PsiV curRatio = PsiV(1);
  };

  struct Walker
  {
    std::vector<RT> R;    // [N][3]
    std::vector<RT> rsoa; // [3][npad_pos]
    Det det[2];
    typename J2::State j2;
    typename J1::State j1;
    RT newpos[3];
    std::vector<RT> dist_new, dist_old; // [4][npad] AA temp rows
    avec<ST> myV, myG, myH;
    std::vector<VT> phi_vgl;            // [5][n] of the proposed move
    double weight = 1.0;
    long n_accept = 0, n_reject = 0;
    RT rr_accepted = 0, rr_proposed = 0; // DMC: sums of tau |delta|^2 over the last sweep (DMCBatched.cpp:139-140)
  };

  struct Crowd
  {
    int w0 = 0, w1 = 0;
    StdRandom rng;
    int delay_count[2] = {0, 0}; // leader's delay_count in batched_engine mode
    std::vector<RT> walker_deltas;
  };

  VMCParams prm;
  int N = 0, nw = 0;
  size_t npad_pos = 0;
  LatticeG latG;
  MinImage<RT> mi;
  SplineTable<ST> tab[2];
  std::vector<ST> kx[2], ky[2], kz[2], mKK[2]; // complex orbitals only (BsplineSet.h:48-52 kPoints / mKK)
  J2 j2;
  J1 j1;
  bool has_j2 = false, has_j1 = false;
  std::vector<Walker> walkers;
  std::vector<Crowd> crowds;
  std::vector<uint8_t> accept_log; // filled by sweep when requested: [step][iat][nw]

  int group_of(int iat) const { return iat < prm.n_up ? 0 : 1; }
  int first_of(int ig) const { return ig == 0 ? 0 : prm.n_up; }
  int last_of(int ig) const { return ig == 0 ? prm.n_up : N; }

  // ref: OhmmsPETE/TensorOps.h:906-923 inverse(Tensor<T,3>), row-major
  static void inverse3(const double* a, double* g)
  {
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    const double vinv = 1.0 / det;
    g[0] = vinv * (a[4] * a[8] - a[5] * a[7]);
    g[1] = vinv * (a[7] * a[2] - a[8] * a[1]);
    g[2] = vinv * (a[1] * a[5] - a[2] * a[4]);
    g[3] = vinv * (a[5] * a[6] - a[3] * a[8]);
    g[4] = vinv * (a[8] * a[0] - a[6] * a[2]);
    g[5] = vinv * (a[2] * a[3] - a[0] * a[5]);
    g[6] = vinv * (a[3] * a[7] - a[4] * a[6]);
    g[7] = vinv * (a[6] * a[1] - a[7] * a[0]);
    g[8] = vinv * (a[0] * a[4] - a[1] * a[3]);
  }

  explicit VMC(const VMCParams& p) : prm(p)
  {
    N        = p.n_up + p.n_dn;
    nw       = p.nw;
    npad_pos = aligned_size<RT>(N);
    // G = inverse(R) (CrystalLattice.cpp:63)
    const double* R = p.lattice;
    // (the inverse of the cell AS GIVEN: the minimum-image object keeps the inverse of the REDUCED basis,
    //  ParticleBConds3DSoa.h:339-386, which is a different matrix for a non-reduced cell)
    double G[9];
    inverse3(p.has_spline_lattice ? p.spline_lattice : R, G);
    latG.set(G, nullptr);
    mi.set(R);
    if (is_cplx != (p.complex_orbitals != 0))
      throw std::runtime_error("oracle VMC: complex_orbitals does not match the instantiated value type");
    for (int s = 0; s < 2; ++s)
    {
      const int n = s == 0 ? p.n_up : p.n_dn;
      tab[s].set(static_cast<const ST*>(p.coefs[s]), p.grid, is_cplx ? 2 * n : n, p.npad);
      if (is_cplx)
      {
        kx[s].resize(n), ky[s].resize(n), kz[s].resize(n), mKK[s].resize(n);
        for (int j = 0; j < n; ++j)
        {
          const double* kk = p.kpts[s] + 3 * (size_t)j;
          kx[s][j] = kk[0], ky[s][j] = kk[1], kz[s][j] = kk[2];
          mKK[s][j] = -(kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2]);
        }
      }
    }
    std::vector<int> gids(N);
    for (int i = 0; i < N; ++i)
      gids[i] = group_of(i);
    has_j2 = p.n_j2 > 0;
    j2.init(N, 2, gids.data());
    if (has_j2)
    {
      // cusp: -1/4 like spin, -1/2 unlike spin (Jastrow/RadialJastrowBuilder.cpp:200-208)
      j2.F[0].set(p.j2_uu, p.n_j2, p.j2_rcut, -0.25);
      j2.F[3].set(p.j2_uu, p.n_j2, p.j2_rcut, -0.25);
      // only the like-spin correlation given: TwoBodyJastrow::addFunc fills every pair slot with it (cusp included)
      const double* ud     = p.j2_ud ? p.j2_ud : p.j2_uu;
      const double cusp_ud = p.j2_ud ? -0.5 : -0.25;
      j2.F[1].set(ud, p.n_j2, p.j2_rcut, cusp_ud);
      j2.F[2].set(ud, p.n_j2, p.j2_rcut, cusp_ud);
    }
    has_j1 = p.nions > 0;
    if (has_j1)
    {
      j1.init(p.nions, p.ion_pos, p.ion_grp, p.n_ion_groups, N);
      for (int g = 0; g < p.n_ion_groups; ++g)
        j1.F[g].set(p.j1_params + (size_t)g * p.n_j1, p.n_j1, p.j1_rcut[g], 0.0);
    }
    walkers.resize(nw);
    for (auto& w : walkers)
    {
      w.R.assign(3 * (size_t)N, RT(0));
      w.rsoa.assign(3 * npad_pos, RT(0));
      for (int s = 0; s < 2; ++s)
      {
        Det& d  = w.det[s];
        d.n     = s == 0 ? p.n_up : p.n_dn;
        d.first = first_of(s);
#ifdef QMC_ORACLE_USE_REFERENCE
        d.lda = d.n; // the reference CPU engine addresses Ainv with lda == norb (DelayedUpdate.h:93)
#else
        d.lda = (int)aligned_size<VT>(d.n);
#endif
        d.psiMinv.assign((size_t)d.n * d.lda, VT(0));
        d.dpsiM.assign((size_t)d.n * d.n * 3, VT(0));
        d.d2psiM.assign((size_t)d.n * d.n, VT(0));
        d.invRow.assign(d.n, VT(0));
        d.eng.resize(d.n, p.delay_rank);
      }
      j2.initState(w.j2);
      if (has_j1)
        j1.initState(w.j1);
      w.dist_new.assign(4 * j2.npad, RT(0));
      w.dist_old.assign(4 * j2.npad, RT(0));
      w.myV.assign(p.npad, ST(0));
      w.myG.assign(3 * (size_t)p.npad, ST(0));
      w.myH.assign(6 * (size_t)p.npad, ST(0));
      w.phi_vgl.assign(5 * (size_t)std::max(p.n_up, p.n_dn), VT(0));
    }
    crowds.resize(p.ncrowds);
    for (int c = 0; c < p.ncrowds; ++c)
      crowds[c].rng = StdRandom(p.seeds[c]);
    dealWalkers();
  }

  // walkers are dealt to crowds in contiguous blocks (MCPopulation::redistributeWalkers / fairDivide:
  // the first (nw % ncrowds) crowds get one extra walker; QMCDrivers/MCPopulation.h + QMCDriverNew.cpp)
  void dealWalkers()
  {
    const int nc = (int)crowds.size();
    int base = nw / nc, extra = nw % nc, w0 = 0;
    for (int c = 0; c < nc; ++c)
    {
      const int cnt = base + (c < extra ? 1 : 0);
      crowds[c].w0  = w0;
      crowds[c].w1  = w0 + cnt;
      w0 += cnt;
    }
  }
  // DMC branching (MCPopulation::spawnWalker / killWalker + Walker copy): the walker array keeps its capacity, the first
  // nw walkers are alive
  void setNumWalkers(int n_active)
  {
    if (n_active < 1 || n_active > (int)walkers.size())
      throw std::runtime_error("oracle VMC: live walker count out of range");
    nw = n_active;
    dealWalkers();
  }
  void copyWalker(int src, int dst)
  {
    if (src != dst)
      walkers[dst] = walkers[src];
  }

  void setPositions(const double* R /*[nw][N][3]*/)
  {
    for (int iw = 0; iw < nw; ++iw)
    {
      Walker& w = walkers[iw];
      for (int i = 0; i < N; ++i)
        for (int d = 0; d < 3; ++d)
        {
          w.R[3 * i + d]             = (RT)R[((size_t)iw * N + i) * 3 + d];
          w.rsoa[d * npad_pos + i] = w.R[3 * i + d];
        }
    }
  }

  // ---- SPO evaluation for one walker at `pos` for spin s: fills phi_vgl[5][n] (SplineR2R.cpp:338-412)
  void spoVGL(Walker& w, int s, const RT pos[3], VT* psi, VT* dpsi /*AoS*/, VT* d2psi)
  {
    if constexpr (is_cplx)
    {
      // SplineC2C::evaluateVGL (SplineC2C.cpp:280-300): toUnit_floor, VGH of the 2n real components, assign_vgl
      ST ru[3];
      for (int j = 0; j < 3; ++j)
      {
        ST v = ST(0);
        for (int i = 0; i < 3; ++i)
          v += ST(pos[i]) * ST(latG.G[i * 3 + j]);
        ru[j] = v;
      }
      for (int i = 0; i < 3; i++) // CrystalLattice.h:187-198
        if (-std::numeric_limits<ST>::epsilon() < ru[i] && ru[i] < 0)
          ru[i] = ST(0.0);
        else
          ru[i] -= std::floor(ru[i]);
      const SplineTable<ST>& t = tab[s];
      kernel_vgh(t, ru[0], ru[1], ru[2], w.myV.data(), w.myG.data(), w.myH.data(), t.npad);
      const ST rc[3] = {ST(pos[0]), ST(pos[1]), ST(pos[2])};
      c2c_assign_vgl<ST, VT>(latG, t.npad, t.ns / 2, w.myV.data(), w.myG.data(), w.myH.data(), kx[s].data(), ky[s].data(),
                             kz[s].data(), mKK[s].data(), rc, psi, dpsi, d2psi);
      return;
    }
    else
      spoVGL_real(w, s, pos, psi, dpsi, d2psi);
  }
  void spoVGL_real(Walker& w, int s, const RT pos[3], VT* psi, VT* dpsi /*AoS*/, VT* d2psi)
  {
    ST ru[3];
    const int bc_sign       = convertPos<ST, RT>(latG, pos, ru);
    const SplineTable<ST>& t = tab[s];
    const size_t np         = t.npad;
    kernel_vgh(t, ru[0], ru[1], ru[2], w.myV.data(), w.myG.data(), w.myH.data(), np);
    const ST signed_one = (bc_sign & 1) ? -1 : 1;
    const ST g00 = latG.G[0], g01 = latG.G[1], g02 = latG.G[2], g10 = latG.G[3], g11 = latG.G[4], g12 = latG.G[5],
             g20 = latG.G[6], g21 = latG.G[7], g22 = latG.G[8];
    const ST symGG[6] = {ST(latG.GGt[0]), ST(latG.GGt[1]) + ST(latG.GGt[3]), ST(latG.GGt[2]) + ST(latG.GGt[6]),
                         ST(latG.GGt[4]), ST(latG.GGt[5]) + ST(latG.GGt[7]), ST(latG.GGt[8])};
    const ST *g0 = w.myG.data(), *g1 = g0 + np, *g2 = g0 + 2 * np;
    const ST *h00 = w.myH.data(), *h01 = h00 + np, *h02 = h00 + 2 * np, *h11 = h00 + 3 * np, *h12 = h00 + 4 * np,
             *h22 = h00 + 5 * np;
    const int n = t.ns;
    for (int j = 0; j < n; ++j)
    {
      psi[j]          = VT(signed_one * w.myV[j]);
      dpsi[3 * j + 0] = VT(signed_one * (g00 * g0[j] + g01 * g1[j] + g02 * g2[j]));
      dpsi[3 * j + 1] = VT(signed_one * (g10 * g0[j] + g11 * g1[j] + g12 * g2[j]));
      dpsi[3 * j + 2] = VT(signed_one * (g20 * g0[j] + g21 * g1[j] + g22 * g2[j]));
      d2psi[j]        = VT(signed_one * SymTrace(h00[j], h01[j], h02[j], h11[j], h12[j], h22[j], symGG));
    }
  }

  // ---- from-scratch: psiM/dpsiM/d2psiM, inverse, Jastrows.  DiracDeterminantBatched.cpp:1122-1198
  void recomputeWalker(Walker& w)
  {
    for (int s = 0; s < 2; ++s)
    {
      Det& d = w.det[s];
      const int n = d.n;
      if (n == 0)
        continue;
      std::vector<VT> psiM((size_t)n * n), dtmp(3 * (size_t)n);
      for (int e = 0; e < n; ++e)
      {
        const RT* pos = &w.R[3 * (size_t)(d.first + e)];
        spoVGL(w, s, pos, &psiM[(size_t)e * n], &d.dpsiM[(size_t)e * n * 3], &d.d2psiM[(size_t)e * n]);
      }
      kernel_invert_transpose(psiM.data(), n, n, d.psiMinv.data(), d.lda, d.log_value);
      d.eng.delay_count = 0;
      d.invRow_id       = -1;
      d.curRatio        = PsiV(1);
    }
    if (has_j2)
      j2.recompute(w.j2, mi, w.rsoa.data(), npad_pos);
    if (has_j1)
      j1.recompute(w.j1, mi, w.rsoa.data(), npad_pos);
  }

  void recompute()
  {
#pragma omp parallel for schedule(dynamic)
    for (int iw = 0; iw < nw; ++iw)
      recomputeWalker(walkers[iw]);
    for (auto& c : crowds)
      c.delay_count[0] = c.delay_count[1] = 0;
  }

  // ---- determinant pieces (per walker)
  void prepareInvRow(Det& d, int row)
  {
    if (d.invRow_id != row)
    {
      d.eng.getInvRow(d.psiMinv.data(), d.lda, row, d.invRow.data());
      d.invRow_id = row;
    }
  }

  RT& w_rr_prop(Crowd& cr, int i) { return walkers[cr.w0 + i].rr_proposed; }

  // TrialWaveFunction::evalGrad + ParticleSet::makeMove + TrialWaveFunction::calcRatioGrad for ONE walker and a given
  // displacement, without accepting (TrialWaveFunction.cpp:568-833 in the order SlaterDet, J2, J1) -- the same blocks
  // advanceCrowd strings together, exposed so that the reference's own TrialWaveFunction tests (which prescribe the
  // displacement) can be replayed.  Gradients keep their complex value (the drift only uses the real part).
  void probeMove(int iw, int iat, const double displ[3], std::complex<double>& ratio_out, std::complex<double> grad_old[3],
                 std::complex<double> grad_new[3])
  {
    Walker& w     = walkers[iw];
    const int ig  = group_of(iat);
    const int row = iat - first_of(ig);
    Det& d        = w.det[ig];
    const int n   = d.n;
    prepareInvRow(d, row);
    VT g[3]      = {VT(0), VT(0), VT(0)};
    const VT* dp = &d.dpsiM[(size_t)row * n * 3];
    for (int j = 0; j < n; ++j)
      for (int dd = 0; dd < 3; ++dd)
        g[dd] += d.invRow[j] * dp[3 * j + dd];
    for (int dd = 0; dd < 3; ++dd)
    {
      std::complex<double> t = std::complex<double>(g[dd]);
      if (has_j2)
        t += (double)w.j2.dUat[dd * j2.npad + iat];
      if (has_j1)
        t += (double)w.j1.Grad[3 * iat + dd];
      grad_old[dd] = t;
    }
    for (int dd = 0; dd < 3; ++dd)
      w.newpos[dd] = w.R[3 * iat + dd] + (RT)displ[dd];
    if (has_j2)
    {
      mi.row(w.newpos, w.rsoa.data(), npad_pos, N, iat, w.dist_new.data());
      RT oldpos[3] = {w.rsoa[iat], w.rsoa[npad_pos + iat], w.rsoa[2 * npad_pos + iat]};
      mi.row(oldpos, w.rsoa.data(), npad_pos, N, iat, w.dist_old.data());
      w.dist_old[iat] = std::numeric_limits<RT>::max();
    }
    std::vector<VT> psi(n), dpsi(3 * (size_t)n), d2psi(n);
    spoVGL(w, ig, w.newpos, psi.data(), dpsi.data(), d2psi.data());
    VT ratio(0), gn[3] = {VT(0), VT(0), VT(0)};
    for (int j = 0; j < n; ++j)
    {
      ratio += d.invRow[j] * psi[j];
      for (int dd = 0; dd < 3; ++dd)
        gn[dd] += d.invRow[j] * dpsi[3 * j + dd];
    }
    std::complex<double> r = std::complex<double>(static_cast<PsiV>(ratio));
    RT gj[3]               = {RT(0), RT(0), RT(0)};
    if (has_j2)
      r *= (double)j2.ratioGrad(w.j2, iat, w.dist_new.data(), gj);
    if (has_j1)
      r *= (double)j1.ratioGrad(w.j1, mi, iat, w.newpos, gj);
    ratio_out = r;
    for (int dd = 0; dd < 3; ++dd)
      grad_new[dd] = std::complex<double>(gn[dd] / ratio) + (double)gj[dd];
  }

  // TrialWaveFunction::mw_evaluateRatios body for ONE walker (TrialWaveFunction.cpp:1079-1110): the ratios
  // psi(R with electron `ref` at r_vp[k]) / psi(R) at the nk virtual positions of a VirtualParticleSet (the quadrature
  // points of the non-local pseudopotential), as the product over the selected components:
  //   determinant  DiracDeterminantBatched::evaluateRatios (DiracDeterminantBatched.cpp:790-797): psiMinv row . psi(r_vp)
  //   J2           TwoBodyJastrow::evaluateRatios (TwoBodyJastrow.cpp:166-171): exp(Uat[ref] - computeU(dist row of r_vp))
  //                with computeU = sum_{j != ref} F[group(ref)][group(j)].evaluate(r_j) (value only, BsplineFunctor.cpp:135-200)
  //   J1           J1OrbitalSoA::evaluateRatios (J1OrbitalSoA.h:295-299): exp(Vat[ref] - computeU(e-ion row of r_vp))
  // ct: 0 ALL, 1 FERMIONIC, 2 NONFERMIONIC (TrialWaveFunction::ComputeType)
  void evaluateRatios(int iw, int ref, int nk, const double* r_vp, int ct, std::complex<double>* out)
  {
    Walker& w     = walkers[iw];
    const int ig  = group_of(ref);
    const int row = ref - first_of(ig);
    Det& d        = w.det[ig];
    const int n   = d.n;
    if (ct != 2)
      prepareInvRow(d, row); // (= the row of psiMinv once the delayed updates are complete)
    std::vector<VT> psi(n), dpsi(3 * (size_t)n), d2psi(n);
    std::vector<RT> dist(4 * npad_pos);
    for (int kq = 0; kq < nk; ++kq)
    {
      const RT pos[3] = {(RT)r_vp[3 * kq], (RT)r_vp[3 * kq + 1], (RT)r_vp[3 * kq + 2]};
      std::complex<double> r(1.0, 0.0);
      if (ct != 2)
      {
        spoVGL(w, ig, pos, psi.data(), dpsi.data(), d2psi.data());
        VT ratio(0);
        for (int j = 0; j < n; ++j)
          ratio += d.invRow[j] * psi[j];
        r *= std::complex<double>(static_cast<PsiV>(ratio));
      }
      if (ct != 1 && has_j2)
      {
        mi.row(pos, w.rsoa.data(), npad_pos, N, ref, dist.data());
        RT usum(0);
        const int igt = j2.grp_ids[ref] * j2.ngroups;
        for (int j = 0; j < N; ++j)
        {
          if (j == ref)
            continue;
          const BsplineFunctor<RT>& f = j2.F[igt + j2.grp_ids[j]];
          RT du, d2u;
          if (f.present && dist[j] < f.cutoff_radius)
            usum += f.evaluate_impl(dist[j], du, d2u);
        }
        r *= std::exp(static_cast<double>(w.j2.Uat[ref] - usum));
      }
      if (ct != 1 && has_j1)
      {
        RT at, lap, g[3];
        j1.compute(mi, pos, at, lap, g);
        r *= std::exp(static_cast<double>(w.j1.Vat[ref] - at));
      }
      out[kq] = r;
    }
  }

  // one sweep step for one crowd
  // forced (optional, [N][nw]): accept flags imposed from outside ("teacher forcing" for mixed-precision parity: the
  // uniform is still drawn under the reference's rule so that the stream stays aligned); ratio_log (optional, [N][nw]):
  // the total wavefunction ratio of every proposed move
  void advanceCrowd(Crowd& cr, uint8_t* acc_log /* [N][nw] or null */, const uint8_t* forced = nullptr,
                    double* ratio_log = nullptr)
  {
    const int cw = cr.w1 - cr.w0;
    if (cw == 0)
      return;
    const bool use_drift = prm.use_drift != 0;
    std::vector<RT> log_gf(cw, RT(0)), log_gb(cw, RT(0)), prob(cw);
    std::vector<PsiV> ratios(cw);
    std::vector<RT> deltas(3 * (size_t)cw), drifts(3 * (size_t)cw), grads_now(3 * (size_t)cw),
        grads_new(3 * (size_t)cw);
    std::vector<char> isAccepted(cw);

    // makeGaussRandomWithEngine(walker_deltas, rng): nw*N*3 Gaussians in one go (VMCBatched.cpp:109)
    cr.walker_deltas.resize(3 * (size_t)cw * N);
    assignGaussRand(cr.walker_deltas.data(), (unsigned)cr.walker_deltas.size(), cr.rng);
    const bool dmc = prm.dmc != 0;
    if (dmc)
      for (int i = 0; i < cw; ++i)
        walkers[cr.w0 + i].rr_accepted = walkers[cr.w0 + i].rr_proposed = RT(0);
    std::vector<RT> rr(cw, RT(0));

    for (int ig = 0; ig < 2; ++ig)
    {
      // TauParams (QMCDrivers/TauParams.hpp:29-40), mass = 1
      const RT tauovermass = RT(prm.tau) * RT(1.0);
      const RT oneover2tau = 0.5 / (tauovermass);
      const RT sqrttau     = std::sqrt(tauovermass);
      for (int iat = first_of(ig); iat < last_of(ig); ++iat)
      {
        const int row = iat - first_of(ig);
        // deltas for this particle: walker_deltas[iat*cw + iw] (VMCBatched.cpp:122), scaled by sqrt(tau)
        for (int i = 0; i < cw; ++i)
        {
          if (dmc)
          {
            // rr = tauovermass * dot(delta_r, delta_r) of the raw Gaussians (DMCBatched.cpp:163-167)
            const RT* dr = &cr.walker_deltas[3 * ((size_t)iat * cw + i)];
            rr[i]        = tauovermass * (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
          }
          for (int d = 0; d < 3; ++d)
            deltas[3 * i + d] = cr.walker_deltas[3 * ((size_t)iat * cw + i) + d] * sqrttau;
        }

        if (use_drift)
        {
          // TWF::mw_evalGrad: components in order SlaterDet, J2, J1; grads summed in ValueType
          for (int i = 0; i < cw; ++i)
          {
            Walker& w = walkers[cr.w0 + i];
            Det& d    = w.det[ig];
            prepareInvRow(d, row);
            VT g[3] = {VT(0), VT(0), VT(0)};
            const VT* dp = &d.dpsiM[(size_t)row * d.n * 3];
            for (int j = 0; j < d.n; ++j)
            {
              g[0] += d.invRow[j] * dp[3 * j];
              g[1] += d.invRow[j] * dp[3 * j + 1];
              g[2] += d.invRow[j] * dp[3 * j + 2];
            }
            // complex gradients reach the drift through convertToReal (DriftModifierUNR.cpp:20-23): real part
            RT gt[3] = {real_of(g[0]), real_of(g[1]), real_of(g[2])};
            if (has_j2)
              for (int dd = 0; dd < 3; ++dd)
                gt[dd] += w.j2.dUat[dd * j2.npad + iat]; // TwoBodyJastrow::evalGrad (:524-528)
            if (has_j1)
              for (int dd = 0; dd < 3; ++dd)
                gt[dd] += w.j1.Grad[3 * iat + dd];
            for (int dd = 0; dd < 3; ++dd)
              grads_now[3 * i + dd] = gt[dd];
          }
          for (int i = 0; i < cw; ++i)
          {
            getDrift<RT>(tauovermass, &grads_now[3 * i], &drifts[3 * i]);
            for (int d = 0; d < 3; ++d)
              drifts[3 * i + d] += deltas[3 * i + d];
          }
        }
        else
          drifts = deltas;

        // ParticleSet::mw_makeMove: newpos = R[iat] + displ; distance-table temp/old rows
        for (int i = 0; i < cw; ++i)
        {
          Walker& w = walkers[cr.w0 + i];
          for (int d = 0; d < 3; ++d)
            w.newpos[d] = w.R[3 * iat + d] + drifts[3 * i + d];
          if (has_j2)
          {
            mi.row(w.newpos, w.rsoa.data(), npad_pos, N, iat, w.dist_new.data());
            // note: dist_new needs stride j2.npad; npad_pos == j2.npad (both aligned_size<RT>(N))
            RT oldpos[3] = {w.rsoa[iat], w.rsoa[npad_pos + iat], w.rsoa[2 * npad_pos + iat]};
            mi.row(oldpos, w.rsoa.data(), npad_pos, N, iat, w.dist_old.data());
            w.dist_old[iat] = std::numeric_limits<RT>::max();
          }
        }

        // TWF::mw_calcRatioGrad
        for (int i = 0; i < cw; ++i)
        {
          Walker& w = walkers[cr.w0 + i];
          Det& d    = w.det[ig];
          const int n = d.n;
          VT* psi   = w.phi_vgl.data();
          std::vector<VT> dpsi(3 * (size_t)n);
          VT* d2psi = w.phi_vgl.data() + 4 * (size_t)n;
          spoVGL(w, ig, w.newpos, psi, dpsi.data(), d2psi);
          for (int j = 0; j < n; ++j)
          {
            w.phi_vgl[1 * (size_t)n + j] = dpsi[3 * j];
            w.phi_vgl[2 * (size_t)n + j] = dpsi[3 * j + 1];
            w.phi_vgl[3 * (size_t)n + j] = dpsi[3 * j + 2];
          }
          prepareInvRow(d, row);
          VT ratio(0), gx(0), gy(0), gz(0);
          for (int j = 0; j < n; ++j)
            ratio += d.invRow[j] * psi[j];
          for (int j = 0; j < n; ++j)
          {
            gx += d.invRow[j] * dpsi[3 * j];
            gy += d.invRow[j] * dpsi[3 * j + 1];
            gz += d.invRow[j] * dpsi[3 * j + 2];
          }
          d.curRatio = static_cast<PsiV>(ratio);
          RT gn[3]   = {real_of(gx / ratio), real_of(gy / ratio), real_of(gz / ratio)};
          PsiV r     = d.curRatio;
          if (has_j2)
            r *= (double)j2.ratioGrad(w.j2, iat, w.dist_new.data(), gn);
          if (has_j1)
            r *= (double)j1.ratioGrad(w.j1, mi, iat, w.newpos, gn);
          ratios[i] = r;
          for (int dd = 0; dd < 3; ++dd)
            grads_new[3 * i + dd] = gn[dd];
        }

        if (use_drift)
        {
          for (int i = 0; i < cw; ++i)
          {
            const RT* dl = &deltas[3 * i];
            log_gf[i]    = -oneover2tau * (dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
            RT dr[3];
            getDrift<RT>(tauovermass, &grads_new[3 * i], dr);
            for (int d = 0; d < 3; ++d)
              dr[d] += drifts[3 * i + d];
            log_gb[i] = -oneover2tau * (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
          }
        }
        for (int i = 0; i < cw; ++i)
          prob[i] = (RT)std::norm(ratios[i]); // VMCBatched.cpp:152 (squared modulus; x*x for a real PsiValue)

        // accept test: RNG consumed only if the first two conditions hold (VMCBatched.cpp:156-158)
        for (int i = 0; i < cw; ++i)
        {
          const bool valid = true; // periodic cell: every move is valid
          if (dmc)
          {
            // DMCBatched.cpp:188-222: checkPhaseChanged (ratio == 0 or SFNBranch::phaseChanged(arg ratio): real
            // wavefunctions reject when cos(arg) < eps, complex builds never), prob = norm(ratio) exp(log_gb - log_gf)
            bool reject = ratios[i] == PsiV(0);
            if constexpr (!is_cplx)
              reject = reject || std::cos(std::arg(ratios[i])) < std::numeric_limits<RT>::epsilon();
            w_rr_prop(cr, i) += rr[i];
            const RT pdmc = (RT)(std::norm(ratios[i]) * std::exp(log_gb[i] - log_gf[i]));
            if (valid && !reject && pdmc >= std::numeric_limits<RT>::epsilon() && cr.rng() < pdmc)
            {
              isAccepted[i] = 1;
              walkers[cr.w0 + i].rr_accepted += rr[i];
            }
            else
              isAccepted[i] = 0;
          }
          else if (valid && prob[i] >= std::numeric_limits<RT>::epsilon() &&
                   cr.rng() < prob[i] * std::exp(log_gb[i] - log_gf[i]))
            isAccepted[i] = 1;
          else
            isAccepted[i] = 0;
          if (forced)
            isAccepted[i] = forced[(size_t)iat * nw + cr.w0 + i];
          if (ratio_log)
          {
            if constexpr (is_cplx)
            {
              ratio_log[2 * ((size_t)iat * nw + cr.w0 + i)]     = ratios[i].real();
              ratio_log[2 * ((size_t)iat * nw + cr.w0 + i) + 1] = ratios[i].imag();
            }
            else
              ratio_log[(size_t)iat * nw + cr.w0 + i] = ratios[i];
          }
        }

        // TWF::mw_accept_rejectMove -> determinant, J2, J1; then ParticleSet::mw_accept_rejectMove
        const bool batched = prm.batched_engine != 0;
        for (int i = 0; i < cw; ++i)
        {
          Walker& w = walkers[cr.w0 + i];
          Det& d    = w.det[ig];
          d.invRow_id = -1;
          if (isAccepted[i])
          {
            if (d.curRatio == PsiV(0))
              throw std::runtime_error("oracle: accepted move with curRatio == 0");
            d.log_value += std::log(std::complex<double>(d.curRatio));
            const int n = d.n;
            // save G/L rows (add_delay_list_save_sigma_VGL: AccelMatrixUpdateOMPTarget.hpp:125-136)
            VT* dp = &d.dpsiM[(size_t)row * n * 3];
            for (int j = 0; j < n; ++j)
            {
              dp[3 * j]     = w.phi_vgl[1 * (size_t)n + j];
              dp[3 * j + 1] = w.phi_vgl[2 * (size_t)n + j];
              dp[3 * j + 2] = w.phi_vgl[3 * (size_t)n + j];
              d.d2psiM[(size_t)row * n + j] = w.phi_vgl[4 * (size_t)n + j];
            }
            // batched engine: sigma = Value(1)/ratios_local (VT arithmetic, DelayedUpdateBatched.h:599);
            // CPU engine: RATIOT = PsiValue(double)
            if (batched)
              d.eng.acceptRow(d.psiMinv.data(), d.lda, row, w.phi_vgl.data(), static_cast<VT>(d.curRatio));
            else
              d.eng.acceptRow(d.psiMinv.data(), d.lda, row, w.phi_vgl.data(), d.curRatio);
            if (has_j2)
              j2.acceptMove(w.j2, iat, w.dist_new.data(), w.dist_old.data());
            if (has_j1)
              j1.acceptMove(w.j1, iat);
            for (int dd = 0; dd < 3; ++dd)
            {
              w.R[3 * iat + dd]             = w.newpos[dd];
              w.rsoa[dd * npad_pos + iat] = w.newpos[dd];
            }
            w.n_accept++;
          }
          else
          {
            if (batched && prm.delay_rank > 1)
              d.eng.pseudoAcceptRow(d.psiMinv.data(), d.lda, row);
            w.n_reject++;
          }
          d.curRatio = PsiV(1);
          if (acc_log)
            acc_log[(size_t)iat * nw + cr.w0 + i] = isAccepted[i];
        }
      }
    }
    // flex_completeUpdates (DiracDeterminantBatched.cpp:545-591): flush pending delays
    for (int i = 0; i < cw; ++i)
      for (int s = 0; s < 2; ++s)
      {
        Det& d = walkers[cr.w0 + i].det[s];
        d.eng.updateInvMat(d.psiMinv.data(), d.lda);
        d.invRow_id = -1;
      }
  }

  void sweep(int nsteps, bool log_accept, const uint8_t* forced = nullptr, double* ratio_log = nullptr)
  {
    if (log_accept)
      accept_log.assign((size_t)nsteps * N * nw, 0);
    for (int step = 0; step < nsteps; ++step)
    {
      uint8_t* lg       = log_accept ? accept_log.data() + (size_t)step * N * nw : nullptr;
      const uint8_t* fc = forced ? forced + (size_t)step * N * nw : nullptr;
      double* rl        = ratio_log ? ratio_log + (size_t)step * N * nw * (is_cplx ? 2 : 1) : nullptr;
#pragma omp parallel for schedule(static, 1)
      for (int c = 0; c < (int)crowds.size(); ++c)
        advanceCrowd(crowds[c], lg, fc, rl);
    }
  }

  // ---- G, L of all particles and kinetic energy -1/2 sum(L + G.G)  (mw_evaluateGL, fromscratch=false;
  // DiracDeterminantBatched.cpp:594-604 computeGL; TwoBodyJastrow.cpp:769-808; J1OrbitalSoA.h:112-123;
  // QMCHamiltonians/BareKineticEnergy: KE = -1/2 (sum L + sum G.G), L = laplacian of log psi)
  // complex value type: G [N][3][2], L [N][2] interleaved; KE = -1/2 (Sum(L) + Dot(G,G)) with the reference's complex
  // form real(CplxDot(G,G) + CplxSum(L)) = sum(re*re - im*im) + sum(re L) (QMCHamiltonians/BareKineticEnergy.cpp:114,
  // Platforms/CPU/VectorOps.h:141-152)
  void evaluateGL(int iw, double* G /*[N][3]*/, double* L /*[N]*/, double* logpsi, double* ke)
  {
    Walker& w = walkers[iw];
    std::vector<VT> g(3 * (size_t)N, VT(0)), l(N, VT(0));
    for (int s = 0; s < 2; ++s)
    {
      Det& d = w.det[s];
      for (int i = 0; i < d.n; ++i)
      {
        const VT* inv = &d.psiMinv[(size_t)i * d.lda];
        const VT* dp  = &d.dpsiM[(size_t)i * d.n * 3];
        const VT* d2  = &d.d2psiM[(size_t)i * d.n];
        VT rv[3] = {VT(0), VT(0), VT(0)}, lap(0);
        for (int j = 0; j < d.n; ++j)
        {
          rv[0] += inv[j] * dp[3 * j];
          rv[1] += inv[j] * dp[3 * j + 1];
          rv[2] += inv[j] * dp[3 * j + 2];
        }
        for (int j = 0; j < d.n; ++j)
          lap += inv[j] * d2[j];
        const int iat = d.first + i;
        for (int dd = 0; dd < 3; ++dd)
          g[3 * iat + dd] += rv[dd];
        l[iat] += lap - (rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
      }
    }
    if (has_j2)
      for (int i = 0; i < N; ++i)
      {
        for (int dd = 0; dd < 3; ++dd)
          g[3 * i + dd] += w.j2.dUat[dd * j2.npad + i];
        l[i] += w.j2.d2Uat[i];
      }
    if (has_j1)
      for (int i = 0; i < N; ++i)
      {
        for (int dd = 0; dd < 3; ++dd)
          g[3 * i + dd] += w.j1.Grad[3 * i + dd];
        l[i] -= w.j1.Lap[i];
      }
    double kin = 0;
    for (int i = 0; i < N; ++i)
    {
      if constexpr (is_cplx)
      {
        kin += (double)l[i].real();
        for (int dd = 0; dd < 3; ++dd)
        {
          const double re = g[3 * i + dd].real(), im = g[3 * i + dd].imag();
          kin += re * re - im * im;
          if (G)
          {
            G[2 * (3 * i + dd)]     = re;
            G[2 * (3 * i + dd) + 1] = im;
          }
        }
        if (L)
        {
          L[2 * i]     = l[i].real();
          L[2 * i + 1] = l[i].imag();
        }
      }
      else
      {
        kin += (double)l[i] + (double)g[3 * i] * g[3 * i] + (double)g[3 * i + 1] * g[3 * i + 1] +
            (double)g[3 * i + 2] * g[3 * i + 2];
        if (G)
          for (int dd = 0; dd < 3; ++dd)
            G[3 * i + dd] = g[3 * i + dd];
        if (L)
          L[i] = l[i];
      }
    }
    if (ke)
      *ke = -0.5 * kin;
    if (logpsi)
      *logpsi = w.det[0].log_value.real() + w.det[1].log_value.real() + (has_j2 ? w.j2.log_value : 0.0) +
          (has_j1 ? w.j1.log_value : 0.0);
  }
};

} // namespace orc
