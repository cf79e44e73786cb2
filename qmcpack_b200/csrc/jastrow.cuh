// qmcpack_b200/csrc/jastrow.cuh -- distance-table rows, two-body and one-body B-spline Jastrow kernels (sm_100a).
//
// Replaces (reference paths under /root/reference/src):
//   SoaDistanceTableAAOMPTarget::mw_move        Particle/SoaDistanceTableAAOMPTarget.h:265-371     (K17)
//   DTD_BConds::computeDistancesOffload         Particle/Lattice/ParticleBConds3DSoa.h:141-168 (PPPO), :452-510 (PPPG)
//   BsplineFunctor::mw_evaluateVGL / mw_updateVGL  QMCWaveFunctions/Jastrow/BsplineFunctor.cpp:26-131, :203-326   (K16)
//   TwoBodyJastrow::mw_ratioGrad / mw_accept_rejectMove / recompute   Jastrow/TwoBodyJastrow.cpp:542-578, :631-664, :667-713
//   J1OrbitalSoA ratioGrad / acceptMove / recompute   Jastrow/J1OrbitalSoA.h:136-185, :441-469 (host loop in the reference)
// One CTA per walker; the five sums of a move are reduced with warp shuffles in a fixed order.  The temp/old rows and
// the per-pair u, u'/r, u'' of the proposed row stay in HBM between the ratio and the accept kernels exactly like
// mw_new_old_dist_displ / mw_cur_allu of the reference (TwoBodyJastrow.cpp:63-86).
#pragma once
#include "common.cuh"

namespace qmcb
{
template<typename RT>
struct FunctorDev
{
  const RT* coefs;
  RT DeltaRInv, rcut;
  int max_index;
};

template<typename RT>
struct CellDev
{
  int ortho;
  RT L[3], Linv[3];
  RT r[9], g[9];
  RT corners[3][8];
};

template<typename RT>
struct JastrowDev
{
  int N, npad, n_up, nw;
  CellDev<RT> cell;
  // J2
  int has_j2;
  FunctorDev<RT> F2[4];     // [group(iat)*2 + group(j)]
  RT* rsoa;                 // [nw][3][npad]  committed positions
  RT* newpos;               // [nw][3]
  RT* rows;                 // [2][nw][4][npad]  new then old: r, dx, dy, dz
  RT* cur_allu;             // [nw][3][npad]
  RT* j2_vgl;               // [nw][5]  cur_Uat, grad(3), -lapl
  RT* Uat;                  // [nw][npad]
  RT* dUat;                 // [nw][3][npad]
  RT* d2Uat;                // [nw][npad]
  double* j2_log;           // [nw]
  // J1
  int has_j1, nions, npad_ion;
  const RT* ion_rsoa;       // [3][npad_ion]
  const int* ion_grp;       // [nions]
  FunctorDev<RT> F1[8];
  RT* j1_cur;               // [nw][5]  curAt, curGrad(3), curLap
  RT* Vat;                  // [nw][N]
  RT* Grad1;                // [nw][3][N]
  RT* Lap1;                 // [nw][N]
  double* j1_log;           // [nw]
};

#ifdef __CUDACC__
constexpr int JAS_TPB = 256;

// ref: Numerics/SplineBound.hpp:37-62 with T = RT
template<typename RT>
__device__ __forceinline__ void spline_bound_rt(RT x, int nmax, int& ind, RT& dx)
{
  if (x < 0)
  {
    ind = 0;
    dx  = RT(0);
  }
  else
  {
    RT ipart;
    dx  = modf(x, &ipart);
    ind = (int)ipart;
    if (ind > nmax)
    {
      ind = nmax;
      dx  = RT(1) - (sizeof(RT) == 4 ? RT(1.1920929e-07f) : RT(2.220446049250313e-16));
    }
  }
}

#ifdef QMCB_JAS_NOINLINE
#define QMCB_JAS_INLINE __noinline__
#else
#define QMCB_JAS_INLINE __forceinline__
#endif
// ref: BsplineFunctor.h:254-285 evaluate_impl; returns u, sets du = u'/r and d2u = u'' (zero beyond the cutoff)
template<typename RT>
__device__ QMCB_JAS_INLINE RT functor_eval(const FunctorDev<RT>& f, RT r, RT& du_over_r, RT& d2u)
{
  RT u(0);
  du_over_r = RT(0);
  d2u       = RT(0);
  if (f.coefs != nullptr && r < f.rcut)
  {
    RT rr = r * f.DeltaRInv;
    int i;
    RT t;
    spline_bound_rt(rr, f.max_index, i, t);
    const RT c0 = f.coefs[i], c1 = f.coefs[i + 1], c2 = f.coefs[i + 2], c3 = f.coefs[i + 3];
    d2u = f.DeltaRInv * f.DeltaRInv *
        (c0 * (RT(-1.0) * t + RT(1.0)) + c1 * (RT(3.0) * t + RT(-2.0)) + c2 * (RT(-3.0) * t + RT(1.0)) +
         c3 * (RT(1.0) * t + RT(0.0)));
    RT dudr = f.DeltaRInv *
        (c0 * ((RT(-0.5) * t + RT(1.0)) * t + RT(-0.5)) + c1 * ((RT(1.5) * t + RT(-2.0)) * t + RT(0.0)) +
         c2 * ((RT(-1.5) * t + RT(1.0)) * t + RT(0.5)) + c3 * ((RT(0.5) * t + RT(0.0)) * t + RT(0.0)));
    u = (c0 * (((RT(-1.0 / 6.0) * t + RT(3.0 / 6.0)) * t + RT(-3.0 / 6.0)) * t + RT(1.0 / 6.0)) +
         c1 * (((RT(3.0 / 6.0) * t + RT(-6.0 / 6.0)) * t + RT(0.0 / 6.0)) * t + RT(4.0 / 6.0)) +
         c2 * (((RT(-3.0 / 6.0) * t + RT(3.0 / 6.0)) * t + RT(3.0 / 6.0)) * t + RT(1.0 / 6.0)) +
         c3 * (((RT(1.0 / 6.0) * t + RT(0.0 / 6.0)) * t + RT(0.0 / 6.0)) * t + RT(0.0 / 6.0)));
    du_over_r = dudr * (RT(1) / r);
  }
  return u;
}

// minimum-image displacement src - pos.  ref: ParticleBConds3DSoa.h:141-168 (ortho), :452-510 (general)
template<typename RT>
__device__ QMCB_JAS_INLINE void min_image(const CellDev<RT>& C, const RT pos[3], RT px, RT py, RT pz, int iel, int flip_ind,
                                          RT& rr, RT& dx, RT& dy, RT& dz)
{
  if (C.ortho)
  {
    const RT x = (px - pos[0]) * C.Linv[0];
    const RT y = (py - pos[1]) * C.Linv[1];
    const RT z = (pz - pos[2]) * C.Linv[2];
    dx         = C.L[0] * (x - round(x));
    dy         = C.L[1] * (y - round(y));
    dz         = C.L[2] * (z - round(z));
    rr         = sqrt(dx * dx + dy * dy + dz * dz);
  }
  else
  {
    const RT flip    = iel < flip_ind ? RT(1) : RT(-1);
    const RT displ_0 = (px - pos[0]) * flip;
    const RT displ_1 = (py - pos[1]) * flip;
    const RT displ_2 = (pz - pos[2]) * flip;
    const RT ar_0    = -floor(displ_0 * C.g[0] + displ_1 * C.g[3] + displ_2 * C.g[6]);
    const RT ar_1    = -floor(displ_0 * C.g[1] + displ_1 * C.g[4] + displ_2 * C.g[7]);
    const RT ar_2    = -floor(displ_0 * C.g[2] + displ_1 * C.g[5] + displ_2 * C.g[8]);
    const RT delx    = displ_0 + ar_0 * C.r[0] + ar_1 * C.r[3] + ar_2 * C.r[6];
    const RT dely    = displ_1 + ar_0 * C.r[1] + ar_1 * C.r[4] + ar_2 * C.r[7];
    const RT delz    = displ_2 + ar_0 * C.r[2] + ar_1 * C.r[5] + ar_2 * C.r[8];
    RT rmin          = delx * delx + dely * dely + delz * delz;
    int ic           = 0;
#pragma unroll
    for (int c = 1; c < 8; ++c)
    {
      const RT x  = delx + C.corners[0][c];
      const RT y  = dely + C.corners[1][c];
      const RT z  = delz + C.corners[2][c];
      const RT r2 = x * x + y * y + z * z;
      ic          = (r2 < rmin) ? c : ic;
      rmin        = (r2 < rmin) ? r2 : rmin;
    }
    rr = sqrt(rmin);
    dx = flip * (delx + C.corners[0][ic]);
    dy = flip * (dely + C.corners[1][ic]);
    dz = flip * (delz + C.corners[2][ic]);
  }
}

// Cheap SUPERSET test for the cutoff lists (orthorhombic cells): squared minimum-image distance with round-to-nearest-even
// and no square root, against rcut^2 (1 + a few ulp).  Whatever passes is evaluated with min_image / functor_eval -- the
// reference's arithmetic, which applies the exact r < rcut test itself -- so the sums are unchanged; a tie of the rounding
// (|d| = L/2) is beyond every cutoff.  About a third fewer instructions per candidate than min_image.
template<typename RT>
__device__ __forceinline__ RT min_image_r2_ortho(const CellDev<RT>& C, const RT pos[3], const RT px, const RT py, const RT pz)
{
  RT x = (px - pos[0]) * C.Linv[0], y = (py - pos[1]) * C.Linv[1], z = (pz - pos[2]) * C.Linv[2];
  x    = C.L[0] * (x - rint(x));
  y    = C.L[1] * (y - rint(y));
  z    = C.L[2] * (z - rint(z));
  return x * x + y * y + z * z;
}
#ifndef QMCB_JAS_FAST
#define QMCB_JAS_FAST 1
#endif
template<typename RT>
__device__ __forceinline__ RT rcut2_inflated(const RT rcut)
{
  return rcut * rcut * (RT(1) + RT(16) * (sizeof(RT) == 4 ? RT(1.1920929e-07f) : RT(2.220446049250313e-16)));
}

// one-body sums at position pos: at = sum u, lap = sum(u'' + 2u'/r), grad = sum (u'/r) d   (J1OrbitalSoA.h:136-185)
// must be called by every thread of the CTA
template<typename RT>
__device__ __forceinline__ void j1_sums(const JastrowDev<RT>& J, const RT pos[3], RT out[5], RT* red)
{
  RT acc[5] = {RT(0), RT(0), RT(0), RT(0), RT(0)};
  for (int j = threadIdx.x; j < J.nions; j += blockDim.x)
  {
    RT r, dx, dy, dz, du, d2u;
    min_image(J.cell, pos, J.ion_rsoa[j], J.ion_rsoa[J.npad_ion + j], J.ion_rsoa[2 * J.npad_ion + j], j, 0, r, dx, dy, dz);
    const RT u = functor_eval(J.F1[J.ion_grp[j]], r, du, d2u);
    acc[0] += u;
    acc[1] += du * dx;
    acc[2] += du * dy;
    acc[3] += du * dz;
    acc[4] += d2u + RT(2) * du;
  }
  block_sum<RT, 5>(acc, red);
#pragma unroll
  for (int e = 0; e < 5; ++e)
    out[e] = acc[e];
}

// ---- cutoff lists.  The functors vanish beyond rcut (5.57 bohr in a 15.76 bohr cell at NiO-a64: 18 % of the pairs are
// inside), so evaluating every pair keeps 4 of 5 lanes busy with zeros and the accept rewrites 15 KB of unchanged sums.
// Pass 1 computes the distances of all pairs and every WARP compacts the indices of its in-range pairs (ballot + popc)
// into its own segment of a shared-memory list; pass 2 evaluates the functors with full warps over the list.  Skipped
// pairs contribute exact zeros, the list order is fixed by (iteration, lane), and no atomics are involved, so the sums
// stay deterministic.  JAS_LIST entries are enough for N + N_ion <= 4096 - (threads); larger systems take the direct path.
constexpr int JAS_LIST = 4608;

// ---- proposed move: J2 sums over the temporary distance row and J1 sums at the proposed position.
// The reference materialises the new AND the old distance row (mw_new_old_dist_displ) and the per-pair functor values
// (mw_cur_allu) in device memory and reads them back in the accept kernel: 54 KB written + 45 KB read per walker per
// move at NiO-a64.  Here the rows are a by-product that is only stored on request (STORE, for qmcb_dtaa_get_temp_rows);
// the accept recomputes the two rows from the positions (9 KB) -- identical arithmetic, so identical values -- and the
// J1 and J2 sums share one pass and one block reduction.
// Body for one walker, executed by ALL threads of the CTA (blockDim.x threads; red >= 10 * 32; jl = JAS_LIST entries).
template<typename RT, bool STORE>
__device__ __forceinline__ void jastrow_move_body(const JastrowDev<RT>& J, const int iw, const int iat, const RT pos[3], RT* red,
                                                  unsigned short* jl)
{
  const int tid = threadIdx.x, N = J.N, np = J.npad, JAS_STEP = blockDim.x;
  const int lane = tid & 31, wg = tid >> 5, nwg = JAS_STEP >> 5;
  const RT* rs = J.rsoa + (size_t)iw * 3 * np;
  RT acc[10]   = {RT(0), RT(0), RT(0), RT(0), RT(0), RT(0), RT(0), RT(0), RT(0), RT(0)};
  const int n2 = J.has_j2 ? N : 0, n1 = J.has_j1 ? J.nions : 0; // combined index space: electrons then ions
  const int gi = (iat < J.n_up ? 0 : 1) * 2;
  RT old[3]    = {RT(0), RT(0), RT(0)};
  RT* rnew     = J.rows + (size_t)iw * 4 * np;
  RT* rold     = J.rows + ((size_t)J.nw + iw) * 4 * np;
  if (STORE && J.has_j2)
    old[0] = rs[iat], old[1] = rs[np + iat], old[2] = rs[2 * np + iat];
  const int iters  = (n2 + n1 + JAS_STEP - 1) / JAS_STEP;
  const int segcap = iters * 32;
  const bool listed = segcap * nwg <= JAS_LIST && n2 + n1 < 65536;
  unsigned short* seg = jl + wg * segcap;
  int cnt = 0;
  for (int it = 0; it < iters; ++it)
  {
    const int idx = it * JAS_STEP + tid;
    bool need     = false;
    RT r = RT(0), dx = RT(0), dy = RT(0), dz = RT(0);
    if (idx < n2)
    {
      const int j = idx;
      const RT px = rs[j], py = rs[np + j], pz = rs[2 * np + j];
      min_image(J.cell, pos, px, py, pz, j, iat, r, dx, dy, dz);
      if (STORE)
      {
        rnew[j]          = r;
        rnew[np + j]     = dx;
        rnew[2 * np + j] = dy;
        rnew[3 * np + j] = dz;
        RT ro, ox, oy, oz;
        min_image(J.cell, old, px, py, pz, j, iat, ro, ox, oy, oz);
        rold[j]          = (j == iat) ? (sizeof(RT) == 4 ? RT(3.402823466e+38f) : RT(1.7976931348623157e+308)) : ro;
        rold[np + j]     = ox;
        rold[2 * np + j] = oy;
        rold[3 * np + j] = oz;
      }
      const FunctorDev<RT>& F = J.F2[gi + (j < J.n_up ? 0 : 1)];
      need                    = j != iat && F.coefs != nullptr && r < F.rcut;
    }
    else if (idx < n2 + n1)
    {
      const int j = idx - n2;
      min_image(J.cell, pos, J.ion_rsoa[j], J.ion_rsoa[J.npad_ion + j], J.ion_rsoa[2 * J.npad_ion + j], j, 0, r, dx, dy, dz);
      const FunctorDev<RT>& F = J.F1[J.ion_grp[j]];
      need                    = F.coefs != nullptr && r < F.rcut;
    }
    if (listed)
    {
      const unsigned m = __ballot_sync(0xffffffffu, need);
      if (need)
        seg[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)idx;
      cnt += __popc(m);
    }
    else if (need)
    {
      RT du, d2u;
      if (idx < n2)
      {
        const RT u = functor_eval(J.F2[gi + (idx < J.n_up ? 0 : 1)], r, du, d2u);
        acc[0] += u;
        acc[1] += du * dx;
        acc[2] += du * dy;
        acc[3] += du * dz;
        acc[4] += d2u + RT(2) * du;
      }
      else
      {
        const RT u = functor_eval(J.F1[J.ion_grp[idx - n2]], r, du, d2u);
        acc[5] += u;
        acc[6] += du * dx;
        acc[7] += du * dy;
        acc[8] += du * dz;
        acc[9] += d2u + RT(2) * du;
      }
    }
  }
  if (listed)
  {
    __syncwarp();
    for (int e = lane; e < cnt; e += 32)
    {
      const int idx = seg[e];
      RT r, dx, dy, dz, du, d2u;
      if (idx < n2)
      {
        min_image(J.cell, pos, rs[idx], rs[np + idx], rs[2 * np + idx], idx, iat, r, dx, dy, dz);
        const RT u = functor_eval(J.F2[gi + (idx < J.n_up ? 0 : 1)], r, du, d2u);
        acc[0] += u;
        acc[1] += du * dx;
        acc[2] += du * dy;
        acc[3] += du * dz;
        acc[4] += d2u + RT(2) * du;
      }
      else
      {
        const int j = idx - n2;
        min_image(J.cell, pos, J.ion_rsoa[j], J.ion_rsoa[J.npad_ion + j], J.ion_rsoa[2 * J.npad_ion + j], j, 0, r, dx, dy, dz);
        const RT u = functor_eval(J.F1[J.ion_grp[j]], r, du, d2u);
        acc[5] += u;
        acc[6] += du * dx;
        acc[7] += du * dy;
        acc[8] += du * dz;
        acc[9] += d2u + RT(2) * du;
      }
    }
  }
  block_sum<RT, 10>(acc, red);
  if (J.has_j2 && tid == 0)
  {
    RT* vgl = J.j2_vgl + (size_t)iw * 5;
    vgl[0]  = acc[0];
    vgl[1]  = acc[1];
    vgl[2]  = acc[2];
    vgl[3]  = acc[3];
    vgl[4]  = -acc[4];
  }
  if (J.has_j1 && tid < 5)
    J.j1_cur[(size_t)iw * 5 + tid] = acc[5 + tid];
}

// ---- the same sums by `nwarps` warps of a CTA WITHOUT any barrier (walker-segment kernel, segment.cuh: the spline
// consumer warps run them while the first stencil slabs are in flight).  Warp `wg` of the set takes the candidates
// it * 32 nwarps + 32 wg + lane, compacts its in-range pairs into its own list segment and leaves its ten partial sums
// {J2: u, gx, gy, gz, lap; J1: the same} in part[0..10) (shared memory, 16 entries per warp); whoever consumes the sums
// adds the warps' partials in index order.  jl: >= ceil((N + nions) / (32 nwarps)) * 32 nwarps entries.
#ifndef QMCB_JB
#define QMCB_JB 1 // position loads of JB iterations in flight at once (B200, NiO-a64, 512 walkers, segment kernels per sweep:
                  // 1 -> 32.5 ms, 2 -> 35.4 ms, 3 -> 40.0 ms: the kernel lives on 64 registers per thread)
#endif
// The same work as a resumable object: init(), one pass1(it) per block of 32 nwarps candidates (each a short chain of
// dependent loads: positions -> distance -> cutoff test), finish() = functors over the compacted list + reduction.  The
// walker-segment kernel interleaves the pass1 steps with the stencil slabs of the spline gather (segment.cuh), so that
// the slab stream never waits for the Jastrow sums and vice versa.
template<typename RT>
struct JastrowMove
{
  int cnt, iters, n2, n1, gi, tid, STEP, np;
  unsigned short* seg;
  const RT* rs;
  __device__ __forceinline__ void init(const int wg, const int nwarps, const JastrowDev<RT>& J, const int iw, const int iat,
                                       unsigned short* jl)
  {
    const int lane = threadIdx.x & 31;
    np = J.npad, STEP = 32 * nwarps, tid = 32 * wg + lane;
    rs    = J.rsoa + (size_t)iw * 3 * np;
    n2    = J.has_j2 ? J.N : 0;
    n1    = J.has_j1 ? J.nions : 0;
    gi    = (iat < J.n_up ? 0 : 1) * 2;
    iters = (n2 + n1 + STEP - 1) / STEP;
    seg   = jl + wg * iters * 32;
    cnt   = 0;
  }
  // pass 1, iteration `it`: distance of candidate it * STEP + tid, in-range pairs compacted into the warp's list
  __device__ __forceinline__ void pass1(const JastrowDev<RT>& J, const int iat, const RT pos[3], const int it)
  {
    const int lane = threadIdx.x & 31;
    const int idx  = it * STEP + tid;
    RT px(0), py(0), pz(0);
#ifdef QMCB_FAKE_POS // (timing experiment: what do the position loads of this pass cost?  results are wrong)
    if (idx < n2)
      px = RT(0.013) * idx, py = RT(0.007) * idx, pz = RT(0.003) * idx;
#else
    if (idx < n2)
      px = rs[idx], py = rs[np + idx], pz = rs[2 * np + idx];
#endif
    else if (idx < n2 + n1)
    {
      const int j = idx - n2;
      px = J.ion_rsoa[j], py = J.ion_rsoa[J.npad_ion + j], pz = J.ion_rsoa[2 * J.npad_ion + j];
    }
    bool need = false;
    RT r, dx, dy, dz;
    if (idx < n2)
    {
      const FunctorDev<RT>& F = J.F2[gi + (idx < J.n_up ? 0 : 1)];
      if (QMCB_JAS_FAST && J.cell.ortho)
        need = idx != iat && F.coefs != nullptr && min_image_r2_ortho(J.cell, pos, px, py, pz) < rcut2_inflated(F.rcut);
      else
      {
        min_image(J.cell, pos, px, py, pz, idx, iat, r, dx, dy, dz);
        need = idx != iat && F.coefs != nullptr && r < F.rcut;
      }
    }
    else if (idx < n2 + n1)
    {
      const FunctorDev<RT>& F = J.F1[J.ion_grp[idx - n2]];
      if (QMCB_JAS_FAST && J.cell.ortho)
        need = F.coefs != nullptr && min_image_r2_ortho(J.cell, pos, px, py, pz) < rcut2_inflated(F.rcut);
      else
      {
        min_image(J.cell, pos, px, py, pz, idx - n2, 0, r, dx, dy, dz);
        need = F.coefs != nullptr && r < F.rcut;
      }
    }
    const unsigned mk = __ballot_sync(0xffffffffu, need);
    if (need)
      seg[cnt + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)idx;
    cnt += __popc(mk);
  }
  // pass 2: functors over the warp's list; the ten partial sums go to part[0..10)
  __device__ __forceinline__ void finish(const JastrowDev<RT>& J, const int iat, const RT pos[3], RT* part)
  {
    const int lane = threadIdx.x & 31;
    RT acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e)
      acc[e] = RT(0);
    __syncwarp();
    for (int e = lane; e < cnt; e += 32)
    {
      // (one copy of the distance + functor code for both kinds of partner: the walker-segment kernel is instruction-cache
      // bound, segment.cuh)
      const int idx  = seg[e];
      const bool is2 = idx < n2;
      const int j    = is2 ? idx : idx - n2;
      const RT* ps   = is2 ? rs : J.ion_rsoa;
      const int st   = is2 ? np : J.npad_ion;
      RT r, dx, dy, dz, du, d2u;
      min_image(J.cell, pos, ps[j], ps[st + j], ps[2 * st + j], j, is2 ? iat : 0, r, dx, dy, dz);
      const FunctorDev<RT>& F = is2 ? J.F2[gi + (idx < J.n_up ? 0 : 1)] : J.F1[J.ion_grp[j]];
      const RT u  = functor_eval(F, r, du, d2u);
      const RT c1 = du * dx, c2 = du * dy, c3 = du * dz, c4 = d2u + RT(2) * du;
      if (is2)
        acc[0] += u, acc[1] += c1, acc[2] += c2, acc[3] += c3, acc[4] += c4;
      else
        acc[5] += u, acc[6] += c1, acc[7] += c2, acc[8] += c3, acc[9] += c4;
    }
    warp_fold<RT, 16>(acc); // lane l holds the total of value l >> 1
    if ((lane & 1) == 0 && (lane >> 1) < 10)
      part[lane >> 1] = acc[0];
  }
};

template<typename RT>
__device__ __forceinline__ void jastrow_move_warps(const int wg, const int nwarps, const JastrowDev<RT>& J, const int iw,
                                                   const int iat, const RT pos[3], unsigned short* jl, RT* part)
{
  JastrowMove<RT> jm;
  jm.init(wg, nwarps, J, iw, iat, jl);
  for (int it = 0; it < jm.iters; ++it)
    jm.pass1(J, iat, pos, it);
  jm.finish(J, iat, pos, part);
}

// ---- ratios at VIRTUAL positions (non-local pseudopotential quadrature points): TwoBodyJastrow::mw_evaluateRatios
// (Jastrow/TwoBodyJastrow.cpp:174-210) -> BsplineFunctor::mw_evaluateV (Jastrow/BsplineFunctor.cpp:135-200), and
// J1OrbitalSoA::evaluateRatios (Jastrow/J1OrbitalSoA.h:295-299).  One warp per virtual position ivp of walker wk[ivp]
// replacing electron ref[ivp]:  out[ivp] = exp(Uat[ref] - sum_{j != ref} u(|r_j - r_vp|)) * exp(Vat[ref] - sum_I u(|R_I - r_vp|)).
// grid = ceil(nvp / 4), block = 128.
template<typename RT>
__global__ void __launch_bounds__(128)
    jastrow_vp_ratio_kernel(const JastrowDev<RT> J, const int nvp, const int* wk, const int* ref, const RT* rvp, double* out)
{
  const int lane = threadIdx.x & 31, ivp = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (ivp >= nvp)
    return;
  const int iw = wk[ivp], iat = ref[ivp], np = J.npad;
  const RT pos[3] = {rvp[3 * ivp], rvp[3 * ivp + 1], rvp[3 * ivp + 2]};
  const RT* rs    = J.rsoa + (size_t)iw * 3 * np;
  RT u2(0), u1(0);
  if (J.has_j2)
  {
    const int gi = (iat < J.n_up ? 0 : 1) * 2;
    for (int j = lane; j < J.N; j += 32)
    {
      if (j == iat)
        continue;
      RT r, dx, dy, dz, du, d2u;
      min_image(J.cell, pos, rs[j], rs[np + j], rs[2 * np + j], j, 0, r, dx, dy, dz);
      u2 += functor_eval(J.F2[gi + (j < J.n_up ? 0 : 1)], r, du, d2u);
    }
  }
  if (J.has_j1)
    for (int j = lane; j < J.nions; j += 32)
    {
      RT r, dx, dy, dz, du, d2u;
      min_image(J.cell, pos, J.ion_rsoa[j], J.ion_rsoa[J.npad_ion + j], J.ion_rsoa[2 * J.npad_ion + j], j, 0, r, dx, dy, dz);
      u1 += functor_eval(J.F1[J.ion_grp[j]], r, du, d2u);
    }
  u2 = warp_sum(u2);
  u1 = warp_sum(u1);
  if (lane == 0)
  {
    double ratio = 1.0;
    if (J.has_j2)
      ratio *= exp((double)(J.Uat[(size_t)iw * np + iat] - u2));
    if (J.has_j1)
      ratio *= exp((double)(J.Vat[(size_t)iw * J.N + iat] - u1));
    out[ivp] = ratio;
  }
}

// grid = nw
template<typename RT, bool STORE>
__global__ void __launch_bounds__(JAS_TPB) jastrow_move_kernel(const JastrowDev<RT> J, const int iat)
{
  __shared__ RT red[10 * 32];
  __shared__ unsigned short jl[JAS_LIST];
  const int iw    = blockIdx.x;
  const RT pos[3] = {J.newpos[3 * iw], J.newpos[3 * iw + 1], J.newpos[3 * iw + 2]};
  jastrow_move_body<RT, STORE>(J, iw, iat, pos, red, jl);
}

// one pair of the accept: both distance rows and both functor evaluations are recomputed from the positions (see
// jastrow_move_body); BsplineFunctor.cpp:262-324
template<typename RT>
struct J2Pair
{
  RT px, py, pz, ua, da, db, dc, l2;
};

// ---- accept (walker iw, thread group g): J2 per-particle sums (BsplineFunctor.cpp:262-324), J1 commit, position commit
// jl: JAS_LIST entries of shared memory for the group
template<typename RT>
__device__ __forceinline__ void jastrow_accept_body(const Group& g, const JastrowDev<RT>& J, const int iw, const int iat,
                                                    unsigned short* jl, long long* tr = nullptr /* debug trace (segment.cuh) */)
{
#define JAS_TR(id)             \
  do                           \
  {                            \
    if (tr && g.tid == 0)      \
      tr[id] = clock64();      \
  } while (0)
  JAS_TR(36);
  const int tid = g.tid, N = J.N, np = J.npad;
  const int lane = tid & 31, wg = tid >> 5, nwg = g.n >> 5;
  if (J.has_j2)
  {
    const RT* rs   = J.rsoa + (size_t)iw * 3 * np;
    const RT pnew[3] = {J.newpos[3 * iw], J.newpos[3 * iw + 1], J.newpos[3 * iw + 2]};
    const RT pold[3] = {rs[iat], rs[np + iat], rs[2 * np + iat]}; // (the commit below waits behind the group barrier)
    RT* Uat        = J.Uat + (size_t)iw * np;
    RT* dU         = J.dUat + (size_t)iw * 3 * np;
    RT* d2U        = J.d2Uat + (size_t)iw * np;
    const RT* vgl  = J.j2_vgl + (size_t)iw * 5;
    const int gi   = (iat < J.n_up ? 0 : 1) * 2;
    const RT Uold_iat = Uat[iat];
    g.sync();
    auto update = [&](const int j, const J2Pair<RT>& q) {
      RT rn, nx, ny, nz, ro, ox, oy, oz, cdu, cd2, du, d2u;
      min_image(J.cell, pnew, q.px, q.py, q.pz, j, iat, rn, nx, ny, nz);
      min_image(J.cell, pold, q.px, q.py, q.pz, j, iat, ro, ox, oy, oz);
      const FunctorDev<RT>& F = J.F2[gi + (j < J.n_up ? 0 : 1)];
      const RT cu = functor_eval(F, rn, cdu, cd2);
      const RT u  = functor_eval(F, ro, du, d2u);
      Uat[j]         = q.ua + (cu - u);
      dU[j]          = q.da - (nx * cdu - ox * du);
      dU[np + j]     = q.db - (ny * cdu - oy * du);
      dU[2 * np + j] = q.dc - (nz * cdu - oz * du);
      d2U[j]         = q.l2 - (cd2 + RT(2) * cdu - (d2u + RT(2) * du));
    };
    auto fetch = [&](const int j, J2Pair<RT>& q) {
      q.px = rs[j];
      q.py = rs[np + j];
      q.pz = rs[2 * np + j];
      q.ua = Uat[j];
      q.da = dU[j];
      q.db = dU[np + j];
      q.dc = dU[2 * np + j];
      q.l2 = d2U[j];
    };
    const int iters  = (N + g.n - 1) / g.n;
    const int segcap = iters * 32;
    const bool listed = segcap * nwg <= JAS_LIST && N < 65536;
    // two elements per thread at a time: every load of a chunk is issued before the first store, so the HBM round trips
    // overlap instead of queueing behind the read-modify-write stores (which the compiler must assume may alias)
#ifndef QMCB_JACC_CH
#define QMCB_JACC_CH 1 // pairs per thread per trip in the accept (B200, a64: 1 -> 31.1 ms, 2 -> 33.3 ms per sweep)
#endif
    constexpr int CH = QMCB_JACC_CH;
    if (listed)
    {
      // pass 1: pairs whose old OR new distance is inside the cutoff (the others add exact zeros to every sum)
      unsigned short* seg = jl + wg * segcap;
      int cnt = 0;
      // (the position loads of up to three iterations are requested before the first is used)
      constexpr int JB = QMCB_JB;
      for (int it0 = 0; it0 < iters; it0 += JB)
      {
        RT px[JB], py[JB], pz[JB];
#pragma unroll
        for (int h = 0; h < JB; ++h)
        {
          const int j = (it0 + h) * g.n + tid;
          px[h] = py[h] = pz[h] = RT(0);
          if (it0 + h < iters && j < N)
            px[h] = rs[j], py[h] = rs[np + j], pz[h] = rs[2 * np + j];
        }
#pragma unroll
        for (int h = 0; h < JB; ++h)
        {
          if (it0 + h >= iters)
            break;
          const int j = (it0 + h) * g.n + tid;
          bool need   = false;
          if (j < N && j != iat)
          {
            const FunctorDev<RT>& F = J.F2[gi + (j < J.n_up ? 0 : 1)];
            if (QMCB_JAS_FAST && J.cell.ortho)
            {
              const RT thr = rcut2_inflated(F.rcut);
              need = F.coefs != nullptr && (min_image_r2_ortho(J.cell, pnew, px[h], py[h], pz[h]) < thr ||
                                            min_image_r2_ortho(J.cell, pold, px[h], py[h], pz[h]) < thr);
            }
            else
            {
              RT rn, ro, t0, t1, t2;
              min_image(J.cell, pnew, px[h], py[h], pz[h], j, iat, rn, t0, t1, t2);
              min_image(J.cell, pold, px[h], py[h], pz[h], j, iat, ro, t0, t1, t2);
              need = F.coefs != nullptr && (rn < F.rcut || ro < F.rcut);
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, need);
          if (need)
            seg[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
          cnt += __popc(m);
        }
      }
      __syncwarp();
      JAS_TR(37);
      // pass 2: full warps over the list
      for (int e0 = lane; e0 < cnt; e0 += CH * 32)
      {
        J2Pair<RT> q[CH];
        int jj[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
        {
          const int e = e0 + c * 32;
          jj[c]       = e < cnt ? (int)seg[e] : -1;
          if (jj[c] >= 0)
            fetch(jj[c], q[c]);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (jj[c] >= 0)
            update(jj[c], q[c]);
      }
    }
    else
    {
      for (int j0 = tid; j0 < N; j0 += CH * g.n)
      {
        J2Pair<RT> q[CH];
        bool on[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
        {
          const int j = j0 + c * g.n;
          on[c]       = j < N && j != iat;
          if (on[c])
            fetch(j, q[c]);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c)
          if (on[c])
            update(j0 + c * g.n, q[c]);
      }
    }
    JAS_TR(38);
    if (tid == 0)
    {
      J.j2_log[iw] += (double)(Uold_iat - vgl[0]); // TwoBodyJastrow.cpp:656-659
      Uat[iat]         = vgl[0];
      dU[iat]          = vgl[1];
      dU[np + iat]     = vgl[2];
      dU[2 * np + iat] = vgl[3];
      d2U[iat]         = vgl[4];
    }
  }
  // (on another warp than the J2 commit above: each is a short chain of dependent global read-modify-writes)
  if (J.has_j1 && tid == (g.n > 32 ? 32 : 0))
  {
    const RT* cur = J.j1_cur + (size_t)iw * 5;
    RT* Vat       = J.Vat + (size_t)iw * N;
    J.j1_log[iw] += (double)(Vat[iat] - cur[0]); // J1OrbitalSoA.h:463
    Vat[iat]                                = cur[0];
    J.Grad1[((size_t)iw * 3 + 0) * N + iat] = cur[1];
    J.Grad1[((size_t)iw * 3 + 1) * N + iat] = cur[2];
    J.Grad1[((size_t)iw * 3 + 2) * N + iat] = cur[3];
    J.Lap1[(size_t)iw * N + iat]            = cur[4];
  }
  // ParticleSet::mw_accept_rejectMove (ParticleSet.cpp:717-758): commit the position
  if (tid < 3)
    J.rsoa[(size_t)iw * 3 * np + tid * np + iat] = J.newpos[3 * iw + tid];
  JAS_TR(39);
#undef JAS_TR
}

template<typename RT>
__global__ void __launch_bounds__(JAS_TPB)
    jastrow_accept_kernel(const JastrowDev<RT> J, const int iat, const unsigned char* accepted)
{
  __shared__ unsigned short jl[JAS_LIST];
  if (!accepted[blockIdx.x])
    return;
  jastrow_accept_body<RT>(cta_group(), J, blockIdx.x, iat, jl);
}

// ---- from scratch (TwoBodyJastrow.cpp:667-713 lower-triangle form; J1OrbitalSoA.h:237-250).  grid = nw.
// The per-particle sums are accumulated row by row exactly like the reference (row iat touches j < iat).
template<typename RT>
__global__ void __launch_bounds__(JAS_TPB) jastrow_recompute_kernel(const JastrowDev<RT> J)
{
  __shared__ RT red[5 * 32];
  const int iw = blockIdx.x, tid = threadIdx.x, N = J.N, np = J.npad;
  const RT* rs = J.rsoa + (size_t)iw * 3 * np;
  if (J.has_j2)
  {
    RT* Uat = J.Uat + (size_t)iw * np;
    RT* dU  = J.dUat + (size_t)iw * 3 * np;
    RT* d2U = J.d2Uat + (size_t)iw * np;
    for (int j = tid; j < np; j += JAS_TPB)
    {
      Uat[j] = RT(0);
      d2U[j] = RT(0);
      dU[j] = dU[np + j] = dU[2 * np + j] = RT(0);
    }
    __syncthreads();
    for (int iat = 0; iat < N; ++iat)
    {
      const RT pos[3] = {rs[iat], rs[np + iat], rs[2 * np + iat]};
      const int gi    = (iat < J.n_up ? 0 : 1) * 2;
      RT acc[5]       = {RT(0), RT(0), RT(0), RT(0), RT(0)};
      for (int j = tid; j < iat; j += JAS_TPB)
      {
        RT r, dx, dy, dz, du, d2u;
        min_image(J.cell, pos, rs[j], rs[np + j], rs[2 * np + j], j, iat, r, dx, dy, dz);
        const RT u = functor_eval(J.F2[gi + (j < J.n_up ? 0 : 1)], r, du, d2u);
        acc[0] += u;
        acc[1] += du * dx;
        acc[2] += du * dy;
        acc[3] += du * dz;
        acc[4] += d2u + RT(2) * du;
        Uat[j] += u;
        d2U[j] -= d2u + RT(2) * du;
        dU[j] -= du * dx;
        dU[np + j] -= du * dy;
        dU[2 * np + j] -= du * dz;
      }
      block_sum<RT, 5>(acc, red);
      if (tid == 0)
      {
        Uat[iat]         = acc[0];
        dU[iat]          = acc[1];
        dU[np + iat]     = acc[2];
        dU[2 * np + iat] = acc[3];
        d2U[iat]         = -acc[4];
      }
      __syncthreads();
    }
    // log_value = -0.5 * sum Uat (TwoBodyJastrow.cpp:769-781)
    double lv[1] = {0.0};
    for (int j = tid; j < N; j += JAS_TPB)
      lv[0] -= 0.5 * (double)Uat[j];
    __shared__ double dred[32];
    block_sum<double, 1>(lv, dred);
    if (tid == 0)
      J.j2_log[iw] = lv[0];
  }
  if (J.has_j1)
  {
    double lv = 0.0;
    for (int iat = 0; iat < N; ++iat)
    {
      const RT pos[3] = {rs[iat], rs[np + iat], rs[2 * np + iat]};
      RT o[5];
      j1_sums(J, pos, o, red);
      if (tid == 0)
      {
        J.Vat[(size_t)iw * N + iat]             = o[0];
        J.Grad1[((size_t)iw * 3 + 0) * N + iat] = o[1];
        J.Grad1[((size_t)iw * 3 + 1) * N + iat] = o[2];
        J.Grad1[((size_t)iw * 3 + 2) * N + iat] = o[3];
        J.Lap1[(size_t)iw * N + iat]            = o[4];
        lv -= (double)o[0];
      }
    }
    if (tid == 0)
      J.j1_log[iw] = lv;
  }
}

// G += dUat, L += d2Uat (J2);  G += Grad, L -= Lap (J1).   grid = (ceil(N/256), nw)
template<typename RT>
__global__ void jastrow_add_gl_kernel(const JastrowDev<RT> J, RT* Gd, RT* Ld, const int cs /* 1 real, 2 complex (re,im) */)
{
  const int iw = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x, N = J.N, np = J.npad;
  if (i >= N)
    return;
  RT g[3] = {RT(0), RT(0), RT(0)}, l(0);
  if (J.has_j2)
  {
    const RT* dU = J.dUat + (size_t)iw * 3 * np;
    g[0] += dU[i];
    g[1] += dU[np + i];
    g[2] += dU[2 * np + i];
    l += J.d2Uat[(size_t)iw * np + i];
  }
  if (J.has_j1)
  {
    g[0] += J.Grad1[((size_t)iw * 3 + 0) * N + i];
    g[1] += J.Grad1[((size_t)iw * 3 + 1) * N + i];
    g[2] += J.Grad1[((size_t)iw * 3 + 2) * N + i];
    l -= J.Lap1[(size_t)iw * N + i];
  }
  // the Jastrows are real: with complex G, L (cs == 2) only the real parts receive a contribution
  RT* go = Gd + ((size_t)iw * N + i) * 3 * cs;
  go[0] += g[0];
  go[cs] += g[1];
  go[2 * cs] += g[2];
  Ld[((size_t)iw * N + i) * cs] += l;
}
#endif

} // namespace qmcb
