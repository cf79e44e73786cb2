// qmcpack_b200/csrc/common.cuh -- shared device/host helpers of libqmcb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

namespace qmcb
{
extern std::atomic<unsigned long long> g_launch_count;
// programmatic dependent launch between the kernels of a sweep (env QMCB_PDL): 0 = off, 1 = boundary -> spline gather,
// 2 = also spline gather -> boundary; +4 = the primary also signals at its START that dependents may become resident
// (otherwise they follow its CTAs out as those exit).  The dependent kernel's CTAs run their prologue while the previous
// kernel drains; every global access of a dependent sits behind pdl_wait().
extern int g_pdl_mode;

struct CudaError : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

#define QMCB_CUDA(call)                                                                                     \
  do                                                                                                        \
  {                                                                                                         \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess)                                                                                 \
      throw qmcb::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " at " + __FILE__ + \
                            ":" + std::to_string(__LINE__));                                                \
  } while (0)

#define QMCB_LAUNCH_CHECK()                                \
  do                                                       \
  {                                                        \
    qmcb::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    QMCB_CUDA(cudaPeekAtLastError());                      \
  } while (0)

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is per DEVICE (and per kernel), and crowds of one
// process may live on several GPUs and call in from concurrent host threads: the record is keyed by (kernel, device)
// and guarded by a mutex -- a process-wide "done once" flag would leave every launch above 48 KB failing on the second GPU.
template<typename K>
inline void ensure_dynamic_smem(K kern, size_t bytes)
{
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  QMCB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = granted[std::make_pair(reinterpret_cast<const void*>(kern), dev)];
  if (have < bytes)
  {
    QMCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
}

// getAlignedSize<T> of the reference (Platforms/CPU/SIMD/aligned_allocator.hpp:41-47): 64-byte rows
template<typename T>
inline size_t aligned_size(size_t n)
{
  constexpr size_t ND = 64 / sizeof(T);
  return ((n + ND - 1) / ND) * ND;
}

// simple owning device buffer
template<typename T>
struct DevBuf
{
  T* p     = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count, bool zero = true)
  {
    release();
    n = count;
    if (count == 0)
      return;
    QMCB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    if (zero)
    {
      // cudaMemset runs on the legacy default stream and is asynchronous to the host; the crowd streams are
      // non-blocking, so without this wait a later kernel or copy on them could be overtaken by the memset
      QMCB_CUDA(cudaMemset(p, 0, count * sizeof(T)));
      QMCB_CUDA(cudaDeviceSynchronize());
    }
  }
  size_t bytes() const { return n * sizeof(T); }
};

// pinned host staging buffer
template<typename T>
struct PinBuf
{
  T* p     = nullptr;
  size_t n = 0;
  PinBuf() = default;
  PinBuf(const PinBuf&) = delete;
  PinBuf& operator=(const PinBuf&) = delete;
  ~PinBuf()
  {
    if (p)
      cudaFreeHost(p);
  }
  void alloc(size_t count)
  {
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    n = count;
    if (count)
      QMCB_CUDA(cudaMallocHost(&p, count * sizeof(T)));
  }
};

// ---- complex value type of the C2C (complex orbital) determinants: plain {re, im} pair, layout-compatible with
// std::complex / the interleaved arrays of the C ABI; trivially constructible so that it can live in __shared__ arrays
#ifdef __CUDACC__
#define QMCB_HD __host__ __device__ __forceinline__
#else
#define QMCB_HD inline
#endif
template<typename T>
struct cx
{
  T re, im;
  cx() = default;
  QMCB_HD cx(T r) : re(r), im(T(0)) {}
  QMCB_HD cx(T r, T i) : re(r), im(i) {}
  template<typename U>
  QMCB_HD explicit cx(const cx<U>& o) : re((T)o.re), im((T)o.im)
  {}
  QMCB_HD cx& operator+=(const cx& o)
  {
    re += o.re;
    im += o.im;
    return *this;
  }
  QMCB_HD cx& operator-=(const cx& o)
  {
    re -= o.re;
    im -= o.im;
    return *this;
  }
};
template<typename T>
QMCB_HD cx<T> operator+(const cx<T>& a, const cx<T>& b)
{
  return cx<T>(a.re + b.re, a.im + b.im);
}
template<typename T>
QMCB_HD cx<T> operator-(const cx<T>& a, const cx<T>& b)
{
  return cx<T>(a.re - b.re, a.im - b.im);
}
template<typename T>
QMCB_HD cx<T> operator-(const cx<T>& a)
{
  return cx<T>(-a.re, -a.im);
}
template<typename T>
QMCB_HD cx<T> operator*(const cx<T>& a, const cx<T>& b)
{
  return cx<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template<typename T>
QMCB_HD cx<T> operator*(const cx<T>& a, const T b)
{
  return cx<T>(a.re * b, a.im * b);
}
template<typename T>
QMCB_HD cx<T> operator/(const cx<T>& a, const cx<T>& b)
{
  // textbook quotient (a * conj(b)) / |b|^2
  const T d = b.re * b.re + b.im * b.im;
  return cx<T>((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
template<typename T>
QMCB_HD bool operator!=(const cx<T>& a, const cx<T>& b)
{
  return a.re != b.re || a.im != b.im;
}
template<typename T>
QMCB_HD bool operator==(const cx<T>& a, const cx<T>& b)
{
  return a.re == b.re && a.im == b.im;
}

// scalar type, double-precision counterpart and a few accessors that read the same for real and complex values
template<typename V>
struct value_traits
{
  using real_t                     = V;
  using dbl_t                      = double;
  static constexpr bool is_complex = false;
  static constexpr int ncomp       = 1;
};
template<typename T>
struct value_traits<cx<T>>
{
  using real_t                     = T;
  using dbl_t                      = cx<double>;
  static constexpr bool is_complex = true;
  static constexpr int ncomp       = 2;
};
QMCB_HD float real_part(float v) { return v; }
QMCB_HD double real_part(double v) { return v; }
template<typename T>
QMCB_HD T real_part(const cx<T>& v)
{
  return v.re;
}
QMCB_HD double to_dbl(float v) { return (double)v; }
QMCB_HD double to_dbl(double v) { return v; }
template<typename T>
QMCB_HD cx<double> to_dbl(const cx<T>& v)
{
  return cx<double>((double)v.re, (double)v.im);
}
template<typename V>
QMCB_HD V from_dbl(double v)
{
  return (V)v;
}
template<typename V>
QMCB_HD V from_dbl(const cx<double>& v)
{
  return V(v);
}
// squared modulus (std::norm)
QMCB_HD double norm2(double v) { return v * v; }
QMCB_HD double norm2(const cx<double>& v) { return v.re * v.re + v.im * v.im; }

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// kernel launch with (optionally) the programmatic-stream-serialization attribute
template<typename... KArgs, typename... Args>
inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                          Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = grid;
  cfg.blockDim           = block;
  cfg.dynamicSmemBytes   = smem;
  cfg.stream             = st;
  cudaLaunchAttribute attr[1];
  if (pdl)
  {
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = 1;
  }
  QMCB_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}

// streaming (read-once / write-once) global accesses: ld.global.cs / st.global.cs mark the line evict-first in L2, so
// one-shot data (the proposed move's orbital rows, the gradient rows written by an accept) does not push the walkers'
// recurring working set (delayed-update rows, Jastrow sums, positions) out of L2
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ cx<float> ld_stream(const cx<float>* p)
{
  const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
  return cx<float>(v.x, v.y);
}
__device__ __forceinline__ cx<double> ld_stream(const cx<double>* p)
{
  const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
  return cx<double>(v.x, v.y);
}
__device__ __forceinline__ void st_stream(float* p, const float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double* p, const double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(cx<float>* p, const cx<float>& v)
{
  __stcs(reinterpret_cast<float2*>(p), make_float2(v.re, v.im));
}
__device__ __forceinline__ void st_stream(cx<double>* p, const cx<double>& v)
{
  __stcs(reinterpret_cast<double2*>(p), make_double2(v.re, v.im));
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template<typename T>
__device__ __forceinline__ cx<T> warp_sum(cx<T> v)
{
  return cx<T>(warp_sum(v.re), warp_sum(v.im));
}

// warp-wide sums of NV (4 or 16) values per lane with the recursive-halving exchange: at every step a lane hands half of
// its values to its partner and keeps the other half, so NV values cost NV - 1 + log2(32 / NV) shuffles instead of
// 5 NV.  On return lane l holds, in v[0], the total of value number l / (32 / NV) (l >> 1 for NV = 16, l >> 3 for NV = 4);
// the order in which a total is added up is fixed (it depends on the lane numbers only).
template<typename T, int NV>
__device__ __forceinline__ void warp_fold(T (&v)[NV])
{
  static_assert(NV == 4 || NV == 16, "warp_fold handles 4 or 16 values");
  const int lane = threadIdx.x & 31;
  int off        = 16;
#pragma unroll
  for (int m = NV / 2; m >= 1; m /= 2, off /= 2)
  {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i)
    {
      const T send = upper ? v[i] : v[i + m];
      const T keep = upper ? v[i + m] : v[i];
      v[i]         = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off >= 1; off /= 2)
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}

// a group of threads of one CTA that synchronise on their own named barrier (bar 0 with the full CTA == __syncthreads)
struct Group
{
  int tid, n, bar;
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory"); }
};
__device__ __forceinline__ Group cta_group() { return Group{(int)threadIdx.x, (int)blockDim.x, 0}; }

// group-wide sum of NV values per thread, result valid in every thread of the group; `red` >= NV*32 T of shared memory;
// g.n must be a multiple of 32 and every thread of the group must call it
template<typename T, int NV>
__device__ __forceinline__ void group_sum(const Group& g, T (&v)[NV], T* red)
{
  const int lane = g.tid & 31, warp = g.tid >> 5, nwarp = g.n >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    v[i] = warp_sum(v[i]);
  g.sync();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i)
      red[i * 32 + warp] = v[i];
  g.sync();
#pragma unroll
  for (int i = 0; i < NV; ++i)
  {
    T x  = lane < nwarp ? red[i * 32 + lane] : T(0);
    v[i] = warp_sum(x);
  }
}

// block-wide sum of NV values per thread, result valid in every thread.  `red` is shared scratch of >= NV*32 T.
// All threads of the block must call it (uses __syncthreads).
template<typename T, int NV>
__device__ __forceinline__ void block_sum(T (&v)[NV], T* red)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    v[i] = warp_sum(v[i]);
  __syncthreads(); // protect `red` from a previous use
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i)
      red[i * 32 + warp] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i)
  {
    T x  = lane < nwarp ? red[i * 32 + lane] : T(0);
    v[i] = warp_sum(x);
  }
}
#endif

} // namespace qmcb
