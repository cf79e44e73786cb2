// qmcpack_b200/csrc/common.cuh -- shared device/host helpers of libqmcb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <atomic>

namespace qmcb
{
extern std::atomic<unsigned long long> g_launch_count;

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

#ifdef __CUDACC__
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
