// Microbenchmark: issue rate of legacy mma.sync TF32 (m16n8k8), FP64 DMMA (m8n8k4) and plain FFMA/DFMA on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template<int NACC>
__global__ void k_tf32(float* out, int iters)
{
  float d[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  float s = 0; for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int NACC>
__global__ void k_f64(double* out, int iters)
{
  double d[NACC][2];
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = 0.0;
  double a = 1.0, b = 1.0;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<typename T, int NACC>
__global__ void k_fma(T* out, int iters, T x)
{
  T d[NACC];
  for (int i = 0; i < NACC; ++i) d[i] = T(i);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) d[i] = d[i] * x + x;
  T s = 0; for (int i = 0; i < NACC; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<typename F> float timeit(F f)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* buf; cudaMalloc(&buf, 1 << 26);
  const int iters = 4096, tpb = 256, blocks = sms * 4;
  const double warps = (double)blocks * tpb / 32;
  { float ms = timeit([&] { k_tf32<8><<<blocks, tpb>>>((float*)buf, iters); });
    double fl = warps * iters * 8 * 2.0 * 16 * 8 * 8; printf("mma.sync m16n8k8 tf32 : %.1f TFLOP/s\n", fl / ms / 1e9); }
  { float ms = timeit([&] { k_f64<8><<<blocks, tpb>>>((double*)buf, iters); });
    double fl = warps * iters * 8 * 2.0 * 8 * 8 * 4; printf("mma.sync m8n8k4 f64   : %.1f TFLOP/s\n", fl / ms / 1e9); }
  { float ms = timeit([&] { k_fma<float, 16><<<blocks, tpb>>>((float*)buf, iters, 1.0001f); });
    double fl = warps * 32 * iters * 16 * 2.0; printf("FFMA                   : %.1f TFLOP/s\n", fl / ms / 1e9); }
  { float ms = timeit([&] { k_fma<double, 16><<<blocks, tpb>>>((double*)buf, iters, 1.0001); });
    double fl = warps * 32 * iters * 16 * 2.0; printf("DFMA                   : %.1f TFLOP/s\n", fl / ms / 1e9); }
  return 0;
}
