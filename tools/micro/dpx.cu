// DPX (VIADDMNMX.S16x2) and SHFL latency / throughput probe.  nvcc -arch=sm_100a -O3 dpx.cu -o dpx
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dpx_tp(uint32_t* out, int iters, long long* cyc) {
  uint32_t a[ILP], b = threadIdx.x * 3 + 1, c = blockIdx.x + 7;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = __viaddmax_s16x2(a[i], b, c);
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void iadd_tp(uint32_t* out, int iters, long long* cyc) {
  uint32_t a[ILP], b = threadIdx.x * 3 + 1;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = __vmaxs2(a[i], b) ;
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_lat(uint32_t* out, int iters, long long* cyc) {
  uint32_t v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v = __vmaxs2(v, __shfl_up_sync(0xFFFFFFFFu, v, 1)) + 1;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void shfl_tp(uint32_t* out, int iters, long long* cyc) {
  uint32_t a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = __shfl_up_sync(0xFFFFFFFFu, a[i], 1);
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <class F>
void run(const char* name, F f, int ilp, int iters) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {1, 4, 8, 16, 32}) {
    f<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    f<<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / iters / ilp;
    printf("%-10s ilp %d warps/SM %2d: %.2f cyc per op per warp; SM throughput %.2f warp-ops/cyc\n", name, ilp, warps, per, warps / per);
  }
}
int main() {
  const int iters = 20000;
  run("viaddmax", dpx_tp<1>, 1, iters);
  run("viaddmax", dpx_tp<8>, 8, iters);
  run("vmaxs2", iadd_tp<1>, 1, iters);
  run("vmaxs2", iadd_tp<8>, 8, iters);
  run("shfl+max", shfl_lat, 1, iters);
  run("shfl", shfl_tp<8>, 8, iters);
  return 0;
}
