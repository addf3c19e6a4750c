// Dependent random-load latency vs footprint (TLB reach probe).  nvcc -arch=sm_100a -O3 randlat.cu -o randlat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void chase(const uint32_t* buf, uint64_t nwords, int iters, uint32_t* out, long long* cyc) {
  uint64_t x = (blockIdx.x * 2654435761u + threadIdx.x * 40503u) % nwords;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v = buf[x];
    acc += v;
    x = ((x * 6364136223846793005ull + 1442695040888963407ull + v) >> 11) % nwords;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = acc; cyc[blockIdx.x] = t1 - t0; }
}
int main() {
  const int iters = 400;
  for (double gb : {0.125, 0.5, 2.0, 8.0, 32.0, 96.0}) {
    uint64_t bytes = (uint64_t)(gb * (1ull << 30));
    uint32_t* buf; if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc %.1f GB failed\n", gb); continue; }
    cudaMemset(buf, 0, bytes);
    for (int blocks : {148, 2000, 8000}) {
      for (int threads : {1, 8}) {
        uint32_t* out; long long* cyc; cudaMalloc(&out, blocks * 4); cudaMalloc(&cyc, blocks * 8);
        chase<<<blocks, threads>>>(buf, bytes / 4, iters, out, cyc);
        cudaDeviceSynchronize();
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a); chase<<<blocks, threads>>>(buf, bytes / 4, iters, out, cyc); cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        long long* h = new long long[blocks]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < blocks; ++i) s += h[i];
        printf("footprint %6.2f GB  warps %5d lanes %d : %7.0f cycles/load (kernel %.3f ms)\n", gb, blocks, threads, s / blocks / iters, ms);
        delete[] h; cudaFree(out); cudaFree(cyc);
      }
    }
    cudaFree(buf);
  }
  return 0;
}
