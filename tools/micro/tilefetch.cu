// Traceback tile fetch: per-lane cp.async (16-byte pieces) vs per-lane TMA bulk copies (cp.async.bulk, one 128-byte
// row per lane, mbarrier completion).  Same geometry as poa_trace.cuh: 32 rows x 128 B out of 1 KB matrix rows, one
// warp per CTA, 16 CTAs per SM, random rows of a 4 GB buffer (DRAM-resident).  nvcc -arch=sm_100a -O3 tilefetch.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kPitch = 144;  // bytes between tile rows in shared memory
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>  // 0: cp.async 16 B x 8 per lane, 1: cp.async.bulk 128 B per lane
__global__ void __launch_bounds__(32, 16) fetch(const uint8_t* buf, uint64_t rows, int iters, uint32_t* out, long long* cyc) {
  __shared__ __align__(128) uint8_t tile[32 * kPitch];
  __shared__ __align__(8) uint64_t mbar;
  const int lane = threadIdx.x;
  uint32_t acc = 0, parity = 0;
  uint64_t x = blockIdx.x * 0x9E3779B97F4A7C15ull + 12345;
  if (MODE == 1 && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    const uint64_t r = (x >> 20) % (rows - 32);
    const uint32_t col = static_cast<uint32_t>((x >> 8) & 7u) * 128u;  // 128-byte column window inside the 1 KB row
    const uint8_t* src = buf + (r + lane) * 1024ull + col;
    const uint32_t dst = smem_u32(tile) + lane * kPitch;
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * 16), "l"(src + q * 16) : "memory");
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
    } else {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(32 * 128) : "memory");
      __syncwarp();
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(dst), "l"(src),
                   "r"(smem_u32(&mbar))
                   : "memory");
      uint32_t done = 0;
      while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
      }
      parity ^= 1u;
    }
    acc += *reinterpret_cast<uint32_t*>(tile + ((x >> 40) & 31u) * kPitch + lane * 4);
    __syncwarp();
  }
  const long long t1 = clock64();
  out[blockIdx.x * 32 + lane] = acc;
  if (lane == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  const uint64_t rows = 4ull << 20;  // 4 M rows x 1 KB = 4 GB
  uint8_t* buf;
  if (cudaMalloc(&buf, rows * 1024) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 1, rows * 1024);
  const int iters = 2000;
  for (int per_sm : {1, 4, 16}) {
    const int blocks = 148 * per_sm;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, blocks * 32 * 4); cudaMalloc(&cyc, blocks * 8);
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) fetch<0><<<blocks, 32>>>(buf, rows, iters, out, cyc);
        else fetch<1><<<blocks, 32>>>(buf, rows, iters, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long* h = new long long[blocks];
      cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < blocks; ++i) s += h[i];
      printf("%-28s warps/SM %2d: %.0f cycles per tile fetch (32 rows x 128 B)\n", mode == 0 ? "cp.async 16 B x 8 per lane" : "cp.async.bulk 128 B per lane", per_sm, s / blocks / iters);
      delete[] h;
    }
  }
  return 0;
}
