// Micro-benchmark: cost of back-to-back tcgen05.mma instructions of the small shapes mlp_tc issues (sm_100a).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I reinforcement-learning-replications_b200/csrc
//             -I include -o tools/bin/tc_mma_bench tools/tc_mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace b200rl;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void umma_f16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}

template <int M, int N, int NACC, bool TS, bool AMN, bool BMN, int NMMA>
__global__ void __launch_bounds__(128, 1) bench(long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t holder;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&holder), 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&mbar), 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t bar = smem_u32(&mbar);
  constexpr uint32_t idesc = make_idesc_bf16(M, N, AMN, BMN);
  if (threadIdx.x < 32) {
    const uint32_t ub = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t tmem = __shfl_sync(0xffffffffu, holder, 0);
    const uint64_t a0 = AMN ? make_smem_desc_sw128(ub, 128 * 128, 1024) : make_smem_desc_sw128(ub, 16, 1024);
    const uint64_t b0 = BMN ? make_smem_desc_sw128(ub + 65536, 256 * 128, 1024) : make_smem_desc_sw128(ub + 65536, 16, 1024);
    const uint32_t alo = (uint32_t)a0, ahi = (uint32_t)(a0 >> 32), blo = (uint32_t)b0, bhi = (uint32_t)(b0 >> 32);
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < NMMA / 4; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = tmem + (uint32_t)(k % NACC) * (uint32_t)N;
          const uint32_t acc = (i > 0 || k >= NACC) ? 1u : 0u;
          if (TS) umma_f16_ts_elect(d, tmem + 448, blo + k * (BMN ? 128 : 2), bhi, idesc, acc);
          else umma_f16_elect2(d, alo + k * (AMN ? 128 : 2), ahi, blo + k * (BMN ? 128 : 2), bhi, idesc, acc);
        }
      }
      const long long t1 = clock64();
      umma_commit_elect(bar);
      mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      __syncwarp();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(holder, 512);
}

template <int M, int N, int NACC, bool TS, bool AMN, bool BMN, int NMMA>
void run(long long* d) {
  auto k = bench<M, N, NACC, TS, AMN, BMN, NMMA>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<1, 128, 200 * 1024>>>(d);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("M=%3d N=%3d acc=%d ts=%d a_mn=%d b_mn=%d n=%3d : issue %.1f cyc/mma, complete %.1f cyc/mma (total %lld)\n", M, N, NACC, (int)TS,
         (int)AMN, (int)BMN, NMMA, (double)h[0] / NMMA, (double)h[1] / NMMA, h[1]);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<128, 64, 1, false, false, false, 96>(d);
  run<128, 64, 4, false, false, false, 96>(d);
  run<128, 256, 1, false, false, false, 96>(d);
  run<128, 128, 1, false, false, false, 96>(d);
  run<128, 32, 1, false, false, false, 96>(d);
  run<128, 16, 1, false, false, false, 96>(d);
  run<128, 16, 4, false, false, false, 96>(d);
  run<64, 64, 1, false, false, false, 96>(d);
  run<64, 64, 4, false, false, false, 96>(d);
  run<128, 64, 1, true, false, false, 96>(d);
  run<128, 64, 4, true, false, false, 96>(d);
  run<128, 16, 1, true, false, false, 96>(d);
  run<128, 64, 1, false, false, true, 96>(d);
  run<64, 64, 1, false, true, true, 96>(d);
  run<64, 64, 4, false, true, true, 96>(d);
  run<64, 16, 1, false, true, true, 96>(d);
  run<64, 32, 1, false, true, true, 96>(d);
  run<128, 64, 1, false, false, false, 8>(d);
  run<128, 64, 1, false, false, false, 24>(d);
  run<128, 64, 1, true, false, false, 24>(d);
  return 0;
}
