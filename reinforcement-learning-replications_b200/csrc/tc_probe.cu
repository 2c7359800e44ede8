// tc_probe: diagnostic kernel that issues an arbitrary chain of tcgen05.mma (kind::tf32, cta_group::1) instructions
// on a host-supplied shared-memory image and returns the raw TMEM contents.  Used by tests/test_gpu_tc_probe.py to pin
// the shared-memory descriptor / instruction descriptor / TMEM layouts that mlp_tc.cu relies on (K-major and
// MN-major SWIZZLE_128B operands, M = 64 / 128 accumulators) against a numpy matmul.  Not on the product path.
#include "common.cuh"
#include "tc_common.cuh"

namespace b200rl {

struct ProbeMma {
  unsigned long long adesc;  // start-address field relative to the 1024-aligned image base
  unsigned long long bdesc;
  unsigned int idesc;
  unsigned int dcol;        // TMEM column offset of D
  unsigned int accumulate;  // 0: D = A*B, 1: D += A*B
  unsigned int kind;        // 0: kind::tf32, 1: kind::f16
};

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const uint32_t* __restrict__ image, int image_words,
                                                          const ProbeMma* __restrict__ mmas, int n_mma, int read_cols,
                                                          float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint32_t* img = reinterpret_cast<uint32_t*>(smem_raw + (base - raw));
  for (int i = tid; i < image_words; i += blockDim.x) img[i] = image[i];
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_holder;
  // zero the columns that will be read back so untouched cells are well defined
  for (int c = 0; c < read_cols; c += 8) tmem_st_zero8(tmem + ((uint32_t)(32 * warp) << 16) + c);
  tmem_wait_st();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) {
    for (int i = 0; i < n_mma; ++i) {
      const ProbeMma m = mmas[i];
      const unsigned long long a = m.adesc + (unsigned long long)(base >> 4);
      const unsigned long long b = m.bdesc + (unsigned long long)(base >> 4);
      if (m.kind == 1) umma_f16(tmem + m.dcol, a, b, m.idesc, m.accumulate);
      else umma_tf32(tmem + m.dcol, a, b, m.idesc, m.accumulate);
    }
    umma_commit(smem_u32(&mbar));
  }
  mbar_wait(smem_u32(&mbar), 0);
  tc_fence_after_sync();
  for (int c = 0; c < read_cols; c += 8) {
    uint32_t v[8];
    tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + c, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(size_t)(32 * warp + lane) * read_cols + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_tc_probe(const uint32_t* image_dev, int image_words, const void* mmas_dev, int n_mma,
                               int read_cols, float* out_dev, void* stream) {
  B200RL_REQUIRE(image_dev && mmas_dev && out_dev, "tc_probe: NULL argument");
  B200RL_REQUIRE(image_words > 0 && image_words * 4 <= 200 * 1024, "tc_probe: image must be 1..200 KiB");
  B200RL_REQUIRE(read_cols > 0 && read_cols <= 512 && read_cols % 8 == 0, "tc_probe: read_cols must be 8..512, x8");
  const size_t smem = (size_t)image_words * 4 + 1024;
  B200RL_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      image_dev, image_words, static_cast<const ProbeMma*>(mmas_dev), n_mma, read_cols, out_dev);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
