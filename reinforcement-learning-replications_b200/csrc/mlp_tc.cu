// mlp_tc: the fused MLP step on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract and data flow as mlp_fused.cu (see there for the reference file:line map); this kernel is selected
// for the shapes the reference's on-policy recipes use: 3 Linear layers, hidden widths 64/64, tanh hidden,
// identity output, obs width <= 32, output width <= 15.  Everything else takes the fp32 kernel.
//
// Precision: tcgen05 has no fp32 MMA kind and plain TF32 violates the 1e-5 parity bar, so every fp32
// operand is split into THREE bf16 values x = h + m + l (24 mantissa bits) and every logical product is the six
// kind::f16 MMAs  mm + hl + lh + hm + mh + hh  with fp32 accumulation in TMEM -- measured 1.3e-7 relative error
// against float64 (tests/test_gpu_tc_probe.py).  bf16 (not tf32) because a SWIZZLE_128B buffer of 16-bit elements can
// be read BOTH K-major (activations as the A operand of the next layer) and MN-major (the same activations as an
// operand of the dW = dZ^T X product, whose reduction runs over the tile's rows); tf32 MN-major needs a different
// swizzle, i.e. a second copy of every activation (pinned on hardware by the probe tests).
//
// Per CTA: 128-row tiles, persistent over tiles (grid = min(#tiles, #SMs)), 8 epilogue warps + 1 MMA-issuing warp.
//   shared memory  operand buffers, 64 bf16 columns x 128-byte rows, SWIZZLE_128B, three splits each:
//                  XD [128][64]: obs in cols 0..31, dOut in cols 32..46, ones in col 47 (bias gradients for free)
//                  H1, H2 [128][64]: activations, overwritten in place by dZ1 / dZ2 during the backward pass
//                  W1, W2 [64][64], W3 [16][64]: torch [out][in] order = K-major B operand in the forward pass and,
//                  unchanged, MN-major B operand of dX = dZ W in the backward pass
//   tensor memory  Z1/H1, Z2/H2 (fp32, kept for tanh'), OUT, dH2, dH1 (M = 128) and the per-CTA gradient
//                  accumulators dW2, dW1, dW3^T, db2, db1 (M = 64) that persist across the CTA's tiles.
//   stages / tile  X -> [F1] -> tanh -> [F2] -> tanh -> [F3] -> log-prob/loss/dOut -> [dW3^T, dH2] -> dZ2 ->
//                  [dW2, db2, dH1] -> dZ1 -> [dW1, db1]; one tcgen05.commit + mbarrier wait per bracketed stage.
#include <cuda_bf16.h>

#include <cmath>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200rl {

constexpr int TC_ROWS = 128;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1) * 32;
constexpr float TC_LOG_SQRT_2PI = 0.91893853320467274178f;
constexpr float TC_ENT_CONST = 1.4189385332046727418f;

// shared-memory map (bytes from the 1024-aligned base)
constexpr uint32_t ACT_BUF = 128 * 128;  // one split of a [128][64] bf16 buffer
constexpr uint32_t W_BUF = 64 * 128;     // one split of a [64][64] weight
constexpr uint32_t W3_BUF = 16 * 128;    // one split of the [16][64] output weight
constexpr uint32_t SM_XD = 0;
constexpr uint32_t SM_H1 = SM_XD + 3 * ACT_BUF;
constexpr uint32_t SM_H2 = SM_H1 + 3 * ACT_BUF;
constexpr uint32_t SM_W1 = SM_H2 + 3 * ACT_BUF;
constexpr uint32_t SM_W2 = SM_W1 + 3 * W_BUF;
constexpr uint32_t SM_W3 = SM_W2 + 3 * W_BUF;
constexpr uint32_t SM_OPERANDS_END = SM_W3 + 3 * W3_BUF;
constexpr uint32_t SM_BIAS = SM_OPERANDS_END;       // b1[64] b2[64] b3[16] floats
constexpr uint32_t SM_DIST = SM_BIAS + 1024;        // var[16], log_scale[16] floats
constexpr uint32_t SM_DB3 = SM_DIST + 256;          // [4 warps][16] floats
constexpr uint32_t SM_STAGE = SM_DB3 + 512;         // fp32 staging of the NEXT tile's observations [128][n_in<=32]
constexpr uint32_t SM_TOTAL = SM_STAGE + 128 * 32 * 4;
constexpr uint32_t TC_SMEM_BYTES = SM_TOTAL + 1024;  // + alignment slack

// tensor-memory column map
constexpr uint32_t TM_Z1 = 0, TM_Z2 = 64, TM_OUT = 128, TM_DH2 = 160, TM_DH1 = 224, TM_DW2 = 288, TM_DW1 = 352,
                   TM_DW3 = 384, TM_DB2 = 400, TM_DB1 = 416;

struct TcArgs {
  int n_in, n_out;
  int h1, h2;  // hidden widths (<= 64, zero-padded to the 64-wide buffers)
  int w_off[3], b_off[3], P;
  int loss, dist;
  long long n_rows;
  float inv_n, clip_lo, clip_hi;
  const float* params;
  const float* obs;
  const float* actions;
  const float* log_std;
  const float* adv_raw;
  const double* adv_stats;
  const float* old_logp;
  const float* target;
  float* row_out;
  float* partials;
  double* scalar_partials;
  const int* skip_flag;
  const unsigned* run_if;  // when set: run only if *run_if == seq (wide-range re-run of an mlp_tc2 launch)
  unsigned seq;
  int total_rows;          // partial rows the consumer reduces (> gridDim.x when standing in for mlp_tc2)
};

// byte offset of element (r, c) inside one split buffer (c < 64)
__device__ __forceinline__ uint32_t rel_rc(int r, int c) {
  return (uint32_t)r * 128u + ((uint32_t)((c >> 3) ^ (r & 7)) << 4) + ((uint32_t)(c & 7) << 1);
}

__device__ __forceinline__ void split2(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l) {
  const __nv_bfloat162 hb = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(hb);
  const float r0 = x0 - hf.x, r1 = x1 - hf.y;
  const __nv_bfloat162 mb = __floats2bfloat162_rn(r0, r1);
  const float2 mf = __bfloat1622float2(mb);
  const __nv_bfloat162 lb = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
  h = *reinterpret_cast<const uint32_t*>(&hb);
  m = *reinterpret_cast<const uint32_t*>(&mb);
  l = *reinterpret_cast<const uint32_t*>(&lb);
}

// write 8 consecutive columns (one 16-byte chunk `ch`) of row r into the three split buffers starting at `buf`
__device__ __forceinline__ void store_chunk3(uint8_t* sm, uint32_t buf, uint32_t split_stride, int r, int ch,
                                             const float (&x)[8]) {
  uint4 h, m, l;
  split2(x[0], x[1], h.x, m.x, l.x);
  split2(x[2], x[3], h.y, m.y, l.y);
  split2(x[4], x[5], h.z, m.z, l.z);
  split2(x[6], x[7], h.w, m.w, l.w);
  const uint32_t off = buf + (uint32_t)r * 128u + ((uint32_t)(ch ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(sm + off) = h;
  *reinterpret_cast<uint4*>(sm + off + split_stride) = m;
  *reinterpret_cast<uint4*>(sm + off + 2 * split_stride) = l;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// One operand of a split product: 32-bit halves of the descriptor of split 0 / k-step 0, the low-word advance per
// split (buffer stride >> 4) and per k-step (bytes >> 4).  All fields are warp-uniform.
struct OpDesc {
  uint32_t lo, hi, split_step, k_step;
};
__device__ __forceinline__ OpDesc op_kmajor(uint32_t addr, uint32_t split_bytes) {  // K along the 128-byte rows
  const uint64_t d = make_smem_desc_sw128(addr, 16, 1024);
  return OpDesc{(uint32_t)d, (uint32_t)(d >> 32), split_bytes >> 4, 32u >> 4};
}
__device__ __forceinline__ OpDesc op_mnmajor(uint32_t addr, uint32_t rows, uint32_t split_bytes) {  // K along rows
  const uint64_t d = make_smem_desc_sw128(addr, rows * 128, 1024);
  return OpDesc{(uint32_t)d, (uint32_t)(d >> 32), split_bytes >> 4, 2048u >> 4};
}

// the six split products, smallest terms first: (m,m) (h,l) (l,h) (h,m) (m,h) (h,h); k-loop kept rolled (code size)
__device__ __forceinline__ void issue6(uint32_t d_tmem, uint32_t idesc, int ksteps, bool accumulate_first,
                                       const OpDesc a, const OpDesc b) {
  constexpr int TI[6] = {1, 0, 2, 0, 1, 0};
  constexpr int TJ[6] = {1, 2, 0, 1, 0, 0};
  uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    uint32_t alo = a.lo + TI[t] * a.split_step, blo = b.lo + TJ[t] * b.split_step;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
      umma_f16_elect2(d_tmem, alo, a.hi, blo, b.hi, idesc, acc);
      acc = 1u;
      alo += a.k_step;
      blo += b.k_step;
    }
  }
}
// A (three splits) times an operand that is exact in bf16 (the ones column, split 0 only): three products
__device__ __forceinline__ void issue3(uint32_t d_tmem, uint32_t idesc, int ksteps, bool accumulate_first,
                                       const OpDesc a, const OpDesc b) {
  uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll
  for (int sp = 2; sp >= 0; --sp) {
    uint32_t alo = a.lo + sp * a.split_step, blo = b.lo;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
      umma_f16_elect2(d_tmem, alo, a.hi, blo, b.hi, idesc, acc);
      acc = 1u;
      alo += a.k_step;
      blo += b.k_step;
    }
  }
}

#ifdef B200RL_TC_TIMING
__device__ unsigned long long g_tc_t[16];
#define TC_T(i)                                   \
  do {                                            \
    if (tid == 0) {                               \
      const long long _n = clock64();             \
      tacc[i] += (unsigned long long)(_n - tlast); \
      tlast = _n;                                 \
    }                                             \
  } while (0)
#else
#define TC_T(i)
#endif

__device__ unsigned long long g_tc_fallbacks;  // launches of this kernel that actually re-ran an mlp_tc2 launch

template <bool BACKWARD>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_holder;
  __shared__ double s_sc[6][TC_EPI_WARPS];
  if (p.skip_flag != nullptr && *p.skip_flag != 0) return;  // early stop: whole launch is a no-op
  if (p.run_if != nullptr && *p.run_if != p.seq) return;    // the fp16 kernel's result stands
  if (p.run_if != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_tc_fallbacks, 1ull);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: role branches need no vote
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1024-byte aligned
  uint8_t* sm = smem_raw + (base - raw);
  float* s_bias = reinterpret_cast<float*>(sm + SM_BIAS);
  float* s_dist = reinterpret_cast<float*>(sm + SM_DIST);
  float* s_db3 = reinterpret_cast<float*>(sm + SM_DB3);
  const int n_in = p.n_in, A_out = p.n_out, h1 = p.h1, h2 = p.h2;
  // partial rows beyond this grid (the consumer was sized for mlp_tc2's two rows per CTA) contribute nothing
  for (int row = (int)gridDim.x + (int)blockIdx.x; row < p.total_rows; row += (int)gridDim.x) {
    if (BACKWARD)
      for (int i = tid; i < p.P; i += TC_THREADS) p.partials[(size_t)row * p.P + i] = 0.f;
    if (p.scalar_partials != nullptr && tid < B200RL_N_SCALARS) p.scalar_partials[(size_t)row * B200RL_N_SCALARS + tid] = 0.0;
  }

  // ---- one-time setup: zero operand buffers, stage W (three bf16 splits), biases, distribution constants ----
  for (uint32_t i = tid; i < SM_OPERANDS_END / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    auto put = [&](uint32_t buf, uint32_t stride, int r, int c, float x) {
      const __nv_bfloat16 hb = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(hb);
      const __nv_bfloat16 mb = __float2bfloat16_rn(r1);
      const __nv_bfloat16 lb = __float2bfloat16_rn(r1 - __bfloat162float(mb));
      const uint32_t off = buf + rel_rc(r, c);
      *reinterpret_cast<__nv_bfloat16*>(sm + off) = hb;
      *reinterpret_cast<__nv_bfloat16*>(sm + off + stride) = mb;
      *reinterpret_cast<__nv_bfloat16*>(sm + off + 2 * stride) = lb;
    };
    for (int idx = tid; idx < h1 * n_in; idx += TC_THREADS)
      put(SM_W1, W_BUF, idx / n_in, idx % n_in, __ldg(p.params + p.w_off[0] + idx));
    for (int idx = tid; idx < h2 * h1; idx += TC_THREADS)
      put(SM_W2, W_BUF, idx / h1, idx % h1, __ldg(p.params + p.w_off[1] + idx));
    for (int idx = tid; idx < A_out * h2; idx += TC_THREADS)
      put(SM_W3, W3_BUF, idx / h2, idx % h2, __ldg(p.params + p.w_off[2] + idx));
    for (int i = tid; i < 64; i += TC_THREADS) {
      s_bias[i] = i < h1 ? __ldg(p.params + p.b_off[0] + i) : 0.f;
      s_bias[64 + i] = i < h2 ? __ldg(p.params + p.b_off[1] + i) : 0.f;
    }
    for (int i = tid; i < 16; i += TC_THREADS) s_bias[128 + i] = i < A_out ? __ldg(p.params + p.b_off[2] + i) : 0.f;
    if (p.dist == B200RL_DIST_GAUSSIAN)
      for (int a = tid; a < A_out; a += TC_THREADS) {
        const float scale = expf(__ldg(p.log_std + a));  // gaussian_policy.py:34
        s_dist[a] = scale * scale;                       // Normal.log_prob: var = scale ** 2
        s_dist[16 + a] = logf(scale);
      }
  }
  if (warp == TC_EPI_WARPS) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_holder;
  const uint32_t bar = smem_u32(&mbar);

  const long long num_tiles = (p.n_rows + TC_ROWS - 1) / TC_ROWS;
  constexpr int STAGES = BACKWARD ? 6 : 3;

  if (warp == TC_EPI_WARPS) {
    // =============================== MMA issuer warp =================================================
    const uint32_t I_128_64_KK = make_idesc_bf16(128, 64, 0, 0), I_128_16_KK = make_idesc_bf16(128, 16, 0, 0);
    const uint32_t I_128_64_KM = make_idesc_bf16(128, 64, 0, 1), I_64_64_MM = make_idesc_bf16(64, 64, 1, 1);
    const uint32_t I_64_32_MM = make_idesc_bf16(64, 32, 1, 1), I_64_16_MM = make_idesc_bf16(64, 16, 1, 1);
    // warp-uniform copies (ptxas keeps them in uniform registers: no per-instruction R2UR)
    const uint32_t ub = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t ut = __shfl_sync(0xffffffffu, tmem, 0);
    const OpDesc XD_K = op_kmajor(ub + SM_XD, ACT_BUF), H1_K = op_kmajor(ub + SM_H1, ACT_BUF),
                 H2_K = op_kmajor(ub + SM_H2, ACT_BUF), W1_K = op_kmajor(ub + SM_W1, W_BUF),
                 W2_K = op_kmajor(ub + SM_W2, W_BUF), W3_K = op_kmajor(ub + SM_W3, W3_BUF);
    const OpDesc XD_K2 = op_kmajor(ub + SM_XD + 64, ACT_BUF);  // cols 32..47 (dOut) as a K-major A operand
    const OpDesc H1_M = op_mnmajor(ub + SM_H1, 128, ACT_BUF), H2_M = op_mnmajor(ub + SM_H2, 128, ACT_BUF),
                 XD_M0 = op_mnmajor(ub + SM_XD, 128, ACT_BUF),        // X    (cols 0..31)
                 XD_M32 = op_mnmajor(ub + SM_XD + 64, 128, ACT_BUF),  // dOut (cols 32..47, col 47 = ones)
                 W2_M = op_mnmajor(ub + SM_W2, 64, W_BUF), W3_M = op_mnmajor(ub + SM_W3, 16, W3_BUF);
    bool first = true;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int s = 0; s < STAGES; ++s) {
        __syncthreads();  // operands of stage s are in shared memory (and the previous stage's MMAs have retired)
        tc_fence_after_sync();
        if (s == 0) {  // Z1 = X W1^T
          issue6(ut + TM_Z1, I_128_64_KK, 2, false, XD_K, W1_K);
        } else if (s == 1) {  // Z2 = H1 W2^T
          issue6(ut + TM_Z2, I_128_64_KK, 4, false, H1_K, W2_K);
        } else if (s == 2) {  // OUT = H2 W3^T
          issue6(ut + TM_OUT, I_128_16_KK, 4, false, H2_K, W3_K);
        } else if (s == 3) {
          // dW3^T[i][o] += sum_r H2[r][i] dOut[r][o]   (both read MN-major: the reduction runs over rows)
          issue6(ut + TM_DW3, I_64_16_MM, 8, !first, H2_M, XD_M32);
          // dH2 = dOut W3   (A: XD cols 32..47; B: W3 read MN-major, K = output index)
          issue6(ut + TM_DH2, I_128_64_KM, 1, false, XD_K2, W3_M);
        } else if (s == 4) {
          // dW2[o][i] += sum_r dZ2[r][o] H1[r][i] ; db2[o] += sum_r dZ2[r][o] * 1 ; dH1 = dZ2 W2
          issue6(ut + TM_DW2, I_64_64_MM, 8, !first, H2_M, H1_M);
          issue3(ut + TM_DB2, I_64_16_MM, 8, !first, H2_M, XD_M32);
          issue6(ut + TM_DH1, I_128_64_KM, 4, false, H2_K, W2_M);
        } else {
          // dW1[o][i] += sum_r dZ1[r][o] X[r][i] ; db1[o] += sum_r dZ1[r][o]
          issue6(ut + TM_DW1, I_64_32_MM, 8, !first, H1_M, XD_M0);
          issue3(ut + TM_DB1, I_64_16_MM, 8, !first, H1_M, XD_M32);
        }
        umma_commit_elect(bar);
        __syncwarp();
      }
      first = false;
    }
  } else {
    // =============================== epilogue warps ==================================================
    const int q = warp & 3, half = warp >> 2;
    const int r = 32 * q + lane;                         // row of the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;  // this warp's TMEM lane quadrant
    const int c0 = 32 * half;                            // this warp's column half
    uint32_t phase = 0;

    float adv_mean = 0.f, adv_std = 1.f;  // normalize_tensor (utils.py:90-92): mean, UNBIASED std, no epsilon
    if (p.adv_stats != nullptr) {
      const double s1 = p.adv_stats[0], s2 = p.adv_stats[1], cnt = p.adv_stats[2];
      const double mean = s1 / cnt;
      adv_mean = (float)mean;
      adv_std = (float)sqrt((s2 - cnt * mean * mean) / (cnt - 1.0));
    }
    double sc[6] = {0, 0, 0, 0, 0, 0};
    float db3[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) db3[a] = 0.f;

    // tanh layer epilogue: Z (TMEM) + bias -> tanh -> fp32 copy back to TMEM (for tanh') + bf16 splits to smem
    auto act_epilogue = [&](uint32_t tm_col, const float* bias, uint32_t dst_buf) {
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {  // 16 columns at a time keeps the live register set small
        const int cs = c0 + 16 * sub;
        uint32_t v[16];
        tmem_ld16(tmem + lane_addr + tm_col + cs, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(tanhf(__uint_as_float(v[j]) + bias[cs + j]));
        if (BACKWARD) tmem_st16(tmem + lane_addr + tm_col + cs, v);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[8 * ch + j]);
          store_chunk3(sm, dst_buf, ACT_BUF, r, (cs >> 3) + ch, x);
        }
      }
      if (BACKWARD) tmem_wait_st();
    };
    // backward epilogue: dZ = dH * (1 - H^2), bf16 splits over the activation buffer (in place)
    auto dz_epilogue = [&](uint32_t tm_dh, uint32_t tm_h, uint32_t dst_buf) {
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int cs = c0 + 16 * sub;
        uint32_t g[16], h[16];
        tmem_ld16(tmem + lane_addr + tm_dh + cs, g);
        tmem_ld16(tmem + lane_addr + tm_h + cs, h);
        tmem_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float hv = __uint_as_float(h[8 * ch + j]);
            x[j] = __uint_as_float(g[8 * ch + j]) * (1.f - hv * hv);
          }
          store_chunk3(sm, dst_buf, ACT_BUF, r, (cs >> 3) + ch, x);
        }
      }
    };
    auto stage_done = [&]() {  // publish smem writes to the tensor core, hand over to the issuer, wait for its MMAs
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      mbar_wait(bar, phase);
      phase ^= 1u;
      tc_fence_after_sync();
    };

    float pf_act[15], pf_adv = 0.f, pf_old = 0.f, pf_tgt = 0.f;
#pragma unroll
    for (int a = 0; a < 15; ++a) pf_act[a] = 0.f;
    const float* s_stage = reinterpret_cast<const float*>(sm + SM_STAGE);
    // stage one tile's observations (contiguous rows_here*n_in floats, 16-byte aligned) with cp.async; warps 4..7
    auto stage_obs = [&](long long t) {
      if (half == 1 && t < num_tiles) {
        const long long r0 = t * TC_ROWS;
        const long long rows_here = (p.n_rows - r0) < TC_ROWS ? (p.n_rows - r0) : TC_ROWS;
        const int n16 = (int)((rows_here * n_in * 4 + 15) / 16);  // the obs buffer is padded to 16 bytes by the engine
        const char* g = reinterpret_cast<const char*>(p.obs + r0 * n_in);
        for (int i = tid - 128; i < n16; i += 128) cp_async16(base + SM_STAGE + 16 * i, g + 16 * (size_t)i);
      }
      cp_async_commit_wait_all();
    };
    stage_obs(blockIdx.x);
    asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");

#ifdef B200RL_TC_TIMING
    unsigned long long tacc[16];
    for (int i = 0; i < 16; ++i) tacc[i] = 0;
    long long tlast = clock64();
#endif
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const long long row = tile * TC_ROWS + r;
      const bool valid = row < p.n_rows;

      // ---- observations: staged as fp32 by cp.async (previous tile / prologue); one row per thread of warps 0..3
      //      converts its row to the three bf16 splits in cols 0..31 of XD ----
      if (half == 0) {
        const float* src = s_stage + r * n_in;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = 8 * ch + j;
            x[j] = (valid && c < n_in) ? src[c] : 0.f;
          }
          store_chunk3(sm, SM_XD, ACT_BUF, r, ch, x);
        }
        // loss inputs of this row: issue the loads now, consume them three stages later
        if (valid) {
          if (p.dist == B200RL_DIST_GAUSSIAN) {
#pragma unroll
            for (int a = 0; a < 15; ++a)
              if (a < A_out) pf_act[a] = __ldg(p.actions + row * A_out + a);
          } else if (p.dist == B200RL_DIST_CATEGORICAL) {
            pf_act[0] = __ldg(p.actions + row);
          }
          if (p.loss != B200RL_LOSS_EVAL && p.adv_raw != nullptr) pf_adv = __ldg(p.adv_raw + row);
          if (p.old_logp != nullptr) pf_old = __ldg(p.old_logp + row);
          if (p.loss == B200RL_LOSS_MSE) pf_tgt = __ldg(p.target + row);
        }
      }
      TC_T(0);
      stage_done();                                   // F1
      TC_T(1);
      act_epilogue(TM_Z1, s_bias, SM_H1);
      TC_T(2);
      stage_done();                                   // F2
      TC_T(3);
      act_epilogue(TM_Z2, s_bias + 64, SM_H2);
      TC_T(4);
      stage_done();                                   // F3
      TC_T(5);

      // warps 4..7 have no loss work: they fetch the next tile's observations into the staging buffer meanwhile
      stage_obs(tile + gridDim.x);

      // ---- distribution / loss epilogue (one thread per row: warps 0..3) ----
      if (half == 0) {
        uint32_t o[16];
        tmem_ld16(tmem + lane_addr + TM_OUT, o);
        tmem_wait_ld();
        float out[16], dout[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) {
          out[a] = __uint_as_float(o[a]) + s_bias[128 + a];
          dout[a] = 0.f;
        }
        if (valid) {
          float coef = 0.f, term = 0.f, lp = 0.f, ent = 0.f;
          if (p.dist == B200RL_DIST_NONE) {
            const float vout = out[0];
            if (p.row_out) p.row_out[row] = vout;
            if (p.loss == B200RL_LOSS_MSE) {  // ppo.py:282-287
              const float diff = vout - pf_tgt;
              term = diff * diff;
              dout[0] = (2.f * diff) * p.inv_n;
            }
            sc[0] += (double)term;
            sc[5] += 1.0;
          } else {
            float dlp[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) dlp[a] = 0.f;
            if (p.dist == B200RL_DIST_GAUSSIAN) {
#pragma unroll
              for (int a = 0; a < 15; ++a)
                if (a < A_out) {
                  const float var = s_dist[a], lsc = s_dist[16 + a];
                  const float d = pf_act[a] - out[a];
                  lp += -(d * d) / (2.f * var) - lsc - TC_LOG_SQRT_2PI;  // torch Normal.log_prob
                  ent += TC_ENT_CONST + lsc;                             // torch Normal.entropy
                  dlp[a] = d / var;
                }
            } else {
              float m = out[0];
#pragma unroll
              for (int a = 1; a < 15; ++a)
                if (a < A_out) m = fmaxf(m, out[a]);
              float se = 0.f;
#pragma unroll
              for (int a = 0; a < 15; ++a)
                if (a < A_out) se += expf(out[a] - m);
              const float lse = m + logf(se);
              const int ai = (int)pf_act[0];  // value.long()
#pragma unroll
              for (int a = 0; a < 15; ++a)
                if (a < A_out) {
                  const float lg = out[a] - lse;
                  const float pa = expf(lg);
                  ent -= lg * pa;
                  if (a == ai) lp = lg;
                  dlp[a] = (a == ai ? 1.f : 0.f) - pa;
                }
            }
            if (p.row_out) p.row_out[row] = lp;
            float adv = 0.f, oldlp = 0.f;
            if (p.loss != B200RL_LOSS_EVAL) {
              adv = pf_adv;
              if (p.adv_stats != nullptr) adv = (adv - adv_mean) / adv_std;  // utils.py:91
            }
            if (p.old_logp != nullptr) oldlp = pf_old;
            if (p.loss == B200RL_LOSS_PPO_CLIP) {  // ppo.py:245-255
              const float ratio = expf(lp - oldlp);
              const float s1 = ratio * adv;
              const float s2 = fminf(fmaxf(ratio, p.clip_lo), p.clip_hi) * adv;
              term = -fminf(s1, s2);
              const bool pass = adv >= 0.f ? (ratio <= p.clip_hi) : (ratio >= p.clip_lo);
              coef = pass ? (-p.inv_n * adv) * ratio : 0.f;
            } else if (p.loss == B200RL_LOSS_VPG) {  // vpg.py:203
              term = -(lp * adv);
              coef = -p.inv_n * adv;
            } else if (p.loss == B200RL_LOSS_TRPO_SURROGATE) {  // trpo.py:161-163
              const float ratio = expf(lp - oldlp);
              term = -(ratio * adv);
              coef = (-p.inv_n * adv) * ratio;
            }
#pragma unroll
            for (int a = 0; a < 15; ++a) dout[a] = coef * dlp[a];
            sc[0] += (double)term;
            if (p.old_logp != nullptr) sc[1] += (double)(oldlp - lp);
            sc[2] += (double)ent;
            sc[3] += (double)lp;
            sc[4] += (double)lp * (double)lp;
            sc[5] += 1.0;
          }
        }
        if (BACKWARD) {
#pragma unroll
          for (int a = 0; a < 15; ++a) db3[a] += dout[a];
          dout[15] = 1.0f;  // ones column: db1 / db2 fall out of the dW tensor-core products
          float x0[8], x1[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            x0[j] = dout[j];
            x1[j] = dout[8 + j];
          }
          store_chunk3(sm, SM_XD, ACT_BUF, r, 4, x0);  // cols 32..39
          store_chunk3(sm, SM_XD, ACT_BUF, r, 5, x1);  // cols 40..47
        }
      }
      TC_T(6);
      if (BACKWARD) {
        stage_done();                                 // dW3^T, dH2
        TC_T(7);
        dz_epilogue(TM_DH2, TM_Z2, SM_H2);
        TC_T(8);
        stage_done();                                 // dW2, db2, dH1
        TC_T(9);
        dz_epilogue(TM_DH1, TM_Z1, SM_H1);
        TC_T(10);
        stage_done();                                 // dW1, db1
        TC_T(11);
      } else {
        // forward only: the next tile's F1 may not overwrite TM_OUT/XD before everyone has read them
        tc_fence_before_sync();
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
        tc_fence_after_sync();
      }
    }

#ifdef B200RL_TC_TIMING
    if (tid == 0 && blockIdx.x == 0 && BACKWARD)
      for (int i = 0; i < 16; ++i) g_tc_t[i] = tacc[i];
#endif
    // ---- per-CTA results: gradient accumulators (TMEM, M = 64 layout: row m -> lane (m%16) + 32*(m/16)) ----
    if (BACKWARD && half == 0) {
      float* dst = p.partials + (size_t)blockIdx.x * p.P;
      const int m = 16 * q + lane;  // valid for lane < 16
      uint32_t v[32];
      for (int cb = 0; cb < 2; ++cb) {  // dW2 [64 o][64 i]
        tmem_ld32(tmem + lane_addr + TM_DW2 + 32 * cb, v);
        tmem_wait_ld();
        if (lane < 16 && m < h2)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (32 * cb + j < h1) dst[p.w_off[1] + m * h1 + 32 * cb + j] = __uint_as_float(v[j]);
      }
      tmem_ld32(tmem + lane_addr + TM_DW1, v);  // dW1 [64 o][32 i]
      tmem_wait_ld();
      if (lane < 16 && m < h1)
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < n_in) dst[p.w_off[0] + m * n_in + j] = __uint_as_float(v[j]);
      uint32_t w[16];
      tmem_ld16(tmem + lane_addr + TM_DW3, w);  // dW3^T [64 i][16 o]
      tmem_wait_ld();
      if (lane < 16 && m < h2)
#pragma unroll
        for (int a = 0; a < 15; ++a)
          if (a < A_out) dst[p.w_off[2] + a * h2 + m] = __uint_as_float(w[a]);
      tmem_ld16(tmem + lane_addr + TM_DB2, w);  // column 15 = sum_r dZ2[r][o]
      tmem_wait_ld();
      if (lane < 16 && m < h2) dst[p.b_off[1] + m] = __uint_as_float(w[15]);
      tmem_ld16(tmem + lane_addr + TM_DB1, w);
      tmem_wait_ld();
      if (lane < 16 && m < h1) dst[p.b_off[0] + m] = __uint_as_float(w[15]);
      // db3: fixed-order reduction of the per-row accumulators
#pragma unroll
      for (int a = 0; a < 15; ++a) {
        float s = db3[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_db3[q * 16 + a] = s;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
    if (BACKWARD && tid < A_out) {
      float s = 0.f;
      for (int w4 = 0; w4 < 4; ++w4) s += s_db3[w4 * 16 + tid];
      p.partials[(size_t)blockIdx.x * p.P + p.b_off[2] + tid] = s;
    }
    if (p.scalar_partials != nullptr) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double v = warp_sum(sc[k]);
        if (lane == 0) s_sc[k][warp] = v;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
      if (tid < B200RL_N_SCALARS) {
        double t = 0.0;
        if (tid < 6)
          for (int w8 = 0; w8 < TC_EPI_WARPS; ++w8) t += s_sc[tid][w8];
        p.scalar_partials[(size_t)blockIdx.x * B200RL_N_SCALARS + tid] = t;
      }
    }
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tmem_dealloc(tmem, 512);
}

#ifdef B200RL_TC_TIMING
extern "C" int b200rl_debug_tc_timing(unsigned long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, g_tc_t, sizeof(unsigned long long) * 16);
}
#endif

bool tc_shape_ok(const b200rl_mlp_desc& d) {
  return d.n_layers == 3 && d.sizes[1] >= 1 && d.sizes[1] <= 64 && d.sizes[2] >= 1 && d.sizes[2] <= 64 &&
         d.sizes[0] >= 1 && d.sizes[0] <= 32 &&
         d.sizes[3] >= 1 && d.sizes[3] <= 15 && d.hidden_act == B200RL_ACT_TANH && d.out_act == B200RL_ACT_IDENTITY;
}

int tc_grid(int64_t n_rows) {
  const int64_t tiles = (n_rows + TC_ROWS - 1) / TC_ROWS;
  const int sms = device_sm_count();
  if (sms <= 0) return -1;
  return (int)(tiles < sms ? (tiles < 1 ? 1 : tiles) : sms);
}

static int launch_mlp_tc_impl(const b200rl_mlp_loss_grad_args* a, int64_t n_glob, const unsigned* run_if, unsigned seq,
                              int partial_rows, cudaStream_t s) {
  TcArgs k{};
  k.run_if = run_if;
  k.seq = seq;
  k.total_rows = partial_rows;
  k.n_in = a->mlp.sizes[0];
  k.n_out = a->mlp.sizes[3];
  k.h1 = a->mlp.sizes[1];
  k.h2 = a->mlp.sizes[2];
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    k.w_off[l] = off;
    off += a->mlp.sizes[l + 1] * a->mlp.sizes[l];
    k.b_off[l] = off;
    off += a->mlp.sizes[l + 1];
  }
  k.P = off;
  k.loss = a->loss;
  k.dist = a->dist;
  k.n_rows = a->n_rows;
  k.inv_n = 1.0f / (float)n_glob;
  k.clip_lo = (float)(1.0 - (double)a->clip_range);
  k.clip_hi = (float)(1.0 + (double)a->clip_range);
  k.params = a->params;
  k.obs = a->obs;
  k.actions = a->actions;
  k.log_std = a->log_std;
  k.adv_raw = a->adv_raw;
  k.adv_stats = a->adv_stats;
  k.old_logp = a->old_logp;
  k.target = a->target;
  k.row_out = a->row_out;
  k.partials = a->partials;
  k.scalar_partials = a->scalar_partials;
  k.skip_flag = a->skip_flag;
  const int grid = tc_grid(a->n_rows);
  B200RL_REQUIRE(grid > 0, "mlp_tc: no CUDA device");
  if (a->loss != B200RL_LOSS_EVAL) {
    B200RL_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)TC_SMEM_BYTES));
    mlp_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(k);
  } else {
    B200RL_CUDA(cudaFuncSetAttribute(mlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)TC_SMEM_BYTES));
    mlp_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(k);
  }
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

int launch_mlp_tc(const b200rl_mlp_loss_grad_args* a, int64_t n_glob, cudaStream_t s) {
  return launch_mlp_tc_impl(a, n_glob, nullptr, 0u, 0, s);
}
// re-run of an mlp_tc2 launch whose values left the fp16 range: predicated on *run_if == seq
int launch_mlp_tc_fallback(const b200rl_mlp_loss_grad_args* a, int64_t n_glob, const unsigned* run_if, unsigned seq,
                           int partial_rows, cudaStream_t s) {
  return launch_mlp_tc_impl(a, n_glob, run_if, seq, partial_rows, s);
}

}  // namespace b200rl

namespace b200rl {
// device address of the counter, for kernels in other translation units (no relocatable device code in this build)
unsigned long long* tc_fallback_counter_ptr() {
  static unsigned long long* ptr = nullptr;
  if (ptr == nullptr) {
    void* q = nullptr;
    if (cudaGetSymbolAddress(&q, g_tc_fallbacks) == cudaSuccess) ptr = static_cast<unsigned long long*>(q);
  }
  return ptr;
}
}  // namespace b200rl

extern "C" int64_t b200rl_tc_fallback_count(void) {
  unsigned long long v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, b200rl::g_tc_fallbacks, sizeof(v)) != cudaSuccess) return -1;
  return (int64_t)v;
}
