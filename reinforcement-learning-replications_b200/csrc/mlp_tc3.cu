// mlp_tc3: ONE kernel per PPO iteration -- the policy step AND the value step over the same 128-row tile of the batch
// (tcgen05 + TMEM, fp16 x 2 operand splitting as in mlp_tc2.cu; sm_100a only).
//
// Why (reference: /root/reference/src/rl_replicas/algorithms/ppo.py:173-181 policy loop, :186-192 value loop): the two
// loops touch disjoint parameters and read the same fixed inputs (advantages and returns are computed before either
// loop, :142-161), so step i of both can share one pass over the batch.  Compared with two mlp_tc2 launches:
//   * the two tiles ("slots") a CTA keeps in flight are now the POLICY chain and the VALUE chain of the SAME tile:
//     the observation operand is staged once for both;
//   * the observations are split into their fp16 pairs ONCE PER UPDATE by pack_obs_kernel (every step of the update
//     reads the same observations) into ready-made SWIZZLE_128B tile images [128 rows][h cols 0..31 | l cols 32..63];
//     the step kernel brings a tile image in with ONE 16 KB bulk copy (cp.async.bulk + mbarrier complete_tx) issued by
//     the MMA warp a tile ahead -- no epilogue job touches the observations any more (E0 of mlp_tc2 is gone);
//   * column 31 of the image is 1.0, so the bias gradients db1 / db2 still fall out of the weight-gradient products;
//   * dLoss/dOut of both chains goes into the X buffer of the OTHER parity (it is idle between the previous tile's
//     dW1 and the next tile's bulk copy), which is what makes two X buffers fit next to 4 x 32 KB of activations and
//     2 x 28 KB of weights;
//   * tanh'(H1) is taken from the fp16 pair in shared memory like tanh'(H2): no fp32 copy of H1 in tensor memory,
//     which is what makes both chains' accumulators fit (416 of 512 columns);
//   * one reduction + Adam launch for both networks (reduce_adam3_kernel), one all-reduce per iteration.
// A range / precision trip (fp16 operands) raises a sticky flag; the engine then restores its snapshot and redoes the
// update on the two-loop path whose wide-range kernels have no such limits.
#include <cuda_fp16.h>

#include <cmath>

#include "common.cuh"
#include "tc2_common.cuh"
#include "tc_common.cuh"
#include "tc3.cuh"

namespace b200rl {

constexpr int T3_ROWS = 128;
constexpr int T3_EPI_WARPS = 16;
constexpr int T3_EPI_THREADS = T3_EPI_WARPS * 32;
constexpr int T3_THREADS = T3_EPI_THREADS + 32;
constexpr float T3_LOG_SQRT_2PI = 0.91893853320467274178f;
constexpr float T3_ENT_CONST = 1.4189385332046727418f;

// ---- shared-memory map (bytes from the 1024-aligned base) ----
constexpr uint32_t S3_XB = 0;                        // X(k) | dOut(k) buffers, parity k & 1 and (k + 1) & 1
constexpr uint32_t S3_H = 2 * T2_ACT;                // per chain: H1 h, H1 l, H2 h, H2 l
constexpr uint32_t S3_CHAIN = 4 * T2_ACT;
constexpr uint32_t T3_W1T = 32 * 128, T3_W2 = 64 * 128, T3_W3 = 16 * 128;  // one split of each weight operand
constexpr uint32_t S3_W = S3_H + 2 * S3_CHAIN;       // per net: W1T h,l | W2 h,l | W3 h,l
constexpr uint32_t S3_WNET = 2 * T3_W1T + 2 * T3_W2 + 2 * T3_W3;
constexpr uint32_t S3_OPERANDS_END = S3_W + 2 * S3_WNET;
constexpr uint32_t S3_BIAS = S3_OPERANDS_END;        // per net: b1[64] b2[64] b3[16] + pad = 160 floats
constexpr uint32_t S3_DIST = S3_BIAS + 2 * 640;      // var[16], log_scale[16], 1/(2 var)[16], 1/var[16]
constexpr uint32_t S3_SCALE = S3_DIST + 256;         // per net 16 floats
constexpr uint32_t S3_XS = S3_SCALE + 128;           // 2^ex_k [32], 2^-ex_k [32]
constexpr uint32_t S3_RED = S3_XS + 256;             // setup reduction scratch [17 warps][8] floats
constexpr uint32_t S3_BARS = S3_RED + 576;           // ready[2] chain[2] xfull[2] free[2] (8 B each), tmem holder, bad flag
constexpr uint32_t S3_TOTAL = S3_BARS + 128;
constexpr uint32_t T3_SMEM_BYTES = S3_TOTAL + 1024;  // + alignment slack
static_assert(T3_SMEM_BYTES <= 227 * 1024, "mlp_tc3 shared memory");
// end-of-kernel scratch, aliased onto the X buffers (every MMA has retired by then)
constexpr uint32_t S3_END_DB3 = 0;                   // [16 warps][16] floats
constexpr uint32_t S3_END_SC = 1024;                 // [16 warps][8] doubles

// ---- tensor-memory column map (fp32) ----
constexpr uint32_t M3_CHAIN = 80, M3_Z = 0, M3_OUT = 64;          // per chain: Z1 -> Z2 -> dH2 -> dH1 share Z
// per net: DW2 (64) | DB2 (16, col 15 = db2) | DW1 (64: products with X's h columns, then with its l columns; col 31 =
// db1) | DW3 (32: products with dOut's h columns, then l).  The B operands of dW1 / dW3 hold their two splits side
// by side in one swizzle atom, so ONE product per k-step covers both; the halves are added when the accumulators
// are read, once per launch.  2 x 80 + 2 x 176 = 512 columns: all of tensor memory.
constexpr uint32_t M3_ACC = 160, M3_ACC_NET = 176;
constexpr uint32_t M3_DW2 = 0, M3_DB2 = 64, M3_DW1 = 80, M3_DW3 = 144;

#ifdef B200RL_TC3_TIMING
// CTA 0: [0..9] job wait cycles (2 * (stage - 1) + chain), [10..19] job work cycles, [20..29] issuer: wait for the
// stage inputs, [30..39] issuer: issuing, [40] issuer: waiting for the bulk copy, [41] tiles of the CTA
__device__ unsigned long long g_tc3_t[48];
#endif

enum { C3_G = 0, C3_U1, C3_U2, C3_U3, C3_UH2, C3_UH1, C3_W1, C3_W2, C3_W3, C3_OW3, C3_OW2, C3_OW1, C3_OB, C3_N };

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one contiguous span global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// pack_obs: fp32 observations -> per-tile operand images.  Row r of a tile is 128 bytes: fp16 h-splits of the 32
// (zero-padded) scaled features, then their l-splits, 16-byte chunk j stored at chunk position j ^ (r & 7)
// (SWIZZLE_128B).  Feature k is scaled by 2^ex_k (its own maximum parked in [2^12, 2^13)); column 31 holds 1.0.
// The range / precision guards of mlp_tc2's E0 job run here, once per update.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pack_obs_kernel(const float* __restrict__ obs, long long n_rows, int n_in,
                                                       const float* __restrict__ absmax, uint8_t* __restrict__ ximg,
                                                       float* __restrict__ xscale, float* __restrict__ bad_flag) {
  __shared__ float s_xs[32];
  const int r = threadIdx.x;
  bool bad = false;
  if (r < 32) {
    const int e = r < n_in ? fit_exp(__ldg(absmax + r), bad) : 0;
    s_xs[r] = pow2i(e);
    if (blockIdx.x == 0) {
      xscale[r] = pow2i(e);
      xscale[32 + r] = pow2i(-e);
    }
  }
  __syncthreads();
  const long long tiles = (n_rows + T3_ROWS - 1) / T3_ROWS;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long row = t * T3_ROWS + r;
    float x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = 0.f;
    if (row < n_rows) {
      const float* src = obs + row * n_in;
      float rmax = 0.f, probe = 0.f;
#pragma unroll
      for (int c = 0; c < 31; ++c)
        if (c < n_in) {
          x[c] = __ldg(src + c) * s_xs[c];
          rmax = fmaxf(rmax, fabsf(x[c]));
          probe += x[c];
        }
      if (!(rmax <= T2_RANGE) || probe != probe) bad = true;
      // a row whose every feature sits 2^17 below its column's maximum has lost its l-splits
      if (rmax > 0.f && rmax < 0.03125f) bad = true;
      x[31] = 1.0f;
    }
    uint4 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split2h(x[8 * j + 0], x[8 * j + 1], h[j].x, l[j].x);
      split2h(x[8 * j + 2], x[8 * j + 3], h[j].y, l[j].y);
      split2h(x[8 * j + 4], x[8 * j + 5], h[j].z, l[j].z);
      split2h(x[8 * j + 6], x[8 * j + 7], h[j].w, l[j].w);
    }
    uint8_t* dst = ximg + (size_t)t * T2_ACT + (size_t)r * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<uint4*>(dst + ((j ^ (r & 7)) << 4)) = h[j];
      *reinterpret_cast<uint4*>(dst + (((4 + j) ^ (r & 7)) << 4)) = l[j];
    }
  }
  if (bad) *bad_flag = 1.0f;
}

// ------------------------------------------------------------------------------------------------------------------
// the step kernel
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T3_THREADS, 1) mlp_tc3_kernel(const Tc3Args p) {
  extern __shared__ uint8_t smem_raw[];
  // chain 0 = policy, chain 1 = value; which of them run is the same for every thread of the grid
  const bool run_p = p.run_policy != 0 && (p.stop_flag == nullptr || *p.stop_flag == 0);
  const bool run_v = p.run_value != 0;
  if (!run_p && !run_v) return;
  if (*p.x_bad != 0.f) return;  // the packed observations left the fp16 range: the engine redoes the update (wide-range path)
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform (see mlp_tc2.cu)
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* s_bias = reinterpret_cast<float*>(sm + S3_BIAS);
  float* s_dist = reinterpret_cast<float*>(sm + S3_DIST);
  float* s_scale = reinterpret_cast<float*>(sm + S3_SCALE);
  float* s_xs = reinterpret_cast<float*>(sm + S3_XS);
  float* s_red = reinterpret_cast<float*>(sm + S3_RED);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + S3_BARS + 64);
  int* s_bad = reinterpret_cast<int*>(sm + S3_BARS + 68);
  // ready[c] at +8c, chain[c] at +16+8c, xfull[b] at +32+8b, free[c] at +48+8c
  const uint32_t bars = base + S3_BARS;
  const int n_in = p.n_in;
  bool bad = false;
#ifdef B200RL_TC3_TIMING
  unsigned long long tacc[48];
  for (int i = 0; i < 48; ++i) tacc[i] = 0;
  const long long t_kernel0 = clock64();
#endif

  // ---- one-time setup: scales; weights of both networks as fp16 pairs; biases ----
  // Only the weight operands need zero padding (X tiles arrive whole by bulk copy, H and dOut are fully written by
  // their jobs before any MMA reads them).  Both parameter vectors (44 KB) are first brought into the still unused
  // activation buffers with independent, coalesced loads: the two passes below (maxima, then conversion) would
  // otherwise pay an L2 round trip per element, one after the other -- 17 of the 18 us this set-up used to take.
  for (uint32_t i = S3_W / 16 + tid; i < S3_OPERANDS_END / 16; i += T3_THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) *s_bad = 0;
  if (tid < 64) s_xs[tid] = __ldg(p.xscale + tid);
  float* s_par = reinterpret_cast<float*>(sm + S3_H);
  static_assert(2 * S3_CHAIN >= 4 * (2 * 6000 + 64), "parameter staging area");
#pragma unroll 1
  for (int net = 0; net < 2; ++net) {
    const float* par = p.params[net];
    float* dst = s_par + (net == 0 ? 0 : p.P[0]);
#pragma unroll 8
    for (int i = tid; i < p.P[net]; i += T3_THREADS) dst[i] = (*(par + i));
  }
  __syncthreads();
#pragma unroll 1
  for (int net = 0; net < 2; ++net) {
    const Tc3Net& nn = p.net[net];
    const float* par = s_par + (net == 0 ? 0 : p.P[0]);
    float m1 = 0.f, m2 = 0.f, m3 = 0.f;
    for (int idx = tid; idx < nn.h1 * n_in; idx += T3_THREADS) {
      const float w = (*(par + nn.w_off[0] + idx)) * s_xs[32 + idx % n_in];
      m1 = fmaxf(m1, fabsf(w));
      if (w != w) bad = true;
    }
    for (int idx = tid; idx < nn.h2 * nn.h1; idx += T3_THREADS) {
      const float w = (*(par + nn.w_off[1] + idx));
      m2 = fmaxf(m2, fabsf(w));
      if (w != w) bad = true;
    }
    for (int idx = tid; idx < nn.n_out * nn.h2; idx += T3_THREADS) {
      const float w = (*(par + nn.w_off[2] + idx));
      m3 = fmaxf(m3, fabsf(w));
      if (w != w) bad = true;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
      m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, o));
    }
    if (lane == 0) {
      s_red[warp * 8 + 4 * net + 0] = m1;
      s_red[warp * 8 + 4 * net + 1] = m2;
      s_red[warp * 8 + 4 * net + 2] = m3;
    }
  }
  __syncthreads();
  if (bad) *s_bad = 1;  // NaN weight
  bad = false;
  if (tid < 2) {
    const int net = tid;
    const Tc3Net& nn = p.net[net];
    float m1 = 0.f, m2 = 0.f, m3 = 0.f;
    for (int w = 0; w < T3_THREADS / 32; ++w) {
      m1 = fmaxf(m1, s_red[w * 8 + 4 * net + 0]);
      m2 = fmaxf(m2, s_red[w * 8 + 4 * net + 1]);
      m3 = fmaxf(m3, s_red[w * 8 + 4 * net + 2]);
    }
    bool b0 = false;
    const int ew1 = fit_exp(m1, b0), ew2 = fit_exp(m2, b0), ew3 = fit_exp(m3, b0);
    float typ;  // typical magnitude of N * dLoss/dOut: the gradient scale parks it near 2^3
    if (net == 1) {
      const float tm = __ldg(p.target_absmax);
      typ = (tm > 0.f && tm < INFINITY) ? 0.25f * tm : 1.f;
    } else if (p.dist == B200RL_DIST_GAUSSIAN) {
      float smin = INFINITY;
      for (int a = 0; a < nn.n_out; ++a) smin = fminf(smin, expf(__ldg(p.log_std + a)));
      typ = (smin > 0.f && smin < INFINITY) ? 1.f / smin : 1.f;
    } else {
      typ = 0.5f;
    }
    int eg = 3 + ilogbf(p.n_glob_f) - ilogbf(typ);
    eg = eg < -100 ? -100 : (eg > 100 ? 100 : eg);
    float* sc = s_scale + 16 * net;
    sc[C3_G] = pow2i(eg);
    sc[C3_U1] = pow2i(-ew1);
    sc[C3_U2] = pow2i(-(T2_H_EXP + ew2));
    sc[C3_U3] = pow2i(-(T2_H_EXP + ew3));
    sc[C3_UH2] = pow2i(-ew3);
    sc[C3_UH1] = pow2i(-ew2);
    sc[C3_W1] = pow2i(ew1);
    sc[C3_W2] = pow2i(ew2);
    sc[C3_W3] = pow2i(ew3);
    sc[C3_OW3] = pow2i(-(T2_H_EXP + eg));
    sc[C3_OW2] = pow2i(-(T2_H_EXP + eg));
    sc[C3_OW1] = pow2i(-eg);  // times 2^-ex_k of the column, applied when the accumulator is read
    sc[C3_OB] = pow2i(-eg);
    if (b0) *s_bad = 1;
  }
  __syncthreads();
#pragma unroll 1
  for (int net = 0; net < 2; ++net) {
    const Tc3Net& nn = p.net[net];
    const float* par = s_par + (net == 0 ? 0 : p.P[0]);
    const uint32_t wb = S3_W + net * S3_WNET;
    auto put = [&](uint32_t buf, uint32_t stride, int r, int c, float x) {
      const __half hb = __float2half_rn(x);
      const __half lb = __float2half_rn(x - __half2float(hb));
      const uint32_t off = buf + (uint32_t)r * 128u + ((uint32_t)((c >> 3) ^ (r & 7)) << 4) + ((uint32_t)(c & 7) << 1);
      *reinterpret_cast<__half*>(sm + off) = hb;
      *reinterpret_cast<__half*>(sm + off + stride) = lb;
    };
    const float* sc = s_scale + 16 * net;
    const float sw1 = sc[C3_W1], sw2 = sc[C3_W2], sw3 = sc[C3_W3];
    for (int idx = tid; idx < nn.h1 * n_in; idx += T3_THREADS)  // W1 transposed: row = input, column = output
      put(wb, T3_W1T, idx % n_in, idx / n_in, ((*(par + nn.w_off[0] + idx)) * s_xs[32 + idx % n_in]) * sw1);
    for (int idx = tid; idx < nn.h2 * nn.h1; idx += T3_THREADS)
      put(wb + 2 * T3_W1T, T3_W2, idx / nn.h1, idx % nn.h1, (*(par + nn.w_off[1] + idx)) * sw2);
    for (int idx = tid; idx < nn.n_out * nn.h2; idx += T3_THREADS)
      put(wb + 2 * T3_W1T + 2 * T3_W2, T3_W3, idx / nn.h2, idx % nn.h2, (*(par + nn.w_off[2] + idx)) * sw3);
    float* bb = s_bias + 160 * net;
    for (int i = tid; i < 64; i += T3_THREADS) {
      bb[i] = i < nn.h1 ? (*(par + nn.b_off[0] + i)) : 0.f;
      bb[64 + i] = i < nn.h2 ? (*(par + nn.b_off[1] + i)) : 0.f;
      if (!(fabsf(bb[i]) < INFINITY) || !(fabsf(bb[64 + i]) < INFINITY)) bad = true;
    }
    for (int i = tid; i < 16; i += T3_THREADS) {
      bb[128 + i] = i < nn.n_out ? (*(par + nn.b_off[2] + i)) : 0.f;
      if (!(fabsf(bb[128 + i]) < INFINITY)) bad = true;
    }
  }
  if (p.dist == B200RL_DIST_GAUSSIAN)
    for (int a = tid; a < p.net[0].n_out; a += T3_THREADS) {
      const float scale = expf(__ldg(p.log_std + a));  // gaussian_policy.py:34
      s_dist[a] = scale * scale;
      s_dist[16 + a] = logf(scale);
      s_dist[32 + a] = 1.f / (2.f * (scale * scale));
      s_dist[48 + a] = 1.f / (scale * scale);
    }
  if (bad) *s_bad = 1;  // non-finite bias
  bad = false;
  if (warp == T3_EPI_WARPS) {
    tmem_alloc(smem_u32(s_tmem), 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      mbar_init(bars + 8 * c, T3_EPI_THREADS);  // ready[c]: every epilogue thread arrives once per job of chain c
      mbar_init(bars + 16 + 8 * c, 1);          // chain[c]: tcgen05.commit
      mbar_init(bars + 32 + 8 * c, 1);          // xfull[b]: arrive.expect_tx by the MMA warp + the copy's bytes
      mbar_init(bars + 48 + 8 * c, 1);          // (spare)
    }
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  const long long num_tiles = (p.n_rows + T3_ROWS - 1) / T3_ROWS;
  const long long cta_tiles = (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // tiles blockIdx.x + k * gridDim.x
  const int c_first = run_p ? 0 : 1, c_last = run_v ? 1 : 0;

  if (warp == T3_EPI_WARPS) {
    // =============================== MMA issuer (and bulk-copy producer) warp =============================
    constexpr uint32_t I_128_64_KK = make_idesc_f16(128, 64, 0, 0), I_128_16_KK = make_idesc_f16(128, 16, 0, 0),
                       I_128_64_KM = make_idesc_f16(128, 64, 0, 1), I_128_64_MM = make_idesc_f16(128, 64, 1, 1),
                       I_128_32_MM = make_idesc_f16(128, 32, 1, 1), I_128_16_MM = make_idesc_f16(128, 16, 1, 1);
    (void)I_128_16_MM;
    const uint32_t ub = base, ubar = bars;
    if (tmem != 0u) __trap();  // a 512-column allocation is the whole tensor memory: base 0 (uniform by construction)
    // views at chain 0 / net 0 / X buffer 0; the others are reached by adding byte offsets to the descriptors
    const Op2 X_K = op2_kmajor(ub + S3_XB, 64);                     // A: X, h at +0, l at +64 bytes
    const Op2 X_M = op2_mnmajor(ub + S3_XB, T2_ACT, 64);            // B, N = 64: features h 0..31 (col 31 = ones) | l
    const Op2 X_M16 = op2_mnmajor(ub + S3_XB + 32, T2_ACT, 64);     // B, N = 16: h cols 16..31 (col 31 = ones)
    const Op2 DO_K = op2_kmajor(ub + S3_XB, 32);                    // A, K = 16: dOut h at +0, l at +32 bytes
    const Op2 DO_M = op2_mnmajor(ub + S3_XB, T2_ACT, 32);           // B, N = 32: dOut h | l
    const Op2 H1_K = op2_kmajor(ub + S3_H, T2_ACT), H2_K = op2_kmajor(ub + S3_H + 2 * T2_ACT, T2_ACT);
    const Op2 H1_M = op2_mnmajor(ub + S3_H, T2_ACT, T2_ACT), H2_M = op2_mnmajor(ub + S3_H + 2 * T2_ACT, T2_ACT, T2_ACT);
    const Op2 W1T_M = op2_mnmajor(ub + S3_W, 32 * 128, T3_W1T);
    const Op2 W2_K = op2_kmajor(ub + S3_W + 2 * T3_W1T, T3_W2), W2_M = op2_mnmajor(ub + S3_W + 2 * T3_W1T, 64 * 128, T3_W2);
    const Op2 W3_K = op2_kmajor(ub + S3_W + 2 * T3_W1T + 2 * T3_W2, T3_W3),
              W3_M = op2_mnmajor(ub + S3_W + 2 * T3_W1T + 2 * T3_W2, 16 * 128, T3_W3);
    uint32_t accmask = 0u;  // bit (3 c + j): the accumulator j of chain c holds data (the first product overwrites it)
    uint32_t par_ready = 0u;  // bit c: phase parity of ready[c]
    auto load_x = [&](long long k) {  // tile k of this CTA -> X buffer k & 1
      const uint32_t b = (uint32_t)(k & 1);
      const long long tile = blockIdx.x + k * gridDim.x;
      uint32_t e;
      asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(e));
      if (e) {
        mbar_arrive_expect_tx(ubar + 32 + 8 * b, T2_ACT);
        bulk_copy_g2s(ub + S3_XB + b * T2_ACT, p.ximg + (size_t)tile * T2_ACT, T2_ACT, ubar + 32 + 8 * b);
      }
      __syncwarp();
    };
    auto issue_z1 = [&](const int c, long long k) {  // Z1 = X W1^T of tile k for chain c
      const uint32_t b = (uint32_t)(k & 1);
#ifdef B200RL_TC3_TIMING
      const long long xt0 = clock64();
#endif
      mbar_wait(ubar + 32 + 8 * b, (uint32_t)((k >> 1) & 1));
      tc_fence_after_sync();
#ifdef B200RL_TC3_TIMING
      tacc[40] += (unsigned long long)(clock64() - xt0);
#endif
      issue_chain3<2>(c * M3_CHAIN + M3_Z, I_128_64_KM, op2_at(X_K, b * T2_ACT), op2_at(W1T_M, c * S3_WNET));
    };
    if (cta_tiles > 0) {
      load_x(0);
#pragma unroll 1
      for (int c = c_first; c <= c_last; ++c) {
        issue_z1(c, 0);
        umma_commit_elect(ubar + 16 + 8 * c);
      }
    }
#pragma unroll 1
    for (long long k = 0; k < cta_tiles; ++k) {
      const uint32_t xo = (uint32_t)(k & 1) * T2_ACT, dob = (uint32_t)((k + 1) & 1) * T2_ACT;
#pragma unroll 1
      for (int stage = 1; stage <= 5; ++stage) {
#pragma unroll 1
        for (int c = c_first; c <= c_last; ++c) {
          const uint32_t co = c * S3_CHAIN, wo = c * S3_WNET;
          const uint32_t tz = c * M3_CHAIN, ta = M3_ACC + c * M3_ACC_NET;
#ifdef B200RL_TC3_TIMING
          const long long it0 = clock64();
#endif
          mbar_wait(ubar + 8 * c, (par_ready >> c) & 1u);  // every epilogue thread has delivered the stage inputs
          par_ready ^= 1u << c;
          tc_fence_after_sync();
#ifdef B200RL_TC3_TIMING
          const long long it1 = clock64();
          tacc[20 + 2 * (stage - 1) + c] += (unsigned long long)(it1 - it0);
#endif
          if (stage == 1) {  // Z2 = H1 W2^T
            issue_chain3<4>(tz + M3_Z, I_128_64_KK, op2_at(H1_K, co), op2_at(W2_K, wo));
          } else if (stage == 2) {  // OUT = H2 W3^T
            issue_chain3<4>(tz + M3_OUT, I_128_16_KK, op2_at(H2_K, co), op2_at(W3_K, wo));
          } else if (stage == 3) {
            // dH2 = dOut W3 ; dW3^T[i][o] += sum_r H2[r][i] dOut[r][o] (must retire before H2 becomes dZ2 in place)
            issue_chain3<1>(tz + M3_Z, I_128_64_KM, op2_at(DO_K, dob + c * 64), op2_at(W3_M, wo));
            issue_stacked<8, 1>(ta + M3_DW3, I_128_32_MM, (accmask >> (3 * c)) & 1u, op2_at(H2_M, co),
                                op2_at(DO_M, dob + c * 64));  // N = 32: dOut h | l
            accmask |= 1u << (3 * c);
          } else if (stage == 4) {
            // both chains' stage-3 products have retired (their E4 jobs waited for them before arriving here): the
            // dOut buffer is free -> bring the NEXT tile's observations into it
            if (c == c_last && k + 1 < cta_tiles) load_x(k + 1);
            // dH1 = dZ2 W2 ; dW2[o][i] += sum_r dZ2[r][o] H1[r][i] ; db2[o] += sum_r dZ2[r][o] * 1 (ones column of X)
            issue_chain3<4>(tz + M3_Z, I_128_64_KM, op2_at(H2_K, co), op2_at(W2_M, wo));
            issue_stacked<8, 2>(ta + M3_DW2, I_128_64_MM, (accmask >> (3 * c + 1)) & 1u, op2_at(H2_M, co), op2_at(H1_M, co));
            issue_stacked<8, 1>(ta + M3_DB2, I_128_16_MM, (accmask >> (3 * c + 1)) & 1u, op2_at(H2_M, co), op2_at(X_M16, xo));
            accmask |= 1u << (3 * c + 1);
          } else {
            // dW1[o][i] += sum_r dZ1[r][o] X[r][i]; column 31 (ones) collects db1.  Then the next tile's Z1.
            issue_stacked<8, 1>(ta + M3_DW1, I_128_64_MM, (accmask >> (3 * c + 2)) & 1u, op2_at(H1_M, co), op2_at(X_M, xo));
            accmask |= 1u << (3 * c + 2);
            if (k + 1 < cta_tiles) issue_z1(c, k + 1);
          }
          umma_commit_elect(ubar + 16 + 8 * c);
          __syncwarp();
#ifdef B200RL_TC3_TIMING
          tacc[30 + 2 * (stage - 1) + c] += (unsigned long long)(clock64() - it1);
#endif
        }
      }
    }
#ifdef B200RL_TC3_TIMING
    if (lane == 0 && blockIdx.x == 0) {
      for (int i = 20; i <= 40; ++i) g_tc3_t[i] = tacc[i];
      g_tc3_t[41] = (unsigned long long)cta_tiles;
    }
#endif
  } else {
    // =============================== epilogue warps: one pool of 16 ===================================
    // Jobs run in the fixed order (policy, E1) (value, E1) (policy, E2) ... (value, E5) | next tile, all 16 warps on
    // one job at a time (16 columns each), so the MMAs a job hands over run under the other chain's next job.
    const int q = warp & 3, part = warp >> 2;
    const int r = 32 * q + lane;                          // row of the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const int cs = 16 * part;
    uint32_t ph_chain = 0u;  // bit c: phase parity of chain[c]
    const float sH = pow2i(T2_H_EXP), hh = pow2i(-2 * T2_H_EXP);
    const int A_out = p.net[0].n_out;
    double sc[5] = {0, 0, 0, 0, 0};  // policy: loss terms, old_logp - logp, entropy, logp, logp^2
    double vs = 0.0;                 // value: squared errors
    int rows_done = 0;
    float db3[15], db3v = 0.f;
#pragma unroll
    for (int a = 0; a < 15; ++a) db3[a] = 0.f;
    float adv_mean = 0.f, adv_inv_std = 1.f;  // normalize_tensor (utils.py:90-92): mean, UNBIASED std, no epsilon
    if (p.adv_stats != nullptr) {
      const double s1 = p.adv_stats[0], s2 = p.adv_stats[1], cnt = p.adv_stats[2];
      const double mean = s1 / cnt;
      adv_mean = (float)mean;
      adv_inv_std = 1.f / (float)sqrt((s2 - cnt * mean * mean) / (cnt - 1.0));
    }

#ifdef B200RL_TC3_TIMING
    long long t_work0 = 0;
#endif
    auto job = [&](const int c, const int stage, const long long k) {
      const uint32_t tz = tmem + lane_addr + (uint32_t)c * M3_CHAIN;
      const uint32_t so = S3_H + (uint32_t)c * S3_CHAIN;
      const uint32_t bar_ready = bars + 8 * c, bar_chain = bars + 16 + 8 * c;
      const float* scl = s_scale + 16 * c;
      const float* bias = s_bias + 160 * c;
      const long long tile = blockIdx.x + k * gridDim.x;
      const long long row = tile * T3_ROWS + r;
      const bool valid = row < p.n_rows;
      const bool loss_warp = part == (int)((k + 2 * c) & 3);  // rotates; the two chains use different warps
      auto arrive = [&]() {
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(bar_ready);
      };
      auto wait_chain = [&]() {
#ifdef B200RL_TC3_TIMING
        const long long w0 = clock64();
#endif
        mbar_wait(bar_chain, (ph_chain >> c) & 1u);
        ph_chain ^= 1u << c;
        tc_fence_after_sync();
#ifdef B200RL_TC3_TIMING
        {
          const long long w1 = clock64();
          tacc[2 * (stage - 1) + c] += (unsigned long long)(w1 - w0);
          t_work0 = w1;
        }
#endif
      };
      if (stage == 1 || stage == 2) {
        // ---- E1 / E2: Z (TMEM) * unscale + bias -> tanh -> fp16 pairs ----
        wait_chain();
        const float unscale = scl[stage == 1 ? C3_U1 : C3_U2];
        const float* bs = bias + (stage == 1 ? 0 : 64);
        const uint32_t dst = so + (stage == 1 ? 0u : 2 * T2_ACT);
        uint32_t v[16];
        tmem_ld16(tz + M3_Z + cs, v);
        tmem_wait_ld();
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = fmaf(__uint_as_float(v[j]), unscale, bs[cs + j]);
        tanh16_scaled(z, sH);  // tanh(z) * 2^14
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = z[8 * ch + j];
          store_chunk2(sm, dst, r, (cs >> 3) + ch, x);
        }
        arrive();
      } else if (stage == 3) {
        // ---- E3: loss epilogue, one row per thread, on this tile's loss warps of this chain ----
        const uint32_t dob = S3_XB + (uint32_t)((k + 1) & 1) * T2_ACT;  // the X buffer of the other parity
        float x0[8], x1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x0[j] = x1[j] = 0.f;
        if (c == 0) {
          float pf_act[15], pf_adv = 0.f, pf_old = 0.f;
#pragma unroll
          for (int a = 0; a < 15; ++a) pf_act[a] = 0.f;
          if (loss_warp && valid) {  // issue the loads before waiting for OUT
            if (p.dist == B200RL_DIST_GAUSSIAN) {
#pragma unroll
              for (int a = 0; a < 15; ++a)
                if (a < A_out) pf_act[a] = __ldg(p.actions + row * A_out + a);
            } else {
              pf_act[0] = __ldg(p.actions + row);
            }
            pf_adv = __ldg(p.adv_raw + row);
            pf_old = __ldg(p.old_logp + row);
          }
          wait_chain();
          if (loss_warp) {
            uint32_t o[16];
            tmem_ld16(tz + M3_OUT, o);
            tmem_wait_ld();
            float out[16], dout[16];
            const float u3 = scl[C3_U3];
#pragma unroll
            for (int a = 0; a < 16; ++a) {
              out[a] = fmaf(__uint_as_float(o[a]), u3, bias[128 + a]);
              dout[a] = 0.f;
            }
            if (valid) {
              float lp = 0.f, ent = 0.f, dlp[16];
#pragma unroll
              for (int a = 0; a < 16; ++a) dlp[a] = 0.f;
              if (p.dist == B200RL_DIST_GAUSSIAN) {
#pragma unroll
                for (int a = 0; a < 15; ++a)
                  if (a < A_out) {
                    const float lsc = s_dist[16 + a];
                    const float d = pf_act[a] - out[a];
                    lp += -(d * d) * s_dist[32 + a] - lsc - T3_LOG_SQRT_2PI;  // torch Normal.log_prob
                    ent += T3_ENT_CONST + lsc;                                // torch Normal.entropy
                    dlp[a] = d * s_dist[48 + a];
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
              float adv = pf_adv;
              if (p.adv_stats != nullptr) adv = (adv - adv_mean) * adv_inv_std;  // utils.py:91
              // ppo.py:245-255
              const float ratio = expf(lp - pf_old);
              const float s1 = ratio * adv;
              const float s2 = fminf(fmaxf(ratio, p.clip_lo), p.clip_hi) * adv;
              const float term = -fminf(s1, s2);
              const bool pass = adv >= 0.f ? (ratio <= p.clip_hi) : (ratio >= p.clip_lo);
              const float coef = pass ? (-p.inv_n * adv) * ratio : 0.f;
#pragma unroll
              for (int a = 0; a < 15; ++a) dout[a] = coef * dlp[a];
              sc[0] += (double)term;
              sc[1] += (double)(pf_old - lp);
              sc[2] += (double)ent;
              sc[3] += (double)lp;
              sc[4] += (double)lp * (double)lp;
              rows_done += 1;
            }
            const float sG = scl[C3_G];
#pragma unroll
            for (int a = 0; a < 15; ++a) db3[a] += dout[a];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              x0[j] = dout[j] * sG;
              x1[j] = j < 7 ? dout[8 + j] * sG : 0.f;
            }
          }
        } else {
          float pf_tgt = 0.f;
          if (loss_warp && valid) pf_tgt = __ldg(p.target + row);
          wait_chain();
          if (loss_warp) {
            uint32_t o[8];
            tmem_ld8(tz + M3_OUT, o);
            tmem_wait_ld();
            if (valid) {  // ppo.py:282-287
              const float vout = fmaf(__uint_as_float(o[0]), scl[C3_U3], bias[128]);
              const float diff = vout - pf_tgt;
              const float dout = (2.f * diff) * p.inv_n;
              vs += (double)(diff * diff);
              db3v += dout;
              x0[0] = dout * scl[C3_G];
            }
          }
        }
        if (loss_warp) {
          if (out_of_range8(x0) || out_of_range8(x1)) bad = true;
          // 16 columns h, then 16 columns l, at fp16 columns 32 c .. 32 c + 31 of the dOut buffer
          uint4 h0, l0, h1, l1;
          split2h(x0[0], x0[1], h0.x, l0.x);
          split2h(x0[2], x0[3], h0.y, l0.y);
          split2h(x0[4], x0[5], h0.z, l0.z);
          split2h(x0[6], x0[7], h0.w, l0.w);
          split2h(x1[0], x1[1], h1.x, l1.x);
          split2h(x1[2], x1[3], h1.y, l1.y);
          split2h(x1[4], x1[5], h1.z, l1.z);
          split2h(x1[6], x1[7], h1.w, l1.w);
          uint8_t* rowp = sm + dob + (uint32_t)r * 128u;
          const uint32_t sw = (uint32_t)(r & 7), c4 = 4u * (uint32_t)c;
          *reinterpret_cast<uint4*>(rowp + (((c4 + 0u) ^ sw) << 4)) = h0;
          *reinterpret_cast<uint4*>(rowp + (((c4 + 1u) ^ sw) << 4)) = h1;
          *reinterpret_cast<uint4*>(rowp + (((c4 + 2u) ^ sw) << 4)) = l0;
          *reinterpret_cast<uint4*>(rowp + (((c4 + 3u) ^ sw) << 4)) = l1;
        }
        arrive();
      } else {
        // ---- E4 / E5: dZ (scaled) = dH_acc * unscale * (1 - H^2), H re-read from its fp16 pair, written in place ----
        // (loading this thread's H chunks BEFORE the wait -- it wrote them itself two jobs ago -- to hide the
        // shared-memory latency under the barrier: 0.551 vs 0.525 ms, the 16 extra live registers spill at the 96 cap;
        // a second commit right behind the dH product, so that this job computes while the weight-gradient products
        // still read H and only its stores wait for them, measured 2 % SLOWER: 0.5515 vs 0.540 ms.  With `setmaxnreg`
        // -- a full fifth warpgroup of 128 threads whose idle warps hand their registers over: 104 per epilogue thread
        // and 64 for the MMA warp, or 112 / 32; registers are conserved inside the CTA, 640 x 96 = 128 AUX + 512 EPI --
        // the spills go away and the prefetch is STILL slower (0.540 vs 0.522 ms): early shared-memory reads compete
        // with the tensor pipe's operand fetches, which are what the waiting job is waiting for.  All reverted.)
        wait_chain();  // dH (and the weight-gradient products that still read H)
        // (1 - H^2) * 2^28 = fma(-Hs, Hs, 2^28) with Hs = H * 2^14 as stored; 2^-28 is folded into the unscale factor
        const float unscale = scl[stage == 4 ? C3_UH2 : C3_UH1] * hh;
        const uint32_t buf = so + (stage == 4 ? 2 * T2_ACT : 0u);
        const float one28 = 268435456.f;
        uint32_t g[16];
        tmem_ld16(tz + M3_Z + cs, g);
        tmem_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float x[8];
          load_chunk2(sm, buf, r, (cs >> 3) + ch, x);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = (__uint_as_float(g[8 * ch + j]) * unscale) * fmaf(-x[j], x[j], one28);
          if (too_large8(x)) bad = true;
          store_chunk2(sm, buf, r, (cs >> 3) + ch, x);
        }
        arrive();
      }
    };

#ifdef B200RL_TC3_TIMING
    const long long t_loop0 = clock64();
#endif
#pragma unroll 1
    for (long long k = 0; k < cta_tiles; ++k) {
#pragma unroll 1
      for (int stage = 1; stage <= 5; ++stage) {
#pragma unroll 1
        for (int c = c_first; c <= c_last; ++c) {
          job(c, stage, k);  // one copy of the job code (instruction cache)
#ifdef B200RL_TC3_TIMING
          tacc[10 + 2 * (stage - 1) + c] += (unsigned long long)(clock64() - t_work0);
#endif
        }
      }
    }

#ifdef B200RL_TC3_TIMING
    const long long t_loop_end = clock64();
#endif
    // ---- per-CTA results ----
    if (cta_tiles > 0) {
#pragma unroll 1
      for (int c = c_first; c <= c_last; ++c) {  // the last dW1 of each chain
        mbar_wait(bars + 16 + 8 * c, (ph_chain >> c) & 1u);
        tc_fence_after_sync();
      }
    }
    tc_fence_before_sync();
    asm volatile("bar.sync 2, %0;" ::"n"(T3_EPI_THREADS) : "memory");  // every MMA of the CTA has retired
    tc_fence_after_sync();
    {
      // stacked accumulators: lanes 0..63 = h-split half (partial row 2b), lanes 64..127 = l-split half (row 2b + 1);
      // 8 jobs per net (dW2 x 4 column blocks, dW1 x 2, dW3, db2); warp `part` takes jobs part, part + 4 of both nets
      float* dst_row = p.partials + ((size_t)blockIdx.x * 2 + (q >> 1)) * (size_t)(p.P[0] + p.P[1]);
      const int m = 32 * (q & 1) + lane;  // feature index
      const uint32_t ta = tmem + lane_addr + M3_ACC;
      uint32_t v[16], w[16];
#pragma unroll 1
      for (int c = c_first; c <= c_last; ++c) {
        const Tc3Net& nn = p.net[c];
        float* dst = dst_row + (c == 0 ? 0 : p.P[0]);
        const float* scl = s_scale + 16 * c;
        const bool have = cta_tiles > 0;
        const uint32_t tn = ta + c * M3_ACC_NET;
#pragma unroll 1
        for (int jb = part; jb < 8; jb += 4) {
          // columns of this job, and of its second half where the operand's l columns went to their own block
          const uint32_t col = jb < 4 ? M3_DW2 + 16 * jb : (jb < 6 ? M3_DW1 + 16 * (jb - 4) : (jb == 6 ? M3_DW3 : M3_DB2));
          const uint32_t col2 = jb < 4 ? col : (jb < 6 ? col + 32 : (jb == 6 ? col + 16 : col));
          if (have) {
            tmem_ld16(tn + col, v);
            tmem_ld16(tn + col2, w);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = w[j] = 0u;  // a CTA without tiles: tensor memory was never written
          }
          if (jb < 4) {  // dW2 [h2 o][h1 i]: columns 16 jb .. +15
            const float u = scl[C3_OW2];
            if (m < nn.h2)
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (16 * jb + j < nn.h1) dst[nn.w_off[1] + m * nn.h1 + 16 * jb + j] = __uint_as_float(v[j]) * u;
          } else if (jb < 6) {  // dW1 [h1 o][n_in i] in columns 0..30, db1 in column 31
            const int c0 = 16 * (jb - 4);
            const float u = scl[C3_OW1];
            if (m < nn.h1) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < n_in)
                  dst[nn.w_off[0] + m * n_in + c0 + j] =
                      ((__uint_as_float(v[j]) + __uint_as_float(w[j])) * u) * s_xs[32 + c0 + j];
              if (jb == 5) dst[nn.b_off[0] + m] = (__uint_as_float(v[15]) + __uint_as_float(w[15])) * scl[C3_OB];
            }
          } else if (jb == 6) {  // dW3^T [h2 i][16 o]
            const float u = scl[C3_OW3];
            if (m < nn.h2)
#pragma unroll
              for (int a = 0; a < 15; ++a)
                if (a < nn.n_out) dst[nn.w_off[2] + a * nn.h2 + m] = (__uint_as_float(v[a]) + __uint_as_float(w[a])) * u;
          } else {  // db2 (column 15 = sum_r dZ2[r][o] * ones)
            if (m < nn.h2) dst[nn.b_off[1] + m] = __uint_as_float(v[15]) * scl[C3_OB];
          }
        }
      }
    }
    // per-thread sums -> per-warp sums (tree) -> the 16 warps in order: fixed order => reproducible.  The scratch
    // aliases the X buffers (idle now).
    float* e_db3 = reinterpret_cast<float*>(sm + S3_XB + S3_END_DB3);
    double* e_sc = reinterpret_cast<double*>(sm + S3_XB + S3_END_SC);
#pragma unroll
    for (int a = 0; a < 15; ++a) {
      float t = db3[a];
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o2);
      if (lane == 0) e_db3[warp * 16 + a] = t;
    }
    {
      float t = db3v;
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o2);
      if (lane == 0) e_db3[warp * 16 + 15] = t;
    }
#pragma unroll
    for (int kk = 0; kk < 5; ++kk) {
      const double t = warp_sum(sc[kk]);
      if (lane == 0) e_sc[warp * 8 + kk] = t;
    }
    {
      const double t = warp_sum((double)rows_done);
      if (lane == 0) e_sc[warp * 8 + 5] = t;
      const double t2 = warp_sum(vs);
      if (lane == 0) e_sc[warp * 8 + 6] = t2;
    }
    asm volatile("bar.sync 2, %0;" ::"n"(T3_EPI_THREADS) : "memory");
    const size_t Ptot = (size_t)(p.P[0] + p.P[1]);
    if (tid < 16) {  // b3 gradients: the 16 per-warp totals in warp order
      const int c = tid == 15 ? 1 : 0, a = tid == 15 ? 0 : tid;
      const bool runs = c == 0 ? run_p : run_v;
      if (runs && a < p.net[c].n_out) {
        float t = 0.f;
        for (int w = 0; w < T3_EPI_WARPS; ++w) t += e_db3[w * 16 + tid];
        const size_t off = (c == 0 ? 0 : (size_t)p.P[0]) + p.net[c].b_off[2] + a;
        p.partials[((size_t)blockIdx.x * 2) * Ptot + off] = t;
        p.partials[((size_t)blockIdx.x * 2 + 1) * Ptot + off] = 0.f;
      }
    }
    if (tid >= 32 && tid < 32 + 2 * B200RL_N_SCALARS) {  // scalar sums: policy 0..7, value 8..15
      const int s = tid - 32;
      double t = 0.0;
      if (s < 6) {
        for (int w = 0; w < T3_EPI_WARPS; ++w) t += e_sc[w * 8 + s];
      } else if (s == 8) {
        for (int w = 0; w < T3_EPI_WARPS; ++w) t += e_sc[w * 8 + 6];
      }
      p.scalar_partials[((size_t)blockIdx.x * 2) * (2 * B200RL_N_SCALARS) + s] = t;
      p.scalar_partials[((size_t)blockIdx.x * 2 + 1) * (2 * B200RL_N_SCALARS) + s] = 0.0;
    }
    if (bad) *s_bad = 1;
#ifdef B200RL_TC3_TIMING
    if (tid == 0 && blockIdx.x == 0) {
      for (int i = 0; i < 20; ++i) g_tc3_t[i] = tacc[i];
      g_tc3_t[42] = (unsigned long long)(t_loop0 - t_kernel0);   // setup
      g_tc3_t[43] = (unsigned long long)(t_loop_end - t_loop0);  // tile loop
      g_tc3_t[44] = (unsigned long long)(clock64() - t_loop_end);  // read-out
    }
#endif
  }

  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0 && *s_bad != 0) *p.status = 1.0f;  // sticky: the engine redoes the update on the wide-range path
  if (warp == T3_EPI_WARPS) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// reduce_adam3: fixed-order reduction of the per-CTA partial rows of BOTH networks (as b200rl_reduce_partials) and
// torch.optim.Adam's single-tensor update of both parameter vectors (as b200rl_adam_step), early-stop test included
// (ppo.py:176-181), in ONE launch.  mode 0: reduce + Adam (single GPU); mode 1: reduce only, scalars appended to the
// flat gradient as float32 (the buffer ONE all-reduce carries); mode 2: Adam only, scalars read back from that tail.
// A block owns 32 parameters; every block derives the stop decision from the same inputs in the same order.
//
// Data-parallel runs on one node replace "mode 1, NCCL all-reduce, mode 2" by a one-shot exchange over peer-mapped
// memory (NVLink / NVSwitch): mode 3 reduces and stores this rank's [gradient | scalars] into its exchange buffer,
// the last block to finish publishes the sequence number (release, system scope); mode 4 waits for every rank's
// sequence number (acquire; wait_peers_kernel, one warp), reads ALL ranks' buffers -- its own included -- and adds them in rank order, so every
// rank applies bit-identical updates, then runs Adam.  22 KB per rank: latency bound, ~2 us per peer read.
// Buffers alternate between two parities: a rank can only be one exchange ahead of its slowest peer (it needs that
// peer's next sequence number to finish its own), so the buffer it overwrites was read by everybody.
// mode 5 = modes 3 and 4 in ONE launch (the blocks wait for the sequence numbers themselves): a data-parallel iteration
// is then two launches, like a single-GPU one.  It needs every block resident at once (ra3_one_wave).
// ------------------------------------------------------------------------------------------------------------------
constexpr int RA3_WARPS = 8;

// One float from every rank's exchange buffer.  The loads are issued back to back (a peer read over NVLink is ~2 us:
// a loop with one load per trip would pay that once per rank) and added in rank order, the same sum on every rank.
template <typename Acc>
__device__ __forceinline__ Acc ra3_gather(float* const* peers, int world, long long offset) {
  float x[RA3_MAX_WORLD];
#pragma unroll
  for (int r = 0; r < RA3_MAX_WORLD; ++r) {
    x[r] = 0.f;
    if (r < world) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x[r]) : "l"(peers[r] + offset) : "memory");
  }
  Acc t = (Acc)0;
#pragma unroll
  for (int r = 0; r < RA3_MAX_WORLD; ++r)
    if (r < world) t += (Acc)x[r];
  return t;
}

__global__ void __launch_bounds__(RA3_WARPS * 32, 3) reduce_adam3_kernel(const Ra3Args a) {
  __shared__ float part[RA3_WARPS][32];
  __shared__ double spart[RA3_WARPS * 4][2 * B200RL_N_SCALARS];
  __shared__ double s_scal[2 * B200RL_N_SCALARS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long Ptot = a.P[0] + a.P[1];
  const long long pidx = (long long)blockIdx.x * 32 + lane;
  const bool stopped_before = a.stop_flag != nullptr && *a.stop_flag != 0;
  const bool run_p = a.run_policy != 0 && !stopped_before, run_v = a.run_value != 0;
  const unsigned parity = a.seq & 1u;
  if (a.mode == 4) {  // every rank's buffer of this exchange has arrived (wait_peers_kernel ran before this launch)
    if (threadIdx.x < 2 * B200RL_N_SCALARS)
      s_scal[threadIdx.x] = ra3_gather<double>(a.peers, a.world, parity * a.xchg_stride + Ptot + threadIdx.x);
  } else if (a.mode != 2) {
    // scalar sums: 32 row classes x 16 scalars, then the classes in order (same in every block)
    const int k = lane & 15, cls = warp * 4 + (lane >> 4) * 2;  // two classes per half-warp pass
    for (int half = 0; half < 2; ++half) {
      double s = 0.0;
      for (int c = cls + half; c < a.rows; c += RA3_WARPS * 4) s += a.scalar_partials[(size_t)c * (2 * B200RL_N_SCALARS) + k];
      spart[cls + half][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 2 * B200RL_N_SCALARS) {
      double t = 0.0;
      for (int c = 0; c < RA3_WARPS * 4; ++c) t += spart[c][threadIdx.x];
      s_scal[threadIdx.x] = t;
    }
  } else {
    if (threadIdx.x < 2 * B200RL_N_SCALARS) s_scal[threadIdx.x] = (double)a.grad[Ptot + threadIdx.x];
  }
  __syncthreads();
  float g = 0.f;
  if (a.mode == 4) {
    if (warp == 0 && pidx < Ptot) {
      g = ra3_gather<float>(a.peers, a.world, parity * a.xchg_stride + pidx);
      if (a.grad != nullptr) a.grad[pidx] = g;
    }
  } else if (a.mode != 2) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (pidx < Ptot) {
      const float* qp = a.partials + pidx;
      int c = warp;
      for (; c + 3 * RA3_WARPS < a.rows; c += 4 * RA3_WARPS) {
        s0 += qp[(size_t)c * Ptot];
        s1 += qp[(size_t)(c + RA3_WARPS) * Ptot];
        s2 += qp[(size_t)(c + 2 * RA3_WARPS) * Ptot];
        s3 += qp[(size_t)(c + 3 * RA3_WARPS) * Ptot];
      }
      for (; c < a.rows; c += RA3_WARPS) s0 += qp[(size_t)c * Ptot];
    }
    part[warp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && pidx < Ptot) {
      g = part[0][lane];
#pragma unroll
      for (int w = 1; w < RA3_WARPS; ++w) g += part[w][lane];
      if (a.grad != nullptr) a.grad[pidx] = g;
    }
  } else if (warp == 0 && pidx < Ptot) {
    g = a.grad[pidx];
  }
  if (a.mode == 1) {
    if (blockIdx.x == 0 && threadIdx.x < 2 * B200RL_N_SCALARS) a.grad[Ptot + threadIdx.x] = (float)s_scal[threadIdx.x];
    return;
  }
  if (a.mode == 3 || a.mode == 5) {
    float* mine = a.peers[a.rank] + parity * a.xchg_stride;
    if (warp == 0 && pidx < Ptot) mine[pidx] = g;
    if (blockIdx.x == 0 && threadIdx.x < 2 * B200RL_N_SCALARS) mine[Ptot + threadIdx.x] = (float)s_scal[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(a.done_counter, 1u);
      if (prev == gridDim.x - 1) {  // every block's part of the buffer is written: publish
        *a.done_counter = 0u;
        __threadfence_system();
        unsigned* flag = reinterpret_cast<unsigned*>(a.peers[a.rank] + 2 * a.xchg_stride) + parity;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(a.seq) : "memory");
      }
    }
    if (a.mode == 3) return;
    // mode 5: the same launch goes on to gather.  Lane r of warp 0 waits for rank r's sequence number -- this rank's own
    // included, which is what tells a block that the OTHER blocks of this grid have written their parts (every block is
    // resident: the host checked that the grid fits the device in one wave, so the spinning blocks cannot starve the
    // publishing one).
    if (warp == 0 && lane < a.world) {
      const unsigned* flag = reinterpret_cast<const unsigned*>(a.peers[lane] + 2 * a.xchg_stride) + parity;
      const long long t0 = clock64();
      unsigned seen;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if (seen == a.seq) break;
        if (clock64() - t0 > 20000000000LL) {  // ~10 s: a rank died or never launched
          *a.comm_error = 1;
          break;
        }
        __nanosleep(64);
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * B200RL_N_SCALARS)
      s_scal[threadIdx.x] = ra3_gather<double>(a.peers, a.world, parity * a.xchg_stride + Ptot + threadIdx.x);
    if (warp == 0 && pidx < Ptot) {
      g = ra3_gather<float>(a.peers, a.world, parity * a.xchg_stride + pidx);
      if (a.grad != nullptr) a.grad[pidx] = g;
    }
    __syncthreads();
  }
  // early stop (ppo.py:176-181): the KL carried by this step's forward pass is the KL after the PREVIOUS update
  bool stop = !run_p;
  if (run_p && a.kl_limit_on) stop = (float)(s_scal[1] / a.n_global) > (float)a.kl_limit;
  if (warp == 0 && pidx < Ptot) {
    const int sidx = pidx < a.P[0] ? 0 : 1;
    const bool apply = sidx == 0 ? (run_p && !stop) : run_v;
    if (apply) {
      const Ra3Seg& sg = a.seg[sidx];
      const long long i = sidx == 0 ? pidx : pidx - a.P[0];
      float m = sg.m[i], v = sg.v[i];
      m = m + sg.one_minus_b1 * (g - m);                      // exp_avg.lerp_(grad, 1 - beta1)
      v = v * sg.b2 + sg.one_minus_b2 * (g * g);              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(v) / sg.bc2_sqrt + sg.eps;    // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
      sg.m[i] = m;
      sg.v[i] = v;
      sg.params[i] = sg.params[i] - sg.step_size * (m / denom);  // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    if (run_p) {
      for (int k = 0; k < B200RL_N_SCALARS; ++k) a.slot_p[k] = s_scal[k];
      if (stop) *a.stop_flag = 1;
      else *a.applied_p += 1;
    }
    if (run_v) {
      for (int k = 0; k < B200RL_N_SCALARS; ++k) a.slot_v[k] = s_scal[B200RL_N_SCALARS + k];
      *a.applied_v += 1;
    }
  }
}

// One warp waits until every rank has published the sequence number of this exchange.  A separate, tiny launch: while
// it spins it holds next to nothing of its SM, and the gather + Adam launch behind it needs no polling at all.
__global__ void __launch_bounds__(32) wait_peers_kernel(float* const* peers, int world, long long xchg_stride,
                                                        unsigned seq, int* comm_error) {
  if ((int)threadIdx.x >= world) return;
  const unsigned* flag = reinterpret_cast<const unsigned*>(peers[threadIdx.x] + 2 * xchg_stride) + (seq & 1u);
  const long long t0 = clock64();
  unsigned seen;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
    if (seen != seq && clock64() - t0 > 20000000000LL) {  // ~10 s: a rank died or never launched
      *comm_error = 1;
      break;
    }
  } while (seen != seq);
}

// ---- host side -------------------------------------------------------------------------------------------------
int launch_wait_peers(const Ra3Args& a, cudaStream_t s) {
  wait_peers_kernel<<<1, 32, 0, s>>>(a.peers, a.world, a.xchg_stride, a.seq, a.comm_error);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

int tc3_configure() {
  static const int rc = []() -> int {
    return (int)cudaFuncSetAttribute(mlp_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3_SMEM_BYTES);
  }();
  if (rc != 0) set_error("mlp_tc3: cudaFuncSetAttribute failed (%s)", cudaGetErrorString((cudaError_t)rc));
  return rc;
}

size_t tc3_ximg_bytes(int64_t n_rows) { return (size_t)((n_rows + T3_ROWS - 1) / T3_ROWS) * T2_ACT; }

int tc3_grid(int64_t n_rows) {
  const int64_t tiles = (n_rows + T3_ROWS - 1) / T3_ROWS;
  const int sms = device_sm_count();
  if (sms <= 0) return -1;
  return (int)(tiles < sms ? (tiles < 1 ? 1 : tiles) : sms);
}

bool tc3_shape_ok(const b200rl_mlp_desc& pol, const b200rl_mlp_desc& val) {
  auto ok = [](const b200rl_mlp_desc& d) {
    return d.n_layers == 3 && d.hidden_act == B200RL_ACT_TANH && d.out_act == B200RL_ACT_IDENTITY && d.sizes[0] >= 1 &&
           d.sizes[0] <= 31 && d.sizes[1] >= 1 && d.sizes[1] <= 64 && d.sizes[2] >= 1 && d.sizes[2] <= 64 &&
           d.sizes[3] >= 1 && d.sizes[3] <= 15;
  };
  return ok(pol) && ok(val) && pol.sizes[0] == val.sizes[0] && val.sizes[3] == 1;
}

int launch_pack_obs(const float* obs, int64_t n_rows, int n_in, const float* absmax, uint8_t* ximg, float* xscale,
                    float* bad_flag, cudaStream_t s) {
  const int64_t tiles = (n_rows + T3_ROWS - 1) / T3_ROWS;
  if (tiles <= 0) return 0;
  const int grid = (int)std::min<int64_t>(tiles, 8LL * 148);
  pack_obs_kernel<<<grid, 128, 0, s>>>(obs, n_rows, n_in, absmax, ximg, xscale, bad_flag);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

int launch_mlp_tc3(const Tc3Args& k, cudaStream_t s) {
  if (tc3_configure()) return 1;
  const int grid = tc3_grid(k.n_rows);
  B200RL_REQUIRE(grid > 0, "mlp_tc3: no CUDA device");
  mlp_tc3_kernel<<<grid, T3_THREADS, T3_SMEM_BYTES, s>>>(k);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

// mode 5 spins inside the grid: only when all of its blocks are resident together
bool ra3_one_wave(long long p_total) {
  static const int per_sm = []() -> int {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reduce_adam3_kernel, RA3_WARPS * 32, 0) != cudaSuccess) {
      (void)cudaGetLastError();
      return 0;
    }
    return n;
  }();
  const int sms = device_sm_count();
  return sms > 0 && (p_total + 31) / 32 <= (long long)per_sm * sms;
}

int launch_reduce_adam3(const Ra3Args& a, cudaStream_t s) {
  const long long Ptot = a.P[0] + a.P[1];
  reduce_adam3_kernel<<<(int)((Ptot + 31) / 32), RA3_WARPS * 32, 0, s>>>(a);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace b200rl

#ifdef B200RL_TC3_TIMING
extern "C" int b200rl_debug_tc3_timing(unsigned long long* out48, int reset) {
  if (reset) {
    unsigned long long z[48] = {0};
    return (int)cudaMemcpyToSymbol(b200rl::g_tc3_t, z, sizeof(z));
  }
  return (int)cudaMemcpyFromSymbol(out48, b200rl::g_tc3_t, sizeof(unsigned long long) * 48);
}
#endif
