// Stand-alone kernels behind the reference's public helper functions (rl_replicas/utils.py), for callers that use
// them outside train():
//   discounted_cumulative_sums (utils.py:14-28), gae (:31-44): float64 in, float64 out like scipy.signal.lfilter on
//   the reference's float64 rewards -- the update engine's own scan (gae_scan.cu) carries float64 but stores the
//   float32 casts train() needs, so it cannot serve a caller who wants the float64 vector back;
//   normalize_tensor (:90-92); polyak_average (:47-57).
// These are utilities, not the hot path: one CTA walks the vector from its end in 4096-item chunks (thread-local
// reverse recurrence, warp-shuffle suffix scan of the affine pairs, carry in a register).
#include "common.cuh"

namespace b200rl {

namespace {
constexpr int SC_THREADS = 1024, SC_ITEMS = 4, SC_CHUNK = SC_THREADS * SC_ITEMS;

struct Affine {  // y_first = s + d * y_after
  double s, d;
};
__device__ __forceinline__ Affine combine(const Affine l, const Affine r) { return Affine{l.s + l.d * r.s, l.d * r.d}; }

// MODE 0: x = in0[i].  MODE 1 / 2 (gae): x = r[i] + gamma * v[i+1] - v[i]  (utils.py:41), in0 = rewards, in1 = values.
// MODE 2: float32 values -- numpy evaluates `gamma * values[1:]` in the array's dtype (a Python-float gamma does not
// promote a float32 array), then adds it to the float64 rewards: reproduced operation by operation.
template <int MODE>
__global__ void __launch_bounds__(SC_THREADS) discounted_cumsum_kernel(const double* __restrict__ in0,
                                                                       const void* __restrict__ in1v, long long n,
                                                                       double gamma, double discount,
                                                                       double* __restrict__ out) {
  __shared__ Affine s_warp[32];
  __shared__ double s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double d_item[SC_ITEMS + 1];
  d_item[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= SC_ITEMS; ++j) d_item[j] = d_item[j - 1] * discount;
  double carry = 0.0;  // y just after the current chunk
  for (long long end = n; end > 0; end -= SC_CHUNK) {
    const long long begin = end - SC_CHUNK;  // may be negative on the first (leftmost) chunk
    const long long i0 = begin + (long long)tid * SC_ITEMS;
    double x[SC_ITEMS];
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
      const long long i = i0 + j;
      if (i >= 0 && i < n) {
        if (MODE == 0) {
          x[j] = in0[i];
        } else if (MODE == 1) {
          const double* in1 = static_cast<const double*>(in1v);
          x[j] = (in0[i] + gamma * in1[i + 1]) - in1[i];
        } else {
          const float* in1 = static_cast<const float*>(in1v);
          x[j] = (in0[i] + (double)__fmul_rn((float)gamma, in1[i + 1])) - (double)in1[i];
        }
      } else {
        x[j] = 0.0;  // in front of the vector: contributes nothing to anyone (everything there is discarded)
      }
    }
    // thread aggregate
    Affine a{0.0, d_item[SC_ITEMS]};
#pragma unroll
    for (int j = SC_ITEMS - 1; j >= 0; --j) a.s = x[j] + discount * a.s;
    // exclusive suffix aggregate over the threads to the right: warp level, then across warps
    Affine incl = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double rs = __shfl_down_sync(0xffffffffu, incl.s, o), rd = __shfl_down_sync(0xffffffffu, incl.d, o);
      if (lane + o < 32) incl = combine(incl, Affine{rs, rd});
    }
    if (lane == 0) s_warp[warp] = incl;
    Affine right{__shfl_down_sync(0xffffffffu, incl.s, 1), __shfl_down_sync(0xffffffffu, incl.d, 1)};
    if (lane == 31) right = Affine{0.0, 1.0};
    __syncthreads();
    Affine wr{0.0, 1.0};  // warps to the right, nearest first
    for (int w = warp + 1; w < SC_THREADS / 32; ++w) wr = combine(wr, s_warp[w]);
    right = combine(right, wr);
    double y = right.s + right.d * carry;  // y just after this thread's items
#pragma unroll
    for (int j = SC_ITEMS - 1; j >= 0; --j) {
      y = x[j] + discount * y;
      const long long i = i0 + j;
      if (i >= 0 && i < n) out[i] = y;
    }
    if (tid == 0) s_carry = y;  // thread 0 holds the chunk's first element
    __syncthreads();
    carry = s_carry;
  }
}

// (x - mean) / std, unbiased std, no epsilon (torch.mean / torch.std of utils.py:90-92); float64 accumulation
__global__ void __launch_bounds__(1024) normalize_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ double s_a[32], s_b[32];
  __shared__ double s_mean, s_inv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a = 0.0;
  for (long long i = tid; i < n; i += 1024) a += (double)x[i];
  a = warp_sum(a);
  if (lane == 0) s_a[warp] = a;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += s_a[w];
    s_mean = t / (double)n;
  }
  __syncthreads();
  const double mean = s_mean;
  double b = 0.0;
  for (long long i = tid; i < n; i += 1024) {
    const double d = (double)x[i] - mean;
    b += d * d;
  }
  b = warp_sum(b);
  if (lane == 0) s_b[warp] = b;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += s_b[w];
    s_inv = sqrt(t / (double)(n - 1));  // n == 1: 0/0 = NaN, like torch.std of one element
  }
  __syncthreads();
  const float meanf = (float)mean, stdf = (float)s_inv;
  for (long long i = tid; i < n; i += 1024) out[i] = (x[i] - meanf) / stdf;
}

// target <- f32(rho) * target + f32(1 - rho) * param, every product and the sum rounded separately (torch evaluates
// utils.py:55-57 as two multiplications and one addition of float32 tensors)
__global__ void polyak_rn_kernel(float* target, const float* __restrict__ param, long long n, float rho, float omr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) target[i] = __fadd_rn(__fmul_rn(rho, target[i]), __fmul_rn(omr, param[i]));
}
}  // namespace

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_discounted_cumsum(const double* x, int64_t n, double discount, double* out, void* stream) {
  B200RL_REQUIRE(n >= 0 && (n == 0 || (x && out)), "discounted_cumsum: bad argument");
  if (n == 0) return 0;
  discounted_cumsum_kernel<0><<<1, SC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, nullptr, n, 0.0, discount, out);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int b200rl_gae_f64(const double* rewards, const void* values, int values_f32, int64_t n, double gamma,
                              double gae_lambda, double* out, void* stream) {
  B200RL_REQUIRE(n >= 0 && (n == 0 || (rewards && values && out)), "gae_f64: bad argument");
  if (n == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (values_f32)
    discounted_cumsum_kernel<2><<<1, SC_THREADS, 0, s>>>(rewards, values, n, gamma, gamma * gae_lambda, out);
  else
    discounted_cumsum_kernel<1><<<1, SC_THREADS, 0, s>>>(rewards, values, n, gamma, gamma * gae_lambda, out);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int b200rl_normalize(const float* x, int64_t n, float* out, void* stream) {
  B200RL_REQUIRE(n >= 1 && x && out, "normalize: bad argument");
  normalize_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int b200rl_polyak(float* target, const float* param, int64_t n, double rho, void* stream) {
  B200RL_REQUIRE(n >= 0 && (n == 0 || (target && param)), "polyak: bad argument");
  if (n == 0) return 0;
  polyak_rn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(target, param, n, (float)rho,
                                                                                             (float)(1.0 - rho));
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
