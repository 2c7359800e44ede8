// TRPO vector kernels: conjugate gradient, step size, backtracking line search -- all device-side with skip flags, so
// the whole ConjugateGradientOptimizer.step() runs without a host round trip.
//
// Replaces (reference: /root/reference/src/rl_replicas/optimizers/conjugate_gradient_optimizer.py):
//   _conjugate_gradient :169-202, the step-size rule :83-95, _backtracking_line_search :204-250.
// The Hessian-vector products themselves are B200RL_LOSS_FVP launches of the fused MLP kernel (mlp_fused.cu).
// Vectors have a few thousand elements: every kernel is ONE CTA of 1024 threads (latency-bound by construction).
#include "common.cuh"

namespace b200rl {

constexpr int VT = 1024;

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < VT / 32; ++w) t += red[w];  // fixed order, every thread gets the same value
  return t;
}

// sc layout (doubles): 0 rdotr, 1 loss_before, 2 step_size, 3 new_loss, 4 kl, 5 xHx, 6 accepted ratio index
// flags (ints): 0 cg converged, 1 line search accepted, 2 step rejected
__global__ void __launch_bounds__(VT) cg_init_kernel(const float* g, float* x, float* r, float* pv, int n, double* sc,
                                                     int* flags) {
  __shared__ double red[VT / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += VT) {
    const float gi = g[i];
    x[i] = 0.f;
    r[i] = gi;
    pv[i] = gi;
    acc += (double)gi * (double)gi;
  }
  const double rdotr = block_sum(acc, red);
  if (threadIdx.x == 0) {
    sc[0] = (double)(float)rdotr;
    sc[6] = -1.0;
    flags[0] = 0;
    flags[1] = 0;
    flags[2] = 0;
  }
}

// one CG iteration given z_raw = F p (conjugate_gradient_optimizer.py:190-201); z = z_raw + damping * p (:165)
__global__ void __launch_bounds__(VT) cg_update_kernel(const float* z_raw, float damping, float* x, float* r, float* pv,
                                                       int n, double* sc, int* flags, float residual_tol) {
  __shared__ double red[VT / 32];
  if (flags[0] != 0) return;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += VT) {
    const float z = z_raw[i] + damping * pv[i];
    acc += (double)pv[i] * (double)z;
  }
  const float pz = (float)block_sum(acc, red);
  const float rdotr = (float)sc[0];
  const float v = rdotr / pz;
  acc = 0.0;
  for (int i = threadIdx.x; i < n; i += VT) {
    const float z = z_raw[i] + damping * pv[i];
    x[i] += v * pv[i];
    const float ri = r[i] - v * z;
    r[i] = ri;
    acc += (double)ri * (double)ri;
  }
  const float newrdotr = (float)block_sum(acc, red);
  const float mu = newrdotr / rdotr;
  for (int i = threadIdx.x; i < n; i += VT) pv[i] = r[i] + mu * pv[i];
  if (threadIdx.x == 0) {
    sc[0] = (double)newrdotr;
    if (newrdotr < residual_tol) flags[0] = 1;
  }
}

__global__ void __launch_bounds__(VT) nan_to_zero_kernel(float* x, int n) {  // step_direction[x != x] = 0 (:83)
  for (int i = threadIdx.x; i < n; i += VT)
    if (x[i] != x[i]) x[i] = 0.f;
}

// step_size = sqrt(2 delta / (x^T H x + 1e-8)), NaN -> 1 (:86-93); descent = step_size * x (:95); prev = params
__global__ void __launch_bounds__(VT) step_size_kernel(const float* x, const float* hx_raw, float damping, float delta,
                                                       float* descent, const float* params, float* prev, int n,
                                                       double* sc) {
  __shared__ double red[VT / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += VT) acc += (double)x[i] * (double)(hx_raw[i] + damping * x[i]);
  const float xhx = (float)block_sum(acc, red);
  float step = sqrtf(2.0f * delta * (1.0f / (xhx + 1e-8f)));
  if (step != step) step = 1.0f;
  for (int i = threadIdx.x; i < n; i += VT) {
    descent[i] = step * x[i];
    prev[i] = params[i];
  }
  if (threadIdx.x == 0) {
    sc[2] = (double)step;
    sc[5] = (double)xhx;
  }
}

__global__ void __launch_bounds__(VT) ls_set_params_kernel(float* params, const float* prev, const float* descent,
                                                           float ratio, int n, const int* flags) {
  if (flags[1] != 0) return;  // a previous ratio was accepted: keep those parameters
  for (int i = threadIdx.x; i < n; i += VT) params[i] = prev[i] - ratio * descent[i];
}

// accept iff new_loss < loss_before and kl <= delta (:230)
__global__ void ls_check_kernel(const double* slot, double n_rows, float delta, double* sc, int* flags, int index) {
  if (threadIdx.x != 0 || flags[1] != 0) return;
  const float new_loss = (float)(slot[0] / n_rows), kl = (float)(slot[6] / n_rows);
  sc[3] = (double)new_loss;
  sc[4] = (double)kl;
  if (new_loss < (float)sc[1] && kl <= delta) {
    flags[1] = 1;
    sc[6] = (double)index;
  }
}

// after the loop: reject iff NaN, or new_loss >= loss_before, or kl >= delta (:233-250) -> restore the parameters
__global__ void __launch_bounds__(VT) ls_final_kernel(float* params, const float* prev, int n, float delta,
                                                      const double* sc, int* flags) {
  const float new_loss = (float)sc[3], kl = (float)sc[4], before = (float)sc[1];
  const bool reject = (new_loss != new_loss) || (kl != kl) || new_loss >= before || kl >= delta;
  if (!reject) return;
  for (int i = threadIdx.x; i < n; i += VT) params[i] = prev[i];
  if (threadIdx.x == 0) flags[2] = 1;
}

__global__ void set_scalar_from_slot_kernel(double* dst, const double* slot, int k, double inv) {
  if (threadIdx.x == 0) *dst = (double)(float)(slot[k] * inv);
}

#define LAUNCH1(kern, s, ...)                    \
  do {                                           \
    kern<<<1, VT, 0, s>>>(__VA_ARGS__);          \
    B200RL_CUDA(cudaGetLastError());             \
    count_launch(1);                             \
  } while (0)

int trpo_cg_init(const float* g, float* x, float* r, float* pv, int n, double* sc, int* flags, cudaStream_t s) {
  LAUNCH1(cg_init_kernel, s, g, x, r, pv, n, sc, flags);
  return 0;
}
int trpo_cg_update(const float* z_raw, float damping, float* x, float* r, float* pv, int n, double* sc, int* flags,
                   cudaStream_t s) {
  LAUNCH1(cg_update_kernel, s, z_raw, damping, x, r, pv, n, sc, flags, 1e-10f);
  return 0;
}
int trpo_nan_to_zero(float* x, int n, cudaStream_t s) {
  LAUNCH1(nan_to_zero_kernel, s, x, n);
  return 0;
}
int trpo_step_size(const float* x, const float* hx_raw, float damping, float delta, float* descent, const float* params,
                   float* prev, int n, double* sc, cudaStream_t s) {
  LAUNCH1(step_size_kernel, s, x, hx_raw, damping, delta, descent, params, prev, n, sc);
  return 0;
}
int trpo_ls_set_params(float* params, const float* prev, const float* descent, float ratio, int n, const int* flags,
                       cudaStream_t s) {
  LAUNCH1(ls_set_params_kernel, s, params, prev, descent, ratio, n, flags);
  return 0;
}
int trpo_ls_check(const double* slot, double n_rows, float delta, double* sc, int* flags, int index, cudaStream_t s) {
  ls_check_kernel<<<1, 32, 0, s>>>(slot, n_rows, delta, sc, flags, index);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
int trpo_ls_final(float* params, const float* prev, int n, float delta, const double* sc, int* flags, cudaStream_t s) {
  LAUNCH1(ls_final_kernel, s, params, prev, n, delta, sc, flags);
  return 0;
}
// the float32 scalar sums piggy-backed behind an all-reduced gradient -> a float64 scalar slot
__global__ void tail_to_slot_kernel(const float* tail, double* slot) {
  if (threadIdx.x < B200RL_N_SCALARS) slot[threadIdx.x] = (double)tail[threadIdx.x];
}
int trpo_tail_to_slot(const float* tail, double* slot, cudaStream_t s) {
  tail_to_slot_kernel<<<1, 32, 0, s>>>(tail, slot);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
int trpo_set_scalar(double* dst, const double* slot, int k, double inv, cudaStream_t s) {
  set_scalar_from_slot_kernel<<<1, 32, 0, s>>>(dst, slot, k, inv);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace b200rl
