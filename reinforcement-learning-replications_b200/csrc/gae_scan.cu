// gae_scan: bootstrapped rewards + discounted returns + TD residuals + GAE as ONE segmented reverse scan.
//
// Replaces (reference: /root/reference/src/rl_replicas/): utils.py:14-28 discounted_cumulative_sums,
// utils.py:31-44 gae, utils.py:74-87 bootstrap_rewards_with_last_values and the per-episode Python loops of
// algorithms/ppo.py:142-161.
//
// Both recurrences are first-order linear:  y_i = b_i + a_i * y_{i+1}  scanned from the END of the flat transition
// array, with a_i = 0 on the last step of every episode (segment reset):
//   returns:    a = gamma,        b_i = r_i            (+ gamma * V(last_obs) on the last step of a NOT-done episode)
//   advantages: a = gamma*lambda, b_i = delta_i = r_i + f32(gamma*v_{i+1}) - v_i     (v_L = V(last_obs) even when done)
// The pair (a,b) composes associatively: (a1,b1) o (a2,b2) = (a1*a2, b1 + a1*b2), so the scan is
//   thread: 8 consecutive items sequentially (exactly the reference's float64 recurrence inside a thread),
//   warp:   suffix scan of the 32 thread aggregates with shuffles,
//   CTA:    8 warp aggregates through shared memory,
//   grid:   single-pass decoupled look-back over tile descriptors (tiles take tickets from the end of the array, so a
//           tile only ever waits for tiles that already started; a tile that contains an episode end has a == 0 and
//           cuts the chain).
// HBM traffic = algorithmic traffic: read r (4 or 8 B) + v (4 B), write adv (4 B) + ret (4 B) per transition.
// All carries are float64 (the reference scans in float64, utils.py:28); outputs are cast to float32 like ppo.py:151,160.
#include "common.cuh"

namespace b200rl {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 transitions per CTA

struct __align__(16) ScanTileState {
  double a_ret, b_ret, a_adv, b_adv;  // tile aggregate, valid once status >= 1
  double y_ret, y_adv;                // recurrence values at the tile's FIRST element, valid once status == 2
  int status;
  int pad[3];
};
static_assert(sizeof(ScanTileState) == 64, "tile state is one 64-byte record");

// Workspace header.  The workspace is zeroed ONCE (at allocation); after that every launch cleans up after itself:
// tile status words carry the launch epoch (a stale word from an earlier launch reads as "not ready"), and the last
// CTA to finish resets the ticket / done counters and bumps the epoch.  => one kernel launch per scan, no memset.
struct ScanHeader {
  int ticket;
  int done;
  int epoch;
  int error;
  int pad[12];
};

struct Aff {
  double a, b;
};
__device__ __forceinline__ Aff compose(const Aff first, const Aff later) {
  return Aff{first.a * later.a, first.b + first.a * later.b};
}

struct ScanArgs {
  const void* rew;
  const float* values;
  const float* last_values;
  const long long* off;
  const unsigned char* done;
  long long n, n_ep;
  double gamma, gl;
  float gamma_f;
  float* adv;
  float* ret;
  ScanHeader* hdr;
  ScanTileState* tiles;
  double2* partial;
  double* stats;
  int num_tiles;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename RewT>
__device__ __forceinline__ void load_rewards(const RewT* src, long long i0, bool full, long long n, double (&r)[SCAN_ITEMS]);

template <>
__device__ __forceinline__ void load_rewards<float>(const float* src, long long i0, bool full, long long n,
                                                    double (&r)[SCAN_ITEMS]) {
  if (full) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(src + i0));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(src + i0 + 4));
    r[0] = x0.x; r[1] = x0.y; r[2] = x0.z; r[3] = x0.w;
    r[4] = x1.x; r[5] = x1.y; r[6] = x1.z; r[7] = x1.w;
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) r[j] = (i0 + j < n) ? (double)src[i0 + j] : 0.0;
  }
}
template <>
__device__ __forceinline__ void load_rewards<double>(const double* src, long long i0, bool full, long long n,
                                                     double (&r)[SCAN_ITEMS]) {
  if (full) {
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS / 2; ++q) {
      const double2 x = __ldg(reinterpret_cast<const double2*>(src + i0 + 2 * q));
      r[2 * q] = x.x;
      r[2 * q + 1] = x.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) r[j] = (i0 + j < n) ? src[i0 + j] : 0.0;
  }
}

// suffix scan of per-thread aggregates inside a warp; returns the EXCLUSIVE suffix (composition of lanes > lane)
// and leaves the inclusive aggregate of the whole warp in lane 0's `incl`.
__device__ __forceinline__ Aff warp_suffix_exclusive(Aff& incl, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double a2 = __shfl_down_sync(0xffffffffu, incl.a, o);
    const double b2 = __shfl_down_sync(0xffffffffu, incl.b, o);
    if (lane + o < 32) {
      incl.b = incl.b + incl.a * b2;
      incl.a = incl.a * a2;
    }
  }
  Aff ex;
  ex.a = __shfl_down_sync(0xffffffffu, incl.a, 1);
  ex.b = __shfl_down_sync(0xffffffffu, incl.b, 1);
  if (lane == 31) {
    ex.a = 1.0;
    ex.b = 0.0;
  }
  return ex;
}

// Warp-cooperative 32-ary search: max e in [0, n_ep) with off[e] <= target (requires off[0] <= target < off[n_ep]).
// ~log32(n_ep) rounds of one coalesced probe per lane instead of log2(n_ep) dependent loads per thread.
__device__ __forceinline__ long long warp_find_episode(const long long* __restrict__ off, long long n_ep,
                                                       long long target, int lane) {
  long long lo = 0, hi = n_ep;  // invariant: off[lo] <= target < off[hi]
  while (hi - lo > 1) {
    const long long step = (hi - lo + 31) / 32;
    const long long idx = lo + (long long)lane * step;
    const bool ok = (idx < hi) && (__ldg(off + idx) <= target);  // monotone in lane; lane 0 is always true
    const int cnt = __popc(__ballot_sync(0xffffffffu, ok));
    lo = lo + (long long)(cnt - 1) * step;
    hi = (lo + step < hi) ? lo + step : hi;
  }
  return lo;
}

template <typename RewT>
__global__ void __launch_bounds__(SCAN_THREADS, 4) gae_scan_kernel(const ScanArgs p) {
  __shared__ int s_tile, s_epoch, s_last;
  __shared__ long long s_e0, s_e1;
  __shared__ Aff s_wret[SCAN_THREADS / 32], s_wadv[SCAN_THREADS / 32];
  __shared__ double s_carry[2];
  __shared__ double s_red[2][SCAN_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_epoch = *reinterpret_cast<volatile int*>(&p.hdr->epoch);  // constant for the whole launch
    s_tile = p.num_tiles - 1 - atomicAdd(&p.hdr->ticket, 1);
  }
  __syncthreads();
  const int tile = s_tile;
  const int st_agg = s_epoch * 4 + 1, st_incl = s_epoch * 4 + 2;  // status words of THIS launch
  const long long n = p.n;
  const long long i0 = (long long)tile * SCAN_TILE + (long long)tid * SCAN_ITEMS;
  const bool full = (i0 + SCAN_ITEMS <= n);

  // ---- loads first (vectorised when the thread's 8 items are all in range): their DRAM latency overlaps the
  //      episode search below ----
  double r[SCAN_ITEMS];
  float v[SCAN_ITEMS + 1];
  load_rewards<RewT>(static_cast<const RewT*>(p.rew), i0, full, n, r);
  if (full) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(p.values + i0));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(p.values + i0 + 4));
    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
    v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) v[j] = (i0 + j < n) ? p.values[i0 + j] : 0.f;
  }
  v[SCAN_ITEMS] = (i0 + SCAN_ITEMS < n) ? __ldg(p.values + i0 + SCAN_ITEMS) : 0.f;

  // ---- episodes that overlap this tile: found once per CTA, cooperatively by warp 0 ----
  if (warp == 0) {
    const long long t0 = (long long)tile * SCAN_TILE;
    const long long t1 = (t0 + SCAN_TILE - 1 < n - 1) ? t0 + SCAN_TILE - 1 : n - 1;
    const long long e0 = warp_find_episode(p.off, p.n_ep, t0, lane);
    const long long e1 = warp_find_episode(p.off, p.n_ep, t1, lane);
    if (lane == 0) {
      s_e0 = e0;
      s_e1 = e1;
    }
  }
  __syncthreads();

  // ---- which episode does item i0 belong to?  e = max{e : off[e] <= i0} ----
  long long e = 0, next_off = 0;
  if (i0 < n) {
    long long lo = s_e0, hi = s_e1 + 1;  // only the episodes that overlap this tile
    while (hi - lo > 1) {
      const long long mid = (lo + hi) >> 1;
      if (__ldg(p.off + mid) <= i0) lo = mid; else hi = mid;
    }
    e = lo;
    next_off = __ldg(p.off + e + 1);
  }

  // ---- per-item (a,b) of both recurrences; a is gamma / gamma*lambda, or 0 on an episode's last step ----
  double b_ret[SCAN_ITEMS], b_adv[SCAN_ITEMS];
  unsigned cut = 0;  // bit j set: a_j == 0 (episode end, or out of range)
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long long i = i0 + j;
    if (i < n) {
      while (i >= next_off) {
        ++e;
        next_off = __ldg(p.off + e + 1);
      }
      const bool last = (i == next_off - 1);
      float vnext = v[j + 1];
      double boot = 0.0;
      if (last) {
        const float vl = __ldg(p.last_values + e);
        vnext = vl;                                            // utils.py:41: delta_{L-1} uses V(last_obs) even when done
        if (!__ldg(p.done + e)) boot = p.gamma * (double)vl;   // utils.py:81-85 + ppo.py:149: ret_{L-1} = r + gamma*R_L
        cut |= 1u << j;
      }
      // utils.py:41: rewards[:-1] (f64) + gamma*values[1:] (evaluated in float32) - values[:-1]
      const double delta = (r[j] + (double)__fmul_rn(p.gamma_f, vnext)) - (double)v[j];
      b_ret[j] = r[j] + boot;
      b_adv[j] = delta;
    } else {
      cut |= 1u << j;
      b_ret[j] = 0.0;
      b_adv[j] = 0.0;
    }
  }

  // ---- thread aggregate: compose items 7..0 ----
  Aff tr{1.0, 0.0}, ta{1.0, 0.0};
#pragma unroll
  for (int j = SCAN_ITEMS - 1; j >= 0; --j) {
    const bool c = (cut >> j) & 1u;
    const double ar = c ? 0.0 : p.gamma, aa = c ? 0.0 : p.gl;
    tr.b = b_ret[j] + ar * tr.b;
    tr.a = ar * tr.a;
    ta.b = b_adv[j] + aa * ta.b;
    ta.a = aa * ta.a;
  }

  // ---- warp + CTA suffix scans ----
  const Aff ex_r = warp_suffix_exclusive(tr, lane);
  const Aff ex_a = warp_suffix_exclusive(ta, lane);
  if (lane == 0) {
    s_wret[warp] = tr;
    s_wadv[warp] = ta;
  }
  __syncthreads();

  // ---- thread 0: publish the tile aggregate, look back over later tiles for the carry-in, publish inclusive ----
  if (tid == 0) {
    Aff agg_r{1.0, 0.0}, agg_a{1.0, 0.0};
    for (int w = SCAN_THREADS / 32 - 1; w >= 0; --w) {
      agg_r = compose(s_wret[w], agg_r);
      agg_a = compose(s_wadv[w], agg_a);
    }
    ScanTileState* me = p.tiles + tile;
    me->a_ret = agg_r.a; me->b_ret = agg_r.b; me->a_adv = agg_a.a; me->b_adv = agg_a.b;
    st_release(&me->status, st_agg);

    Aff acc_r{1.0, 0.0}, acc_a{1.0, 0.0};
    for (int j = tile + 1; j < p.num_tiles; ++j) {
      if (acc_r.a == 0.0 && acc_a.a == 0.0) break;  // an episode end in between: nothing further matters
      const ScanTileState* t = p.tiles + j;
      int st = 0;
      long long spins = 0;
      while ((st = ld_acquire(&t->status)) != st_agg && st != st_incl) {
        if (++spins > (1ll << 22)) break;  // bounded: never hang the GPU; flag and bail out
        __nanosleep(32);
      }
      if (st != st_agg && st != st_incl) {
        p.hdr->error = 1;
        break;
      }
      if (st == st_incl) {
        acc_r = Aff{0.0, acc_r.b + acc_r.a * __ldcg(&t->y_ret)};
        acc_a = Aff{0.0, acc_a.b + acc_a.a * __ldcg(&t->y_adv)};
        break;
      }
      acc_r = compose(acc_r, Aff{__ldcg(&t->a_ret), __ldcg(&t->b_ret)});
      acc_a = compose(acc_a, Aff{__ldcg(&t->a_adv), __ldcg(&t->b_adv)});
    }
    // beyond the end of the array the recurrence value is 0
    s_carry[0] = acc_r.b;
    s_carry[1] = acc_a.b;
    me->y_ret = agg_r.b + agg_r.a * acc_r.b;
    me->y_adv = agg_a.b + agg_a.a * acc_a.b;
    st_release(&me->status, st_incl);
  }
  __syncthreads();

  // ---- carry-in of this thread = (lanes after me in my warp) o (warps after mine) applied to the tile carry ----
  Aff xw_r{1.0, 0.0}, xw_a{1.0, 0.0};
  for (int w = SCAN_THREADS / 32 - 1; w > warp; --w) {
    xw_r = compose(s_wret[w], xw_r);
    xw_a = compose(s_wadv[w], xw_a);
  }
  const Aff x_r = compose(ex_r, xw_r), x_a = compose(ex_a, xw_a);
  double y_r = x_r.b + x_r.a * s_carry[0];
  double y_a = x_a.b + x_a.a * s_carry[1];

  // ---- final sequential recurrence over the thread's items (the reference's own float64 loop) ----
  float o_ret[SCAN_ITEMS], o_adv[SCAN_ITEMS];
  double s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int j = SCAN_ITEMS - 1; j >= 0; --j) {
    const bool c = (cut >> j) & 1u;
    y_r = b_ret[j] + (c ? 0.0 : p.gamma) * y_r;
    y_a = b_adv[j] + (c ? 0.0 : p.gl) * y_a;
    o_ret[j] = (float)y_r;
    o_adv[j] = (float)y_a;
    if (i0 + j < n) {
      const double af = (double)o_adv[j];  // statistics of the float32 tensor, like normalize_tensor's input
      s1 += af;
      s2 += af * af;
    }
  }
  if (full) {
    *reinterpret_cast<float4*>(p.ret + i0) = make_float4(o_ret[0], o_ret[1], o_ret[2], o_ret[3]);
    *reinterpret_cast<float4*>(p.ret + i0 + 4) = make_float4(o_ret[4], o_ret[5], o_ret[6], o_ret[7]);
    *reinterpret_cast<float4*>(p.adv + i0) = make_float4(o_adv[0], o_adv[1], o_adv[2], o_adv[3]);
    *reinterpret_cast<float4*>(p.adv + i0 + 4) = make_float4(o_adv[4], o_adv[5], o_adv[6], o_adv[7]);
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
      if (i0 + j < n) {
        p.ret[i0 + j] = o_ret[j];
        p.adv[i0 + j] = o_adv[j];
      }
  }

  // ---- per-tile advantage statistics (fixed order => deterministic) ----
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    s_red[0][warp] = s1;
    s_red[1][warp] = s2;
  }
  __syncthreads();
  if (tid == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
      t1 += s_red[0][w];
      t2 += s_red[1][w];
    }
    p.partial[tile] = make_double2(t1, t2);
    __threadfence();
    s_last = (atomicAdd(&p.hdr->done, 1) == p.num_tiles - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;

  // ---- last CTA: fixed-order sum of the per-tile statistics -> stats = (sum, sum of squares, n); reset the header ----
  __threadfence();
  __shared__ double f1[SCAN_THREADS], f2[SCAN_THREADS];
  double fa = 0.0, fb = 0.0;
  for (int i = tid; i < p.num_tiles; i += SCAN_THREADS) {
    const double2 x = __ldcg(p.partial + i);
    fa += x.x;
    fb += x.y;
  }
  f1[tid] = fa;
  f2[tid] = fb;
  __syncthreads();
  for (int o = SCAN_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o) {
      f1[tid] += f1[tid + o];
      f2[tid] += f2[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const bool bad = p.hdr->error != 0;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    p.stats[0] = bad ? nan : f1[0];
    p.stats[1] = bad ? nan : f2[0];
    p.stats[2] = (double)n;
    p.hdr->ticket = 0;
    p.hdr->done = 0;
    p.hdr->error = 0;
    p.hdr->epoch = (s_epoch + 1) & 0x0fffffff;
  }
}

static inline int scan_tiles(int64_t n) { return (int)((n + SCAN_TILE - 1) / SCAN_TILE); }

}  // namespace b200rl

using namespace b200rl;

extern "C" size_t b200rl_gae_scan_workspace_bytes(int64_t n) {
  const size_t t = (size_t)scan_tiles(n < 1 ? 1 : n);
  return sizeof(ScanHeader) + t * sizeof(ScanTileState) + t * sizeof(double2);
}

extern "C" int b200rl_gae_scan(const void* rewards, int rewards_f64, const float* values, const float* last_values,
                               const int64_t* ep_offsets, const uint8_t* ep_done, int64_t n, int64_t n_ep,
                               double gamma, double gae_lambda, float* adv_raw, float* ret, double* stats,
                               void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200RL_REQUIRE(n >= 0 && n_ep >= 0, "gae_scan: negative size");
  B200RL_REQUIRE(stats != nullptr, "gae_scan: stats is NULL");
  if (n == 0) {
    B200RL_CUDA(cudaMemsetAsync(stats, 0, 3 * sizeof(double), s));
    return 0;
  }
  B200RL_REQUIRE(n_ep >= 1, "gae_scan: %lld transitions but no episode", (long long)n);
  B200RL_REQUIRE(rewards && values && last_values && ep_offsets && ep_done && adv_raw && ret && workspace,
                 "gae_scan: NULL pointer argument");
  B200RL_REQUIRE(aligned16(rewards) && aligned16(values) && aligned16(adv_raw) && aligned16(ret) &&
                     aligned16(workspace),
                 "gae_scan: rewards/values/adv/ret/workspace must be 16-byte aligned");
  B200RL_REQUIRE(workspace_bytes >= b200rl_gae_scan_workspace_bytes(n), "gae_scan: workspace too small");
  const int tiles = scan_tiles(n);
  ScanArgs a;
  a.rew = rewards;
  a.values = values;
  a.last_values = last_values;
  a.off = reinterpret_cast<const long long*>(ep_offsets);
  a.done = ep_done;
  a.n = n;
  a.n_ep = n_ep;
  a.gamma = gamma;
  a.gl = gamma * gae_lambda;
  a.gamma_f = (float)gamma;
  a.adv = adv_raw;
  a.ret = ret;
  a.hdr = static_cast<ScanHeader*>(workspace);
  a.tiles = reinterpret_cast<ScanTileState*>(static_cast<char*>(workspace) + sizeof(ScanHeader));
  a.partial = reinterpret_cast<double2*>(reinterpret_cast<char*>(a.tiles) + (size_t)tiles * sizeof(ScanTileState));
  a.num_tiles = tiles;
  a.stats = stats;
  if (rewards_f64)
    gae_scan_kernel<double><<<tiles, SCAN_THREADS, 0, s>>>(a);
  else
    gae_scan_kernel<float><<<tiles, SCAN_THREADS, 0, s>>>(a);
  B200RL_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
