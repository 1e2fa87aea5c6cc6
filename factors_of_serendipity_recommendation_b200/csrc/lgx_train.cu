// Training-side kernels: fused BPR loss forward / backward scatter, Adam, the device BPR sampler
// and the ranking metrics.
//
// Replaces getEmbedding + bpr_loss (PT/model.py:186-209: 6 gathers + ~12 elementwise kernels), their
// autograd backward (index_select backward = scatter-add), torch.optim.Adam.step (PT/utils.py:50),
// UniformSample_original (PT/utils.py:55-99, PT/sources/sampling.cpp:27-56) and the numpy metrics
// (PT/utils.py:218-285).
#include <algorithm>
#include <cmath>

#include "lgx_common.cuh"
#include "lgx_topk.cuh"

namespace lgx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per (user, pos, neg) sample.  per[b] = coef, per[B+b] = softplus term, per[2B+b] = reg term
__global__ void __launch_bounds__(256)
k_bpr_forward(const float* __restrict__ light, const float* __restrict__ E0, const int64_t* __restrict__ users,
              const int64_t* __restrict__ pos, const int64_t* __restrict__ neg, int B, int n_users, int d,
              float* __restrict__ per) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  // lgx_sample_bpr marks a triple it could not draw (a user without positives in per-user mode, a user who interacted
  // with every item) with -1: such a triple contributes nothing (it used to index the last user's row as an item)
  if (users[b] < 0 || users[b] >= n_users || pos[b] < 0 || neg[b] < 0) {
    if (lane == 0) { per[b] = 0.f; per[B + b] = 0.f; per[2 * B + b] = 0.f; }
    return;
  }
  const int64_t ur = users[b], pr = (int64_t)n_users + pos[b], nr = (int64_t)n_users + neg[b];
  float sp = 0.f, sn = 0.f, reg = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float u = light[ur * d + c];
    sp = fmaf(u, light[pr * d + c], sp);
    sn = fmaf(u, light[nr * d + c], sn);
    const float u0 = E0[ur * d + c], p0 = E0[pr * d + c], n0 = E0[nr * d + c];
    reg += u0 * u0 + p0 * p0 + n0 * n0;
  }
  sp = warp_sum(sp); sn = warp_sum(sn); reg = warp_sum(reg);
  if (lane == 0) {
    const float x = sn - sp;
    // torch.nn.functional.softplus (beta=1, threshold=20), PT/model.py:207
    const float soft = x > 20.f ? x : log1pf(expf(x));
    const float sig = x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
    per[b] = sig / (float)B;
    per[B + b] = soft;
    per[2 * B + b] = reg;
  }
}

// fixed-order block reduction of the per-sample terms -> out2 = {mean softplus, 0.5*sum(reg)/B}
__global__ void __launch_bounds__(1024)
k_bpr_reduce(const float* __restrict__ per, int B, float* __restrict__ out2) {
  __shared__ float s0[1024], s1[1024];
  float a = 0.f, r = 0.f;
  for (int b = threadIdx.x; b < B; b += 1024) { a += per[B + b]; r += per[2 * B + b]; }
  s0[threadIdx.x] = a; s1[threadIdx.x] = r;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = s0[0] / (float)B; out2[1] = 0.5f * s1[0] / (float)B; }
}

// d loss / d light rows: G[u] += s (n - p); G[p] -= s u; G[n] += s u   with s = coef * grad_scale
__global__ void __launch_bounds__(256)
k_bpr_backward_light(const float* __restrict__ light, const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
                     const int64_t* __restrict__ neg, const float* __restrict__ coef, int B, int n_users, int d,
                     float grad_scale, const float* __restrict__ grad_scale_dev, float* __restrict__ G) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  if (users[b] < 0 || users[b] >= n_users || pos[b] < 0 || neg[b] < 0) return;      // undrawn triple (see k_bpr_forward)
  const int64_t ur = users[b], pr = (int64_t)n_users + pos[b], nr = (int64_t)n_users + neg[b];
  const float s = coef[b] * grad_scale * (grad_scale_dev ? __ldg(grad_scale_dev) : 1.0f);
  for (int c = lane; c < d; c += 32) {
    const float u = light[ur * d + c], p = light[pr * d + c], n = light[nr * d + c];
    atomicAdd(G + ur * d + c, s * (n - p));
    atomicAdd(G + pr * d + c, -s * u);
    atomicAdd(G + nr * d + c, s * u);
  }
}

// d reg / d E0 rows: reg = 0.5 * sum(x^2) / B  ->  x / B per occurrence
__global__ void __launch_bounds__(256)
k_bpr_backward_reg(const float* __restrict__ E0, const int64_t* __restrict__ users, const int64_t* __restrict__ pos,
                   const int64_t* __restrict__ neg, int B, int n_users, int d, float scale,
                   const float* __restrict__ scale_dev, float* __restrict__ dE0) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  if (users[b] < 0 || users[b] >= n_users || pos[b] < 0 || neg[b] < 0) return;      // undrawn triple (see k_bpr_forward)
  if (scale_dev) scale *= __ldg(scale_dev);
  const int64_t rows[3] = {users[b], (int64_t)n_users + pos[b], (int64_t)n_users + neg[b]};
#pragma unroll
  for (int r = 0; r < 3; ++r)
    for (int c = lane; c < d; c += 32) atomicAdd(dE0 + rows[r] * d + c, scale * E0[rows[r] * d + c]);
}

// torch.optim.Adam single-tensor update (no amsgrad / weight decay / maximize):
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); p.addcdiv_(m, sqrt(v)/bc2_sqrt + eps, -lr/bc1)
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
       float w1, float b2, float w2, float step_size, float bc2_sqrt, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(w1, gi - m[i], m[i]);
    const float vi = fmaf(w2 * gi, gi, v[i] * b2);
    m[i] = mi;
    v[i] = vi;
    const float denom = __fdiv_rn(sqrtf(vi), bc2_sqrt) + eps;
    p[i] = p[i] - step_size * __fdiv_rn(mi, denom);
  }
}

// CUDA-graph friendly Adam: the step counter lives on the device.  state = {int32 step, float step_size,
// float bc2_sqrt, pad}; k_adam_prep advances it (bias corrections in double, like torch's host scalars).
__global__ void k_adam_prep(int32_t* __restrict__ state, float lr, float beta1, float beta2) {
  const int step = state[0] + 1;
  state[0] = step;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  reinterpret_cast<float*>(state)[1] = (float)((double)lr / bc1);
  reinterpret_cast<float*>(state)[2] = (float)sqrt(bc2);
}
__global__ void __launch_bounds__(256)
k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
           float w1, float b2, float w2, const int32_t* __restrict__ state, float eps) {
  const float step_size = __ldg(reinterpret_cast<const float*>(state) + 1);
  const float bc2_sqrt = __ldg(reinterpret_cast<const float*>(state) + 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(w1, gi - m[i], m[i]);
    const float vi = fmaf(w2 * gi, gi, v[i] * b2);
    m[i] = mi;
    v[i] = vi;
    const float denom = __fdiv_rn(sqrtf(vi), bc2_sqrt) + eps;
    p[i] = p[i] - step_size * __fdiv_rn(mi, denom);
  }
}

// splitmix64: counter-based, every (seed, sample, draw) triple gives an independent 64-bit word
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
k_sample_bpr(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int32_t n_users, int32_t m_items,
             int64_t n_samples, int32_t per_user, uint64_t seed, int64_t* __restrict__ out) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_samples; s += (int64_t)gridDim.x * blockDim.x) {
    uint64_t ctr = mix64(seed ^ mix64((uint64_t)s));
    int64_t u;
    if (per_user > 0) {
      u = s / per_user;                               // sampling.cpp:35-42: every user, per_user triples
    } else {
      do {                                            // utils.py:76,84: uniform user, skip users with no positives
        ctr = mix64(ctr);
        u = (int64_t)(ctr % (uint64_t)n_users);
      } while (indptr[u + 1] == indptr[u]);
    }
    const int64_t beg = indptr[u], len = indptr[u + 1] - beg;
    int64_t p = -1, n = -1;
    if (len > 0) {
      ctr = mix64(ctr);
      p = (int64_t)indices[beg + (int64_t)(ctr % (uint64_t)len)] - n_users;
      if (len < m_items) {
        while (true) {                                // rejection against the sorted train row
          ctr = mix64(ctr);
          n = (int64_t)(ctr % (uint64_t)m_items);
          const int32_t key = (int32_t)(n_users + n);
          int64_t lo = beg, hi = beg + len;
          while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (indices[mid] < key) lo = mid + 1; else hi = mid; }
          if (!(lo < beg + len && indices[lo] == key)) break;
        }
      }
    }
    out[3 * s] = u; out[3 * s + 1] = p; out[3 * s + 2] = n;
  }
}

// one warp per user: hits against the ground-truth list, recall / hit count / ndcg accumulated in fp64
__global__ void __launch_bounds__(256)
k_rank_metrics(const int64_t* __restrict__ topk, int B, int k_stride, int k, const int64_t* __restrict__ gt_ptr,
               const int64_t* __restrict__ gt_items, double* __restrict__ sums3) {
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (u >= B) return;
  const int64_t g0 = gt_ptr[u], glen = gt_ptr[u + 1] - g0;
  double dcg = 0.0;
  int hits = 0;
  for (int i = lane; i < k; i += 32) {
    const int64_t it = topk[(int64_t)u * k_stride + i];
    bool hit = false;
    for (int64_t q = 0; q < glen; ++q) hit |= (gt_items[g0 + q] == it);
    if (hit) { hits += 1; dcg += 1.0 / log2((double)(i + 2)); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    hits += __shfl_xor_sync(0xffffffffu, hits, o);
    dcg += __shfl_xor_sync(0xffffffffu, dcg, o);
  }
  if (lane == 0) {
    double idcg = 0.0;
    const int64_t lim = glen < k ? glen : k;
    for (int64_t i = 0; i < lim; ++i) idcg += 1.0 / log2((double)(i + 2));
    if (idcg == 0.0) idcg = 1.0;
    atomicAdd(sums3 + 0, glen > 0 ? (double)hits / (double)glen : 0.0);
    atomicAdd(sums3 + 1, (double)hits);
    atomicAdd(sums3 + 2, dcg / idcg);
  }
}

}  // namespace lgx

using namespace lgx;

extern "C" {

int lgx_bpr_forward(const float* light, const float* E0, const int64_t* users, const int64_t* pos, const int64_t* neg,
                    int32_t B, int32_t n_users, int32_t d, float* out2, float* coef, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(light && E0 && users && pos && neg && out2 && coef, "NULL argument");
  LGX_REQUIRE(B > 0 && d > 0, "B and d must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  k_bpr_forward<<<(B * 32 + 255) / 256, 256, 0, st>>>(light, E0, users, pos, neg, B, n_users, d, coef);
  k_bpr_reduce<<<1, 1024, 0, st>>>(coef, B, out2);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_bpr_backward_light(const float* light, const int64_t* users, const int64_t* pos, const int64_t* neg,
                           const float* coef, int32_t B, int32_t n_users, int32_t d, float grad_scale,
                           const float* grad_scale_dev, float* G, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(light && users && pos && neg && coef && G, "NULL argument");
  LGX_REQUIRE(B > 0 && d > 0, "B and d must be positive");
  k_bpr_backward_light<<<(B * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(light, users, pos, neg, coef, B, n_users,
                                                                              d, grad_scale, grad_scale_dev, G);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_bpr_backward_reg(const float* E0, const int64_t* users, const int64_t* pos, const int64_t* neg, int32_t B,
                         int32_t n_users, int32_t d, float grad_scale, const float* grad_scale_dev, float* dE0,
                         lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(E0 && users && pos && neg && dE0, "NULL argument");
  LGX_REQUIRE(B > 0 && d > 0, "B and d must be positive");
  k_bpr_backward_reg<<<(B * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(E0, users, pos, neg, B, n_users, d,
                                                                            grad_scale / (float)B, grad_scale_dev, dE0);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int32_t step, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(param && grad && exp_avg && exp_avg_sq, "NULL argument");
  LGX_REQUIRE(n > 0 && step >= 1, "n must be positive and step >= 1");
  // bias corrections in double like torch's python scalars (torch/optim/adam.py _single_tensor_adam)
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)std::sqrt(bc2);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
  k_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, 1.0f - beta1, beta2,
                                                  1.0f - beta2, step_size, bc2_sqrt, eps);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, int32_t* state4, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(param && grad && exp_avg && exp_avg_sq && state4, "NULL argument");
  LGX_REQUIRE(n > 0, "n must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  k_adam_prep<<<1, 1, 0, st>>>(state4, lr, beta1, beta2);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
  k_adam_dev<<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, 1.0f - beta1, beta2, 1.0f - beta2, state4, eps);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_sample_bpr(const lgx_graph* g, int64_t n_samples, int32_t per_user, uint64_t seed, int64_t* out,
                   lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(g && out, "NULL argument");
  LGX_REQUIRE(g->n_users > 0 && g->m_items > 0 && g->n_rows >= g->n_users, "graph has no user rows");
  LGX_REQUIRE(g->nnz > 0, "graph has no interactions to sample from");
  if (per_user > 0) n_samples = (int64_t)g->n_users * per_user;
  LGX_REQUIRE(n_samples > 0, "n_samples must be positive");
  const int blocks = (int)std::min<int64_t>((n_samples + 255) / 256, (int64_t)sm_count() * 16);
  k_sample_bpr<<<blocks, 256, 0, (cudaStream_t)stream>>>(g->indptr, g->indices, g->n_users, g->m_items, n_samples,
                                                        per_user, seed, out);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_rank_metrics(const int64_t* topk_idx, int32_t B, int32_t k_stride, int32_t k, const int64_t* gt_ptr,
                     const int64_t* gt_items, double* sums3, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(topk_idx && gt_ptr && gt_items && sums3, "NULL argument");
  LGX_REQUIRE(B > 0 && k > 0 && k <= k_stride, "need 0 < k <= k_stride");
  k_rank_metrics<<<(B * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(topk_idx, B, k_stride, k, gt_ptr, gt_items,
                                                                        sums3);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

}  // extern "C"
