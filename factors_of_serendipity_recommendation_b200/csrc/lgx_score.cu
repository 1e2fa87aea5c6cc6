// Scoring on CUDA cores: dense sigma(U I^T) (getUsersRating), the exact-fp32 fused score + mask +
// top-K, operand packing for the tcgen05 path, and the final per-user merge.
//
// Replaces PT/model.py:181-183 (index_select + cuBLAS SGEMM + sigmoid), PT/Procedure.py:129-134
// (index_put_ of -1024 over train items) and :135 (torch.topk).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "lgx_common.cuh"
#include "lgx_score_plan.cuh"
#include "lgx_topk.cuh"

namespace lgx {

// ---------------------------------------------------------------------------------------------
// 64 x 64 output tile per CTA (256 threads, 4 x 4 per thread), K-slabs of 16 staged in shared memory.
constexpr int TU = 64, TI = 64, TK = 16;

struct TileAcc {
  float c[4][4];
};

// acc[i][j] = <U[u0 + ty*4 + i], I[j0 + tx*4 + j]>   (rows past B / M read as zero)
__device__ __forceinline__ void sgemm_tile(const float* __restrict__ U, const int64_t* __restrict__ rows, int B,
                                           const float* __restrict__ I, int M, int d, int u0, int j0,
                                           float (*Us)[TU + 4], float (*Is)[TI + 4], TileAcc& acc) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.c[i][j] = 0.f;
  // loader mapping: thread loads 4 consecutive k of one row: row = tid / 4, k4 = (tid % 4) * 4
  const int lr = tid >> 2, lk = (tid & 3) << 2;
  const int ur = u0 + lr, ir = j0 + lr;
  const float* up = nullptr;
  if (ur < B) up = U + (rows ? rows[ur] : (int64_t)ur) * d;
  const float* ip = ir < M ? I + (int64_t)ir * d : nullptr;
  const bool vec = (d & 3) == 0;
  for (int k0 = 0; k0 < d; k0 += TK) {
    float uv[4] = {0.f, 0.f, 0.f, 0.f}, iv[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + lk;
    if (vec && k + 3 < d) {
      if (up) { const float4 t = __ldg(reinterpret_cast<const float4*>(up + k)); uv[0] = t.x; uv[1] = t.y; uv[2] = t.z; uv[3] = t.w; }
      if (ip) { const float4 t = __ldg(reinterpret_cast<const float4*>(ip + k)); iv[0] = t.x; iv[1] = t.y; iv[2] = t.z; iv[3] = t.w; }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k + q < d) {
          if (up) uv[q] = __ldg(up + k + q);
          if (ip) iv[q] = __ldg(ip + k + q);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Us[lk + q][lr] = uv[q];
      Is[lk + q][lr] = iv[q];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Us[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Is[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.c[i][j] = fmaf(av[i], bv[j], acc.c[i][j]);
    }
  }
}

__global__ void __launch_bounds__(256)
k_score_dense(const float* __restrict__ U, const int64_t* __restrict__ users, int B, const float* __restrict__ I,
              int M, int d, float* __restrict__ out, int apply_sigmoid) {
  __shared__ __align__(16) float Us[TK][TU + 4];
  __shared__ __align__(16) float Is[TK][TI + 4];
  const int j0 = blockIdx.x * TI, u0 = blockIdx.y * TU;
  TileAcc acc;
  sgemm_tile(U, users, B, I, M, d, u0, j0, Us, Is, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int u = u0 + ty * 4 + i;
    if (u >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = j0 + tx * 4 + j;
      if (it >= M) continue;
      float s = acc.c[i][j];
      if (apply_sigmoid) s = 1.0f / (1.0f + expf(-s));   // nn.Sigmoid, PT/model.py:183
      out[(int64_t)u * M + it] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The candidate-bucketing primitive of the serendipity pipeline (/root/reference/recommend.py:375-380):
//     mat_dis = np.dot(emb_user, emb_item.T).astype(np.float16); max_dis, min_dis = mat_dis.max() + eps, mat_dis.min()
//     mat_label = np.floor((mat_dis - min_dis) / ((max_dis - min_dis) / num_fold)).astype(np.int8)
// The reference materialises the [n_user, n_item] fp16 matrix on the host; here pass 1 reduces min / max of the
// fp16-rounded scores on the fly and pass 2 writes the int8 labels of a user batch, in the same fp16 arithmetic
// (subtract and divide rounded to half, like numpy's float16 ufuncs).
__device__ __forceinline__ unsigned f32_ordered(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
k_score_minmax(const float* __restrict__ U, int B, const float* __restrict__ I, int M, int d, unsigned* __restrict__ mm) {
  __shared__ __align__(16) float Us[TK][TU + 4];
  __shared__ __align__(16) float Is[TK][TI + 4];
  __shared__ unsigned s_min, s_max;
  if (threadIdx.x == 0) { s_min = 0xffffffffu; s_max = 0u; }
  const int j0 = blockIdx.x * TI, u0 = blockIdx.y * TU;
  TileAcc acc;
  sgemm_tile(U, nullptr, B, I, M, d, u0, j0, Us, Is, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  unsigned lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (u0 + ty * 4 + i < B && j0 + tx * 4 + j < M) {
        const unsigned o = f32_ordered(__half2float(__float2half_rn(acc.c[i][j])));
        lo = min(lo, o);
        hi = max(hi, o);
      }
  __syncthreads();
  atomicMin(&s_min, lo);
  atomicMax(&s_max, hi);
  __syncthreads();
  if (threadIdx.x == 0) { atomicMin(mm, s_min); atomicMax(mm + 1, s_max); }
}
__global__ void k_minmax_decode(const unsigned* __restrict__ mm, float* __restrict__ out2) {
  for (int k = 0; k < 2; ++k) {
    const unsigned u = mm[k];
    out2[k] = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
  }
}
__global__ void __launch_bounds__(256)
k_score_bucket(const float* __restrict__ U, const int64_t* __restrict__ users, int B, const float* __restrict__ I, int M,
               int d, float min_dis, float inter, int8_t* __restrict__ labels) {
  __shared__ __align__(16) float Us[TK][TU + 4];
  __shared__ __align__(16) float Is[TK][TI + 4];
  const int j0 = blockIdx.x * TI, u0 = blockIdx.y * TU;
  TileAcc acc;
  sgemm_tile(U, users, B, I, M, d, u0, j0, Us, Is, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const __half hmin = __float2half_rn(min_dis), hint = __float2half_rn(inter);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int u = u0 + ty * 4 + i;
    if (u >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = j0 + tx * 4 + j;
      if (it >= M) continue;
      const __half q = __hdiv(__hsub(__float2half_rn(acc.c[i][j]), hmin), hint);
      labels[(int64_t)u * M + it] = (int8_t)(int)floorf(__half2float(q));
    }
  }
}

// Exact-fp32 fused top-K: CTA = (user tile of 64, item split).  Scores above the row threshold are
// queued per row, then one thread per row drains its queue into the sorted list (mask checked there).
__global__ void __launch_bounds__(256)
k_score_topk_fp32(const float* __restrict__ U, const int64_t* __restrict__ users, int B, const float* __restrict__ I,
                  int M, int d, int K, TrainMask mask, int64_t item_offset, int n_splits, int tiles_per_split,
                  float* __restrict__ ws_val, int32_t* __restrict__ ws_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float(*Us)[TU + 4] = reinterpret_cast<float(*)[TU + 4]>(smem_raw);
  float(*Is)[TI + 4] = reinterpret_cast<float(*)[TI + 4]>(smem_raw + sizeof(float) * TK * (TU + 4));
  unsigned char* p = smem_raw + 2 * sizeof(float) * TK * (TU + 4);
  float* thresh = reinterpret_cast<float*>(p); p += sizeof(float) * TU;
  int* qcnt = reinterpret_cast<int*>(p); p += sizeof(int) * TU;
  float* qval = reinterpret_cast<float*>(p); p += sizeof(float) * TU * TI;
  int32_t* qidx = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * TU * TI;
  float* lval = reinterpret_cast<float*>(p); p += sizeof(float) * TU * K;
  int32_t* lidx = reinterpret_cast<int32_t*>(p);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int u0 = blockIdx.x * TU, split = blockIdx.y;
  for (int e = tid; e < TU * K; e += 256) { lval[e] = -CUDART_INF_F; lidx[e] = INT32_MAX; }
  if (tid < TU) { thresh[tid] = -CUDART_INF_F; qcnt[tid] = 0; }
  __syncthreads();
  const int n_tiles = (M + TI - 1) / TI;
  const int t_begin = split * tiles_per_split, t_end = min(n_tiles, t_begin + tiles_per_split);
  for (int t = t_begin; t < t_end; ++t) {
    const int j0 = t * TI;
    TileAcc acc;
    sgemm_tile(U, nullptr, B, I, M, d, u0, j0, Us, Is, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float th = thresh[r];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int it = j0 + tx * 4 + j;
        const float s = acc.c[i][j];
        if (it < M && s >= th) {   // >=: an equal score with a lower id could still win the tie
          const int slot = atomicAdd(&qcnt[r], 1);
          qval[r * TI + slot] = s;
          qidx[r * TI + slot] = it;
        }
      }
    }
    __syncthreads();
    if (tid < TU) {
      const int r = tid, n = qcnt[r];
      if (n > 0) {
        const int64_t uid = (u0 + r < B) ? (users ? users[u0 + r] : (int64_t)(u0 + r)) : -1;
        float th = thresh[r];
        for (int q = 0; q < n; ++q) {
          const float s = qval[r * TI + q];
          const int32_t it = qidx[r * TI + q];
          if (better(s, it, lval[(K - 1) * TU + r], lidx[(K - 1) * TU + r]) &&
              !mask.contains(uid, item_offset + it))
            th = topk_insert(lval + r, lidx + r, K, TU, s, it);
        }
        thresh[r] = th;
        qcnt[r] = 0;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < TU * K; e += 256) {
    const int r = e % TU, k = e / TU;
    if (u0 + r < B) {
      const int64_t o = ((int64_t)split * B + (u0 + r)) * K + k;
      ws_val[o] = lval[k * TU + r];
      ws_idx[o] = lidx[k * TU + r];
    }
  }
}

// Merge P sorted candidate lists per user into the final top-K (one thread per user walks the
// P list heads -- P*K is tiny).  If fewer than K unmasked items exist the tail is filled with the user's
// train items at -1024, which is what the reference's index_put_ + topk returns (PT/Procedure.py:134-135).
template <typename IdxT>
__global__ void __launch_bounds__(128)
k_topk_merge(const float* __restrict__ cand_val, const IdxT* __restrict__ cand_idx, int P, int B, int K,
             int64_t item_offset, int64_t m_local, TrainMask mask, const int64_t* __restrict__ users,
             int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= B) return;
  uint8_t head[kMaxSplits];   // K <= 255
  for (int p = 0; p < P; ++p) head[p] = 0;
  int got = 0;
  for (; got < K; ++got) {
    int best = -1;
    float bv = 0.f;
    int64_t bi = 0;
    for (int p = 0; p < P; ++p) {
      const int h = head[p];
      if (h >= K) continue;
      const int64_t o = ((int64_t)p * B + u) * K + h;
      const float v = cand_val[o];
      const int64_t i = (int64_t)cand_idx[o];
      if (v == -CUDART_INF_F) continue;   // empty slot: this list is exhausted
      if (best < 0 || v > bv || (v == bv && i < bi)) { best = p; bv = v; bi = i; }
    }
    if (best < 0) break;
    head[best]++;
    out_val[(int64_t)u * K + got] = bv;
    out_idx[(int64_t)u * K + got] = bi + item_offset;
  }
  if (got < K && mask.indptr != nullptr) {
    const int64_t uid = users ? users[u] : (int64_t)u;
    for (int64_t q = mask.indptr[uid]; q < mask.indptr[uid + 1] && got < K; ++q) {
      const int64_t it = (int64_t)mask.indices[q] - mask.n_users;
      if (it >= item_offset && it < item_offset + m_local) {
        out_val[(int64_t)u * K + got] = kMaskValue;
        out_idx[(int64_t)u * K + got] = it;
        ++got;
      }
    }
  }
  for (; got < K; ++got) {   // only reachable when a shard holds fewer than K items; merged away later
    out_val[(int64_t)u * K + got] = -CUDART_INF_F;
    out_idx[(int64_t)u * K + got] = -1;
  }
}

// fp32 rows -> bf16 operand rows [rows, Ktot]; BF16X3 splits x = hi + lo and lays out
// users [hi | hi | lo], items [hi | lo | hi] so one K=3d GEMM gives hi.hi + hi.lo + lo.hi.
__global__ void __launch_bounds__(256)
k_pack(const float* __restrict__ src, const int64_t* __restrict__ row_ids, int rows, int d, int mode, int is_items,
       __nv_bfloat16* __restrict__ dst) {
  const int64_t total = (int64_t)rows * d;
  const int ktot = mode == LGX_SCORE_BF16X3 ? 3 * d : d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / d), c = (int)(e % d);
    const float x = src[(row_ids ? row_ids[r] : (int64_t)r) * d + c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    __nv_bfloat16* o = dst + (int64_t)r * ktot;
    if (mode == LGX_SCORE_BF16X3) {
      const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
      o[c] = hi;
      o[d + c] = is_items ? lo : hi;
      o[2 * d + c] = is_items ? hi : lo;
    } else {
      o[c] = hi;
    }
  }
}

int launch_merge_i32(const float* ws_val, const int32_t* ws_idx, int P, int B, int K, int64_t item_offset,
                     int64_t m_local, TrainMask mask, const int64_t* users, int64_t* out_idx, float* out_val,
                     cudaStream_t st) {
  k_topk_merge<int32_t><<<(B + 127) / 128, 128, 0, st>>>(ws_val, ws_idx, P, B, K, item_offset, m_local, mask, users,
                                                        out_idx, out_val);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

TrainMask make_mask(const lgx_graph* g) {
  TrainMask m{nullptr, nullptr, 0};
  if (g) { m.indptr = g->indptr; m.indices = g->indices; m.n_users = g->n_users; }
  return m;
}

static size_t fp32_smem_bytes(int K) {
  return 2 * sizeof(float) * TK * (TU + 4) + sizeof(float) * TU + sizeof(int) * TU + 8 * (size_t)TU * TI +
         8 * (size_t)TU * K;
}

int score_topk_fp32(const lgx_graph* g, const float* U, const int64_t* users, int B, const float* I, int M, int d,
                    int K, int64_t item_offset, int64_t* out_idx, float* out_val, void* workspace, cudaStream_t st) {
  const ScorePlan plan = plan_score(B, M, TU, TI, sm_count(), 4);
  float* ws_val = reinterpret_cast<float*>(workspace);
  int32_t* ws_idx = reinterpret_cast<int32_t*>(ws_val + (size_t)plan.n_splits * B * K);
  const size_t smem = fp32_smem_bytes(K);
  static size_t configured[kMaxDevices] = {};     // the opt-in is a per-device function attribute
  const int dev = current_device();
  if (dev >= kMaxDevices || smem > configured[dev]) {
    LGX_CHECK_CUDA(cudaFuncSetAttribute(k_score_topk_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev < kMaxDevices) configured[dev] = smem;
  }
  const TrainMask mask = make_mask(g);
  dim3 grid(plan.n_user_tiles, plan.n_splits);
  k_score_topk_fp32<<<grid, 256, smem, st>>>(U, users, B, I, M, d, K, mask, item_offset, plan.n_splits,
                                             plan.tiles_per_split, ws_val, ws_idx);
  LGX_CHECK_LAUNCH();
  return launch_merge_i32(ws_val, ws_idx, plan.n_splits, B, K, item_offset, M, mask, users, out_idx, out_val, st);
}

}  // namespace lgx

using namespace lgx;

extern "C" {

int lgx_score_dense(const float* U, const int64_t* users, int32_t B, const float* I, int32_t M, int32_t d, float* out,
                    int32_t apply_sigmoid, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(U && I && out, "NULL argument");
  LGX_REQUIRE(B > 0 && M > 0 && d > 0, "B, M, d must be positive");
  dim3 grid((M + TI - 1) / TI, (B + TU - 1) / TU);
  LGX_REQUIRE(grid.y <= 65535, "batch too large for one call (max 65535*64 users)");
  k_score_dense<<<grid, 256, 0, (cudaStream_t)stream>>>(U, users, B, I, M, d, out, apply_sigmoid);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_score_minmax(const float* U, int32_t B, const float* I, int32_t M, int32_t d, float* out2, void* workspace8,
                     lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(U && I && out2 && workspace8, "NULL argument");
  LGX_REQUIRE(B > 0 && M > 0 && d > 0, "B, M, d must be positive");
  dim3 grid((M + TI - 1) / TI, (B + TU - 1) / TU);
  LGX_REQUIRE(grid.y <= 65535, "batch too large for one call (max 65535*64 users)");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* mm = reinterpret_cast<unsigned*>(workspace8);
  const unsigned init[2] = {0xffffffffu, 0u};
  LGX_CHECK_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
  k_score_minmax<<<grid, 256, 0, st>>>(U, B, I, M, d, mm);
  k_minmax_decode<<<1, 1, 0, st>>>(mm, out2);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_score_bucket(const float* U, const int64_t* users, int32_t B, const float* I, int32_t M, int32_t d, float min_dis,
                     float inter, int8_t* labels, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(U && I && labels, "NULL argument");
  LGX_REQUIRE(B > 0 && M > 0 && d > 0, "B, M, d must be positive");
  LGX_REQUIRE(inter > 0.0f, "inter must be positive");
  dim3 grid((M + TI - 1) / TI, (B + TU - 1) / TU);
  LGX_REQUIRE(grid.y <= 65535, "batch too large for one call (max 65535*64 users)");
  k_score_bucket<<<grid, 256, 0, (cudaStream_t)stream>>>(U, users, B, I, M, d, min_dis, inter, labels);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

size_t lgx_pack_bytes(int32_t rows, int32_t d, int32_t mode) {
  if (rows <= 0 || d <= 0 || mode == LGX_SCORE_FP32) return 0;
  const size_t ktot = mode == LGX_SCORE_BF16X3 ? 3 * (size_t)d : (size_t)d;
  return (size_t)rows * ktot * sizeof(__nv_bfloat16);
}

int lgx_pack_operand(const float* src, const int64_t* row_ids, int32_t rows, int32_t d, int32_t mode, int32_t is_items,
                     void* dst, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(src && dst, "NULL argument");
  LGX_REQUIRE(rows > 0 && d > 0, "rows and d must be positive");
  LGX_REQUIRE(mode == LGX_SCORE_BF16 || mode == LGX_SCORE_BF16X3, "pack is only defined for the bf16 modes");
  const int64_t total = (int64_t)rows * d;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
  k_pack<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, row_ids, rows, d, mode, is_items,
                                                   reinterpret_cast<__nv_bfloat16*>(dst));
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_topk_merge(const int64_t* cand_idx, const float* cand_val, int32_t P, int32_t B, int32_t k, int64_t* out_idx,
                   float* out_val, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(cand_idx && cand_val && out_idx && out_val, "NULL argument");
  LGX_REQUIRE(P > 0 && P <= kMaxSplits && B > 0 && k > 0 && k <= 255, "P must be in [1,160], B positive, k in [1,255]");
  TrainMask none{nullptr, nullptr, 0};
  k_topk_merge<int64_t><<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cand_val, cand_idx, P, B, k, 0, 0, none,
                                                                           nullptr, out_idx, out_val);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

}  // extern "C"
