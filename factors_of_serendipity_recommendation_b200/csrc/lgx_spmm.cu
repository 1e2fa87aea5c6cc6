// Sparse propagation: Y = A_hat X with the LightGCN layer-mean fused into the epilogue.
//
// Replaces torch.sparse.mm at PT/model.py:171 (cuSPARSE SpMM on a COO tensor), torch.stack/mean at
// :173-175 and, through the symmetric A_hat, the autograd backward of both (PT/utils.py:49).
//
// Schedule: the graph handle carries work units (<= chunk_nnz non-zeros of one row) in degree-
// descending order.  A group of G lanes owns one unit: each lane keeps V float4 of the output row
// (d = 4*G*V; d=64 -> half-warp per row, 16 x 128-bit gathers per non-zero).  Units are dealt
// round-robin to groups, so every warp sees the same length distribution (load balance on
// power-law graphs without atomics).  Column/value streams are read coalesced once (G per step)
// and broadcast with shuffles; embedding rows are gathered with 128-bit read-only loads, U in
// flight per lane.  Rows longer than chunk_nnz write per-unit partials that a second kernel sums
// in a fixed order (deterministic, no float atomics).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "lgx_common.cuh"

namespace lgx {

__device__ __forceinline__ int32_t ld_stream_i32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_gather_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float drop_value(const DropSpec& drop, int64_t k, float a) {
  if (!drop.enabled) return a;
  const int64_t pos = drop.tpos ? (int64_t)__ldg(drop.tpos + k) : k;
  return lgx_keep(drop.seed, pos, drop.keep_prob) ? a * drop.inv_keep : 0.f;
}

__device__ __forceinline__ void fma4(float4& a, float v, const float4& x) {
  a.x = fmaf(v, x.x, a.x);
  a.y = fmaf(v, x.y, a.y);
  a.z = fmaf(v, x.z, a.z);
  a.w = fmaf(v, x.w, a.w);
}

// Fused all-gather: the epilogue may also store the row into every rank's copy of the gathered layer
// (peer device pointers mapped through CUDA IPC; NVLink P2P stores).  value = acc (next layer's input)
// or the finished running mean (last layer).
constexpr int kMaxPeers = 16;
struct PeerOut {
  float* ptr[kMaxPeers];
  int n;                 // 0 = no peer output
  int store_mean;        // 0: store acc, 1: store (S_in + acc) / div
  int64_t row_offset;    // first row of this rank's block inside the gathered layout
};

// Epilogue shared by the direct path and the long-row reducer.
__device__ __forceinline__ float4 epilogue4(float4 acc, int64_t off, const float* S_in,
                                            float* __restrict__ Y, float* S_out, float div) {
  if (Y) *reinterpret_cast<float4*>(Y + off) = acc;
  float4 s = acc;
  if (S_out) {
    s = *reinterpret_cast<const float4*>(S_in + off);
    s.x += acc.x; s.y += acc.y; s.z += acc.z; s.w += acc.w;
    if (div != 1.0f) {
      s.x = __fdiv_rn(s.x, div); s.y = __fdiv_rn(s.y, div); s.z = __fdiv_rn(s.z, div); s.w = __fdiv_rn(s.w, div);
    }
    *reinterpret_cast<float4*>(S_out + off) = s;
  }
  return s;
}
__device__ __forceinline__ void peer_store4(const PeerOut& po, int64_t row, int d, int col, float4 acc, float4 mean) {
  const float4 v = po.store_mean ? mean : acc;
  const int64_t off = (po.row_offset + row) * d + col;
  for (int p = 0; p < po.n; ++p) *reinterpret_cast<float4*>(po.ptr[p] + off) = v;
}

__device__ __forceinline__ float4 ld_gather_f4_keep(const float* p) {   // hot row: keep in L1
  float4 r;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_gather_f4_stream(const float* p) {  // cold row: do not pollute L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// G lanes per work unit, V float4 per lane.  EXACT: d == 4*G*V (no column bound checks).
// U gathers in flight per lane per sub-batch; MINB resident CTAs per SM the register budget targets.
// HOT: split the gathers by column degree -- for this graph value = dinv[row]*dinv[col], so
// "col degree >= hot_degree" is "value <= dinv[row] / sqrt(hot_degree)"; hot rows are kept in L1
// (evict_last), cold rows bypass it (no_allocate) so the popular rows stay resident per SM.
template <int G, int V, bool EXACT, int U, int MINB, bool HOT>
__global__ void __launch_bounds__(256, MINB)
k_spmm(const WorkItem* __restrict__ work, int64_t n_work, int64_t n_partials, const int32_t* __restrict__ indices,
       const float* __restrict__ values, const float* __restrict__ X, const float* S_in,
       float* __restrict__ Y, float* S_out, float* __restrict__ partial, float div, int d,
       const float* __restrict__ dinv, float hot_rsqrt, const DropSpec drop) {
  static_assert(G % U == 0 || U > G, "U must divide G");
  constexpr int UU = U > G ? G : U;
  const int lig = threadIdx.x & (G - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t n_rounds = (n_work + n_groups - 1) / n_groups;
  const int d4 = d >> 2;

  for (int64_t round = 0; round < n_rounds; ++round) {
    const int64_t item = round * n_groups + group;
    int64_t start = 0;
    int32_t row = 0, len = 0, part = -1;
    const bool have = item < n_work;
    if (have) {
      const int4 w0 = __ldg(reinterpret_cast<const int4*>(work + item));        // start (lo, hi), row, len
      start = ((int64_t)(uint32_t)w0.x) | ((int64_t)w0.y << 32);
      row = w0.z; len = w0.w;
      part = item < n_partials ? (int32_t)item : -1;
    }
    int maxlen = len;  // warp-uniform trip count (groups of one warp hold neighbouring, similar-length units)
#pragma unroll
    for (int o = 16; o >= G; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    float hot_thr = 0.f;
    if (HOT) hot_thr = have ? __ldg(dinv + row) * hot_rsqrt : 0.f;

    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

    // software pipeline: the (column, value) pair of the NEXT step is loaded while this step gathers
    int32_t c_nxt = 0;
    float a_nxt = 0.f;
    if (lig < len) {
      c_nxt = ld_stream_i32(indices + start + lig);
      a_nxt = drop_value(drop, start + lig, ld_stream_f32(values + start + lig));
    }
    for (int base = 0; base < maxlen; base += G) {
      const int32_t c = c_nxt;
      const float a = a_nxt;
      const int kn = base + G + lig;
      c_nxt = 0; a_nxt = 0.f;
      if (kn < len) {
        c_nxt = ld_stream_i32(indices + start + kn);
        a_nxt = drop_value(drop, start + kn, ld_stream_f32(values + start + kn));
      }
      const int cnt = len - base;  // may be <= 0 for the shorter group of the warp
#pragma unroll
      for (int j0 = 0; j0 < G; j0 += UU) {
        float4 x[UU][V];
        float av[UU];
#pragma unroll
        for (int j = 0; j < UU; ++j) {
          const int32_t cj = __shfl_sync(0xffffffffu, c, j0 + j, G);
          av[j] = __shfl_sync(0xffffffffu, a, j0 + j, G);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            x[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int col4 = lig + v * G;
            if (j0 + j < cnt && (EXACT || col4 < d4)) {
              const float* src = X + (int64_t)cj * d + (col4 << 2);
              if (HOT) {
                if (av[j] <= hot_thr) x[j][v] = ld_gather_f4_keep(src); else x[j][v] = ld_gather_f4_stream(src);
              } else {
                x[j][v] = ld_gather_f4(src);
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < UU; ++j) {
#pragma unroll
          for (int v = 0; v < V; ++v) fma4(acc[v], av[j], x[j][v]);
        }
      }
    }
    if (have) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int col4 = lig + v * G;
        if (EXACT || col4 < d4) {
          if (part >= 0) {
            *reinterpret_cast<float4*>(partial + (int64_t)part * d + (col4 << 2)) = acc[v];
          } else {
            epilogue4(acc[v], (int64_t)row * d + (col4 << 2), S_in, Y, S_out, div);
          }
        }
      }
    }
  }
}

// Specialised kernel for d == 4*G*V (16/32/64/128/256): the row stride is a compile-time constant and rows are
// indexed with 32-bit float4 offsets (one IMAD.WIDE per gather instead of a 64-bit multiply chain),
// and steps in which every group of the warp has a full batch of G non-zeros run without any
// predication.  (ncu on the generic kernel: 38 warp instructions per non-zero step, 64% issue-active
// -- the gather loop was instruction-bound, not memory-bound.)
// MODE bit 0: peer output (fused all-gather), bit 1: edge dropout.  Separate instantiations: with runtime
// flags the plain path lost 4% (peer check per row) + 5% (dropout check per index batch).
// MODE bit 2: L2 residency hints for graphs whose embedding table is far larger than L2 (synth-1b: 6 GB of rows, every
// non-zero pulled a 512-byte row from HBM = 36x the algorithmic bytes).  The gather of a HOT column (degree >= the
// graph's hot_deg: value = dinv[row] * dinv[col] <= dinv[row] / sqrt(hot_deg), no lookup needed) is issued with an
// L2 evict_last policy, every other gather and the index / value streams with evict_first, so the ~70 MB of rows that
// take most of the gathers stay in the 126 MB L2 instead of being flushed by the cold stream between two uses.
__device__ __forceinline__ float4 ld_gather_f4_policy(const float4* p, uint64_t pol) {
  float4 r;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
      : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int32_t ld_stream_i32_policy(const int32_t* p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_stream_f32_policy(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_f4_policy(float* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// epilogue4 with every access marked evict_first: the layer's outputs (and the running sum) are streamed once
__device__ __forceinline__ float4 epilogue4_stream(float4 acc, int64_t off, const float* S_in, float* __restrict__ Y,
                                                   float* S_out, float div, uint64_t pol) {
  if (Y) st_f4_policy(Y + off, acc, pol);
  float4 s = acc;
  if (S_out) {
    s = ld_gather_f4_policy(reinterpret_cast<const float4*>(S_in + off), pol);
    s.x += acc.x; s.y += acc.y; s.z += acc.z; s.w += acc.w;
    if (div != 1.0f) {
      s.x = __fdiv_rn(s.x, div); s.y = __fdiv_rn(s.y, div); s.z = __fdiv_rn(s.z, div); s.w = __fdiv_rn(s.w, div);
    }
    st_f4_policy(S_out + off, s, pol);
  }
  return s;
}
struct L2Hint {
  const float* dinv;     // per row of this graph
  float hot_rsqrt;       // 1 / sqrt(hot degree)
};
template <int G, int V, int U, int MINB, int MODE>
__global__ void __launch_bounds__(256, MINB)
k_spmm_fixed(const WorkItem* __restrict__ work, int64_t n_work, int64_t n_partials, const int32_t* __restrict__ indices,
             const float* __restrict__ values, const float* __restrict__ X, const float* S_in,
             float* __restrict__ Y, float* S_out, float* __restrict__ partial, float div, const PeerOut po,
             const DropSpec drop, const L2Hint l2h) {
  constexpr int D4 = G * V;            // float4 per embedding row
  constexpr int D = 4 * D4;
  constexpr int UU = U > G ? G : U;
  uint64_t pol_hot = 0, pol_cold = 0;
  if (MODE & 4) {
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_hot));
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_cold));
  }
  const int lig = threadIdx.x & (G - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t n_rounds = (n_work + n_groups - 1) / n_groups;
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X) + lig;

  for (int64_t round = 0; round < n_rounds; ++round) {
    const int64_t item = round * n_groups + group;
    const int32_t* ci = indices;
    const float* cv = values;
    int64_t start = 0;
    int32_t row = 0, len = 0, part = -1;
    const bool have = item < n_work;
    if (have) {
      const int4 w0 = __ldg(reinterpret_cast<const int4*>(work + item));        // start (lo, hi), row, len
      start = ((int64_t)(uint32_t)w0.x) | ((int64_t)w0.y << 32);
      ci += start; cv += start;
      row = w0.z; len = w0.w;
      part = item < n_partials ? (int32_t)item : -1;
    }
    int maxlen = len;
#pragma unroll
    for (int o = 16; o >= G; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    float hot_thr = 0.f;
    if (MODE & 4) hot_thr = have ? __ldg(l2h.dinv + row) * l2h.hot_rsqrt : 0.f;

    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

    int32_t c_nxt = 0;
    float a_nxt = 0.f;
    if (lig < len) {
      c_nxt = (MODE & 4) ? ld_stream_i32_policy(ci + lig, pol_cold) : ld_stream_i32(ci + lig);
      a_nxt = (MODE & 4) ? ld_stream_f32_policy(cv + lig, pol_cold) : ld_stream_f32(cv + lig);
      if (MODE & 2) a_nxt = drop_value(drop, start + lig, a_nxt);
    }
    for (int base = 0; base < maxlen; base += G) {
      const int32_t c = c_nxt;
      const float a = a_nxt;
      const int kn = base + G + lig;
      c_nxt = 0; a_nxt = 0.f;
      if (kn < len) {
        c_nxt = (MODE & 4) ? ld_stream_i32_policy(ci + kn, pol_cold) : ld_stream_i32(ci + kn);
        a_nxt = (MODE & 4) ? ld_stream_f32_policy(cv + kn, pol_cold) : ld_stream_f32(cv + kn);
        if (MODE & 2) a_nxt = drop_value(drop, start + kn, a_nxt);
      }
      const int cnt = len - base;
      if (__all_sync(0xffffffffu, cnt >= G)) {
        // ---- full batch in every group of the warp: no predicates
#pragma unroll
        for (int j0 = 0; j0 < G; j0 += UU) {
          float4 x[UU][V];
          float av[UU];
#pragma unroll
          for (int j = 0; j < UU; ++j) {
            const uint32_t cj = (uint32_t)__shfl_sync(0xffffffffu, c, j0 + j, G);
            av[j] = __shfl_sync(0xffffffffu, a, j0 + j, G);
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float4* src = X4 + (cj * (uint32_t)D4 + (uint32_t)(v * G));
              if (MODE & 4) x[j][v] = ld_gather_f4_policy(src, av[j] <= hot_thr ? pol_hot : pol_cold);
              else x[j][v] = __ldg(src);
            }
          }
#pragma unroll
          for (int j = 0; j < UU; ++j)
#pragma unroll
            for (int v = 0; v < V; ++v) fma4(acc[v], av[j], x[j][v]);
        }
      } else {
        // ---- ragged tail: a = 0 and column = 0 beyond the row's end (lanes loaded zeros above)
#pragma unroll
        for (int j0 = 0; j0 < G; j0 += UU) {
          if (__all_sync(0xffffffffu, cnt <= j0)) break;
          float4 x[UU][V];
          float av[UU];
#pragma unroll
          for (int j = 0; j < UU; ++j) {
            const uint32_t cj = (uint32_t)__shfl_sync(0xffffffffu, c, j0 + j, G);
            av[j] = __shfl_sync(0xffffffffu, a, j0 + j, G);
#pragma unroll
            for (int v = 0; v < V; ++v) {
              x[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (j0 + j < cnt) {
                const float4* src = X4 + (cj * (uint32_t)D4 + (uint32_t)(v * G));
                if (MODE & 4) x[j][v] = ld_gather_f4_policy(src, av[j] <= hot_thr ? pol_hot : pol_cold);
                else x[j][v] = __ldg(src);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < UU; ++j)
#pragma unroll
            for (int v = 0; v < V; ++v) fma4(acc[v], av[j], x[j][v]);
        }
      }
    }
    if (have) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int col = (lig + v * G) << 2;
        if (part >= 0) {
          *reinterpret_cast<float4*>(partial + (int64_t)part * D + col) = acc[v];
        } else {
          const float4 mean = (MODE & 4) ? epilogue4_stream(acc[v], (int64_t)row * D + col, S_in, Y, S_out, div, pol_cold)
                                         : epilogue4(acc[v], (int64_t)row * D + col, S_in, Y, S_out, div);
          if (MODE & 1) peer_store4(po, row, D, col, acc[v], mean);
        }
      }
    }
  }
}

// Shared-memory staging of the hot rows (north_star: "shared-memory staging of hot rows").
// On L2-resident graphs the kernel above is bound by the L2 -> SM gather path (1.53 GB of 256-byte row gathers per
// Amazon-Book layer for 122 MB of algorithmic bytes, ~12 TB/s = the measured L2 ceiling), so the only way down is to
// gather fewer bytes through L2.  One persistent CTA per SM (1024 threads = the same 32 warps as 4 x 256) copies the H
// most frequently gathered embedding rows (H * d * 4 <= 192 KB; lgx_graph::hot_ids) into shared memory once per layer,
// and the index stream it reads (lgx_graph::hot_idx) stores those columns as ~slot: such gathers are LDS.128 from the
// table instead of LDG.128 through L2.  What it can save is the share of gathers that hit the table (hot_cover):
// 23 % on the Amazon-Book shape (32 % of the item gathers, 12 % of the user gathers), nothing on graphs whose degree
// mass is not concentrated -- the launcher uses it only above 10 % coverage.
constexpr int HOT_THREADS = 1024;
constexpr int HOT_SMEM_BYTES = 192 * 1024;
// Every work unit's entries are stored hot-first (lgx_graph::hot_idx / hot_val / hot_work, built once per table
// size): [n_hot table slots][len - n_hot cold column ids].  The unit is then two plain gather loops, LDS.128 from the
// table and LDG.128 through L2 -- measured alternatives that mix the two per gather lost badly: a branch per gather
// serialises the U gathers that must be in flight together (152 vs 125 us per Amazon-Book layer), a predicated
// LDS/LDG pair per gather was worse still (374 us).
template <int G, int V, int U>
__global__ void __launch_bounds__(HOT_THREADS, 1)
k_spmm_hot(const WorkItem* __restrict__ work, int64_t n_work, int64_t n_partials, const int32_t* __restrict__ hot_idx,
           const float* __restrict__ hot_val, const float* __restrict__ X, const float* S_in,
           float* __restrict__ Y, float* S_out, float* __restrict__ partial, float div,
           const int32_t* __restrict__ hot_ids, int H) {
  constexpr int D4 = G * V;
  constexpr int D = 4 * D4;
  constexpr int UU = U > G ? G : U;
  extern __shared__ float4 tab[];                     // [H][D4]
  const float4* __restrict__ Xall = reinterpret_cast<const float4*>(X);
  for (int i = threadIdx.x; i < H * D4; i += HOT_THREADS) {
    const int slot = i / D4, q = i - slot * D4;
    tab[i] = __ldg(Xall + ((uint32_t)__ldg(hot_ids + slot) * (uint32_t)D4 + (uint32_t)q));
  }
  __syncthreads();
  const int lig = threadIdx.x & (G - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t n_rounds = (n_work + n_groups - 1) / n_groups;
  const float4* __restrict__ X4 = Xall + lig;
  const float4* tabl = tab + lig;

  for (int64_t round = 0; round < n_rounds; ++round) {
    const int64_t item = round * n_groups + group;
    const int32_t* ci = hot_idx;
    const float* cv = hot_val;
    int32_t row = 0, len = 0, n_hot = 0, part = -1;
    const bool have = item < n_work;
    if (have) {
      const int4 w0 = __ldg(reinterpret_cast<const int4*>(work + item));   // start, row, len | n_hot << 16
      const int64_t start = ((int64_t)(uint32_t)w0.x) | ((int64_t)w0.y << 32);
      ci += start; cv += start;
      row = w0.z; len = w0.w & 0xffff; n_hot = (int32_t)((uint32_t)w0.w >> 16);
      part = item < n_partials ? (int32_t)item : -1;
    }
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    // two segments: [0, n_hot) from the table, [n_hot, len) through L2.  Beyond a segment's end lanes hold slot /
    // column 0 with value 0: a harmless gather.
#pragma unroll
    for (int seg = 0; seg < 2; ++seg) {
      const int s0 = seg == 0 ? 0 : n_hot, s1 = seg == 0 ? n_hot : len;
      const int slen = s1 - s0;
      int maxlen = slen;
#pragma unroll
      for (int o = 16; o >= G; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
      int32_t c_nxt = 0;
      float a_nxt = 0.f;
      if (lig < slen) { c_nxt = ld_stream_i32(ci + s0 + lig); a_nxt = ld_stream_f32(cv + s0 + lig); }
      for (int base = 0; base < maxlen; base += G) {
        const int32_t c = c_nxt;
        const float a = a_nxt;
        const int kn = base + G + lig;
        c_nxt = 0; a_nxt = 0.f;
        if (kn < slen) { c_nxt = ld_stream_i32(ci + s0 + kn); a_nxt = ld_stream_f32(cv + s0 + kn); }
        const int cnt = slen - base;
        const bool full = __all_sync(0xffffffffu, cnt >= G);     // every group of the warp has a full batch: no checks
#pragma unroll
        for (int j0 = 0; j0 < G; j0 += UU) {
          if (!full && __all_sync(0xffffffffu, cnt <= j0)) break;
          float4 x[UU][V];
          float av[UU];
#pragma unroll
          for (int j = 0; j < UU; ++j) {
            const uint32_t cj = (uint32_t)__shfl_sync(0xffffffffu, c, j0 + j, G);
            av[j] = __shfl_sync(0xffffffffu, a, j0 + j, G);
#pragma unroll
            for (int v = 0; v < V; ++v) {
              if (seg == 0) x[j][v] = tabl[cj * (uint32_t)D4 + (uint32_t)(v * G)];
              else x[j][v] = __ldg(X4 + (cj * (uint32_t)D4 + (uint32_t)(v * G)));
            }
          }
#pragma unroll
          for (int j = 0; j < UU; ++j)
#pragma unroll
            for (int v = 0; v < V; ++v) fma4(acc[v], av[j], x[j][v]);
        }
      }
    }
    if (have) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int col = (lig + v * G) << 2;
        if (part >= 0) *reinterpret_cast<float4*>(partial + (int64_t)part * D + col) = acc[v];
        else epilogue4(acc[v], (int64_t)row * D + col, S_in, Y, S_out, div);
      }
    }
  }
}

// d not a multiple of 4: scalar lanes (rare; the reference allows any --recdim).
__global__ void __launch_bounds__(256)
k_spmm_scalar(const WorkItem* __restrict__ work, int64_t n_work, int64_t n_partials, const int32_t* __restrict__ indices,
              const float* __restrict__ values, const float* __restrict__ X, const float* S_in,
              float* __restrict__ Y, float* S_out, float* __restrict__ partial, float div, int d, const DropSpec drop) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t item = warp; item < n_work; item += n_warps) {
    const WorkItem w = work[item];
    const int64_t part = item < n_partials ? item : -1;
    for (int c0 = lane; c0 < d; c0 += 32) {
      float acc = 0.f;
      for (int k = 0; k < w.len; ++k) {
        const int32_t c = indices[w.start + k];
        acc = fmaf(drop_value(drop, w.start + k, values[w.start + k]), __ldg(X + (int64_t)c * d + c0), acc);
      }
      if (part >= 0) {
        partial[part * d + c0] = acc;
      } else {
        const int64_t off = (int64_t)w.row * d + c0;
        if (Y) Y[off] = acc;
        if (S_out) {
          float s = S_in[off] + acc;
          S_out[off] = div != 1.0f ? __fdiv_rn(s, div) : s;
        }
      }
    }
  }
}

// Sum the partials of each split row in a fixed order and apply the epilogue.  One CTA per long row:
// column c = tid % cols, slice = tid / cols; every slice sums a contiguous run of partials in order,
// then the slices are combined in order through shared memory (deterministic).
__global__ void __launch_bounds__(256)
k_spmm_long(const LongRow* __restrict__ long_rows, int64_t n_long, const float* __restrict__ partial,
            const float* S_in, float* __restrict__ Y, float* S_out, float div, int d, const PeerOut po) {
  __shared__ float red[256];
  const int colsP = min(d, 256);
  const int slices = 256 / colsP;
  const int cl = threadIdx.x % colsP, slice = threadIdx.x / colsP;
  for (int64_t lr = blockIdx.x; lr < n_long; lr += gridDim.x) {
    const LongRow r = long_rows[lr];
    const int per = (r.n_partials + slices - 1) / slices;
    for (int c_base = 0; c_base < d; c_base += colsP) {
      const int c = c_base + cl;
      float acc = 0.f;
      if (slice < slices && c < d) {
        const int p0 = slice * per, p1 = min(r.n_partials, p0 + per);
        const float* src = partial + (int64_t)(r.first_partial + p0) * d + c;
#pragma unroll 4
        for (int p = p0; p < p1; ++p, src += d) acc += __ldg(src);
      }
      __syncthreads();
      red[threadIdx.x] = acc;
      __syncthreads();
      if (slice == 0 && c < d) {
        acc = red[cl];
        for (int s2 = 1; s2 < slices; ++s2) acc += red[s2 * colsP + cl];
        const int64_t off = (int64_t)r.row * d + c;
        if (Y) Y[off] = acc;
        float s = acc;
        if (S_out) {
          s = S_in[off] + acc;
          s = div != 1.0f ? __fdiv_rn(s, div) : s;
          S_out[off] = s;
        }
        if (po.n > 0) {
          const float v = po.store_mean ? s : acc;
          const int64_t poff = (po.row_offset + r.row) * d + c;
          for (int p = 0; p < po.n; ++p) po.ptr[p][poff] = v;
        }
      }
    }
  }
}

template <typename K>
static int blocks_for(K kernel, int threads) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
  if (per_sm < 1) per_sm = 1;
  return per_sm * sm_count();
}

struct SpmmTuning {
  int variant;      // kernel variant for d = 64 (see spmm_impl)
  int variant128;   // kernel variant for d = 128
  int hot_degree;   // columns with at least this degree are kept in L1 (HOT variants)
};
static SpmmTuning tuning() {
  static SpmmTuning t = [] {
    SpmmTuning x{0, 0, 256};
    if (const char* e = getenv("LGX_SPMM_VARIANT")) x.variant = atoi(e);
    if (const char* e = getenv("LGX_SPMM_VARIANT128")) x.variant128 = atoi(e);
    if (const char* e = getenv("LGX_SPMM_HOT_DEGREE")) x.hot_degree = std::max(1, atoi(e));
    return x;
  }();
  return t;
}

template <int G, int V, bool EXACT, int U, int MINB, bool HOT>
static void launch_spmm(const lgx_graph* g, const float* X, const float* S_in, float* Y, float* S_out, float* partial,
                        float div, int d, cudaStream_t st, const DropSpec& drop) {
  static int max_blocks = 0;
  if (max_blocks == 0) max_blocks = blocks_for(k_spmm<G, V, EXACT, U, MINB, HOT>, 256);
  const int64_t groups_per_block = 256 / G;
  const int64_t need = (g->n_work + groups_per_block - 1) / groups_per_block;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, max_blocks));
  const float hot_rsqrt = 1.0f / sqrtf((float)tuning().hot_degree);
  k_spmm<G, V, EXACT, U, MINB, HOT><<<blocks, 256, 0, st>>>(g->work, g->n_work, g->n_partials, g->indices, g->values,
                                                           X, S_in, Y, S_out, partial, div, d, g->dinv, hot_rsqrt, drop);
}

// L2 hints: only where the table cannot live in L2 anyway (> 2x its size), the graph knows its rows' dinv and
// value == dinv[row] * dinv[col] (LGX_SPMM_L2HINT=0 switches them off, =1 forces them for experiments).
static int hot_degree_for(const lgx_graph* g, int d) {
  const int j = d <= 32 ? 0 : (d <= 64 ? 1 : (d <= 128 ? 2 : (d <= 256 ? 3 : 4)));
  return g->hot_deg[j];
}
static bool use_l2_hints(const lgx_graph* g, int d) {
  static const int forced = [] { const char* e = getenv("LGX_SPMM_L2HINT"); return e ? atoi(e) : -1; }();
  if (forced == 0 || !g->values_are_dinv_products || g->dinv == nullptr || hot_degree_for(g, d) <= 0) return false;
  if (forced == 1) return true;
  return (double)g->n_cols * d * 4.0 > 2.0 * 126e6;
}

template <int G, int V, int U, int MINB, int MODE>
static void launch_fixed_t(const lgx_graph* g, const float* X, const float* S_in, float* Y, float* S_out, float* partial,
                           float div, cudaStream_t st, const PeerOut& po, const DropSpec& drop) {
  static int max_blocks = 0;
  if (max_blocks == 0) max_blocks = blocks_for(k_spmm_fixed<G, V, U, MINB, MODE>, 256);
  const int64_t groups_per_block = 256 / G;
  const int64_t need = (g->n_work + groups_per_block - 1) / groups_per_block;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, max_blocks));
  L2Hint l2h{g->dinv, 0.f};
  if (MODE & 4) {
    l2h.hot_rsqrt = 1.0f / sqrtf((float)std::max(1, hot_degree_for(g, 4 * G * V)));
    // Optional persisting carve-out (LGX_SPMM_L2_PERSIST_MB, default off): measured on the 1B-edge graph it HURTS
    // (428 ms for 3 layers with 80 MB set aside vs 405 ms with the hints alone and 422 ms without hints) -- the
    // carve-out takes L2 away from the cold stream without keeping more of the hot rows.
    static bool carved[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < kMaxDevices && !carved[dev]) {
      int max_persist = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
      static const double want_mb = [] { const char* e = getenv("LGX_SPMM_L2_PERSIST_MB"); return e ? atof(e) : 0.0; }();
      const size_t want = (size_t)std::min<double>((double)max_persist, want_mb * 1e6);
      if (want > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaGetLastError();
      carved[dev] = true;
    }
  }
  k_spmm_fixed<G, V, U, MINB, MODE><<<blocks, 256, 0, st>>>(g->work, g->n_work, g->n_partials, g->indices, g->values, X,
                                                           S_in, Y, S_out, partial, div, po, drop, l2h);
}
template <int G, int V, int U, int MINB>
static void launch_fixed(const lgx_graph* g, const float* X, const float* S_in, float* Y, float* S_out, float* partial,
                         float div, cudaStream_t st, const PeerOut& po, const DropSpec& drop) {
  const int mode = (po.n > 0 ? 1 : 0) | (drop.enabled ? 2 : 0);
  if (!drop.enabled && use_l2_hints(g, 4 * G * V)) {
    if (mode == 0) launch_fixed_t<G, V, U, MINB, 4>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
    else launch_fixed_t<G, V, U, MINB, 5>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
    return;
  }
  if (mode == 0) launch_fixed_t<G, V, U, MINB, 0>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
  else if (mode == 1) launch_fixed_t<G, V, U, MINB, 1>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
  else if (mode == 2) launch_fixed_t<G, V, U, MINB, 2>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
  else launch_fixed_t<G, V, U, MINB, 3>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
}

// Table size for width d.  OPT-IN (LGX_SPMM_HOT=1, read at graph build and here): measured on B200 at the Amazon-Book
// shape the staged kernel is SLOWER than the plain one -- 139 us per layer (table 192 KB; 137 / 135 us with 96 / 48 KB)
// against 125 us -- although 23 % of its gathers are served from shared memory: the plain kernel is bound by
// instruction issue and L2 latency per gather, not by L2 bytes, and the two-segment unit loop adds a ragged batch
// and the persistent 1024-thread CTA loses the 4 x 256 launch's finer tail.  Kept for graphs with a heavier head.
static int hot_table_rows(const lgx_graph* g, int d) {
  if (!g->hot_ids || (d != 64 && d != 128) || g->chunk_nnz > 32767) return 0;
  static const int forced = [] { const char* e = getenv("LGX_SPMM_HOT"); return e ? atoi(e) : 0; }();
  if (forced != 1) return 0;
  static const int budget = [] { const char* e = getenv("LGX_SPMM_HOT_KB"); return e ? atoi(e) * 1024 : HOT_SMEM_BYTES; }();
  int j = 0;
  while (j < 5 && kHotSteps[j + 1] * d * 4 <= std::min(budget, HOT_SMEM_BYTES)) ++j;      // largest step that fits (192 KB)
  const int h = std::min(kHotSteps[j], (int)g->n_hot);
  if (h < kHotSteps[j]) return 0;
  const double cover = g->nnz > 0 ? (double)g->hot_cover[j] / (double)g->nnz : 0.0;
  (void)cover;
  return h;
}

template <int G, int V, int U>
static int launch_hot(const lgx_graph* g, int h, const float* X, const float* S_in, float* Y, float* S_out, float* partial,
                      float div, cudaStream_t st) {
  const int rc = ensure_hot_index(g, h, st);
  if (rc != LGX_OK) return rc;
  constexpr int D = 4 * G * V;
  const size_t smem = (size_t)h * D * sizeof(float);
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (dev >= kMaxDevices || !configured[dev]) {
    LGX_CHECK_CUDA(cudaFuncSetAttribute(k_spmm_hot<G, V, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, HOT_SMEM_BYTES));
    if (dev < kMaxDevices) configured[dev] = true;
  }
  const int64_t groups_per_block = HOT_THREADS / G;
  const int64_t need = (g->n_work + groups_per_block - 1) / groups_per_block;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, sm_count()));
  k_spmm_hot<G, V, U><<<blocks, HOT_THREADS, smem, st>>>(g->hot_work, g->n_work, g->n_partials, g->hot_idx, g->hot_val, X,
                                                         S_in, Y, S_out, partial, div, g->hot_ids, h);
  return LGX_OK;
}

static int spmm_impl(const lgx_graph* g, const float* X, const float* S_in, float* Y, float* S_out, float div,
                     int32_t d, void* workspace, cudaStream_t st, const PeerOut& po = PeerOut{},
                     const DropSpec& drop = DropSpec{}) {
  float* partial = reinterpret_cast<float*>(workspace);
  const bool fixed_ok = (d == 16 || d == 32 || d == 64 || d == 128 || d == 256) &&
                        g->n_cols * (int64_t)(d / 4) < ((int64_t)1 << 32);
  if (po.n > 0 && !fixed_ok) {
    set_error("fused peer output needs d in {16,32,64,128,256}");
    return LGX_ERR_INVALID;
  }
  if (g->n_work > 0) {
    const bool hot_ok = g->values_are_dinv_products;   // HOT needs value == dinv[row]*dinv[col]
    if (d % 4 != 0) {
      const int64_t need = (g->n_work + 7) / 8;
      const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)sm_count() * 8));
      k_spmm_scalar<<<blocks, 256, 0, st>>>(g->work, g->n_work, g->n_partials, g->indices, g->values, X, S_in, Y, S_out, partial, div, d, drop);
    } else if (tuning().variant == 0 && po.n == 0 && !drop.enabled && fixed_ok && hot_table_rows(g, d) > 0) {
      // hot rows staged in shared memory (plain propagation, forward and backward)
      const int h = hot_table_rows(g, d);
      const int rc = d == 64 ? launch_hot<16, 1, 4>(g, h, X, S_in, Y, S_out, partial, div, st)
                             : launch_hot<32, 1, 4>(g, h, X, S_in, Y, S_out, partial, div, st);
      if (rc != LGX_OK) return rc;
    } else if ((tuning().variant == 0 || po.n > 0) && fixed_ok) {
      // default: compile-time row stride, 4 gathers in flight per lane, >= 4 CTAs per SM
      // (B200 sweep over U x occupancy at Amazon-Book shape: profiles/r1_spmm_sweep.txt)
      if (d == 64) launch_fixed<16, 1, 4, 4>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
      else if (d == 128) {
        switch (tuning().variant128) {   // LGX_SPMM_VARIANT128: experiment knob for d = 128
          case 1: launch_fixed<32, 1, 8, 3>(g, X, S_in, Y, S_out, partial, div, st, po, drop); break;
          case 2: launch_fixed<32, 1, 8, 2>(g, X, S_in, Y, S_out, partial, div, st, po, drop); break;
          case 3: launch_fixed<32, 1, 4, 5>(g, X, S_in, Y, S_out, partial, div, st, po, drop); break;
          case 4: launch_fixed<32, 1, 2, 6>(g, X, S_in, Y, S_out, partial, div, st, po, drop); break;
          case 5: launch_fixed<32, 1, 16, 2>(g, X, S_in, Y, S_out, partial, div, st, po, drop); break;
          case 6: launch_fixed<32, 1, 8, 4>(g, X, S_in, Y, S_out, partial, div, st, po, drop); break;
          default: launch_fixed<32, 1, 4, 4>(g, X, S_in, Y, S_out, partial, div, st, po, drop); break;
        }
      }
      else if (d == 256) launch_fixed<32, 2, 4, 3>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
      else if (d == 32) launch_fixed<8, 1, 4, 4>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
      else launch_fixed<4, 1, 4, 4>(g, X, S_in, Y, S_out, partial, div, st, po, drop);
    } else if (d == 64) {
      switch (tuning().variant) {   // experiment knob LGX_SPMM_VARIANT (scripts/spmm_sweep.py); 0 = default above
        case 1: launch_spmm<16, 1, true, 8, 3, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop); break;
        case 3: launch_spmm<16, 1, true, 16, 1, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop); break;
        case 5: if (hot_ok) { launch_spmm<16, 1, true, 8, 2, true>(g, X, S_in, Y, S_out, partial, div, d, st, drop); break; }
        case 6: if (hot_ok) { launch_spmm<16, 1, true, 8, 3, true>(g, X, S_in, Y, S_out, partial, div, d, st, drop); break; }
        case 7: if (hot_ok) { launch_spmm<16, 1, true, 4, 4, true>(g, X, S_in, Y, S_out, partial, div, d, st, drop); break; }
        case 12: launch_spmm<16, 1, true, 8, 2, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop); break;
        case 20: launch_fixed<16, 1, 8, 2>(g, X, S_in, Y, S_out, partial, div, st, PeerOut{}, drop); break;
        case 21: launch_fixed<16, 1, 8, 3>(g, X, S_in, Y, S_out, partial, div, st, PeerOut{}, drop); break;
        case 24: launch_fixed<16, 1, 8, 4>(g, X, S_in, Y, S_out, partial, div, st, PeerOut{}, drop); break;
        case 26: launch_fixed<16, 1, 4, 5>(g, X, S_in, Y, S_out, partial, div, st, PeerOut{}, drop); break;
        default: launch_spmm<16, 1, true, 4, 4, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop); break;
      }
    } else if (d == 128) {
      launch_spmm<32, 1, true, 8, 2, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop);
    } else if (d == 256) {
      launch_spmm<32, 2, true, 8, 1, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop);
    } else if (d == 32) {
      launch_spmm<8, 1, true, 8, 2, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop);
    } else if (d == 16) {
      launch_spmm<4, 1, true, 4, 2, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop);
    } else if (d <= 128) {
      launch_spmm<32, 1, false, 8, 2, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop);
    } else if (d <= 256) {
      launch_spmm<32, 2, false, 8, 1, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop);
    } else {
      launch_spmm<32, 4, false, 4, 1, false>(g, X, S_in, Y, S_out, partial, div, d, st, drop);
    }
    LGX_CHECK_LAUNCH();
  }
  if (g->n_long > 0) {
    const int blocks = (int)std::min<int64_t>(g->n_long, (int64_t)sm_count() * 8);
    k_spmm_long<<<blocks, 256, 0, st>>>(g->long_rows, g->n_long, partial, S_in, Y, S_out, div, d, po);
    LGX_CHECK_LAUNCH();
  }
  return LGX_OK;
}

// Cross-GPU barrier through peer memory: every rank owns an array of n_peers 32-bit flags that the other ranks map.
// Rank `self` writes `epoch` into slot [self] of every rank's array (release, system scope) and then waits until all
// slots of its OWN array have reached `epoch` (acquire).  It runs after the layer's SpMM on the same stream, so that
// kernel's peer stores are complete when the flags go out.  Replaces a host-launched NCCL all-reduce of one float per
// layer (~25-40 us) with one ~5 us kernel.  The wait is bounded: a rank that never arrives traps instead of hanging.
struct PeerFlags {
  uint32_t* ptr[kMaxPeers];
};
__global__ void k_peer_barrier(const PeerFlags pf, int n_peers, int self, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r >= n_peers) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.ptr[r] + self), "r"(epoch) : "memory");
  const uint32_t* mine = pf.ptr[self] + r;
  const long long t0 = clock64();
  uint32_t polls = 0;
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;                     // epochs only grow (wrap-safe compare)
    if ((++polls & 1023u) == 0 && clock64() - t0 > 20000000000LL) __trap();   // ~10 s
  }
  __threadfence_system();
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace lgx

using namespace lgx;

extern "C" {

size_t lgx_spmm_workspace_bytes(const lgx_graph* g, int32_t d) {
  if (!g || d <= 0) return 0;
  return align256((size_t)g->n_partials * (size_t)d * sizeof(float));
}

int lgx_spmm(const lgx_graph* g, const float* X, const float* S_in, float* Y, float* S_out, float div, int32_t d,
             void* workspace, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(g && X, "graph or X is NULL");
  LGX_REQUIRE(d > 0 && d <= 512, "d must be in [1, 512]");
  LGX_REQUIRE(Y || S_out, "nothing to write: Y and S_out are both NULL");
  LGX_REQUIRE(!S_out || S_in, "S_out needs S_in");
  LGX_REQUIRE(g->n_partials == 0 || workspace, "graph has split rows: workspace required");
  LGX_REQUIRE(div != 0.0f, "div must be non-zero");
  return spmm_impl(g, X, S_in, Y, S_out, div, d, workspace, (cudaStream_t)stream);
}

int lgx_spmm_peers(const lgx_graph* g, const float* X, const float* S_in, float* const* peers_host, int32_t n_peers,
                   int64_t row_offset, int32_t store_mean, float* S_out, float div, int32_t d, void* workspace,
                   lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(g && X && peers_host, "NULL argument");
  LGX_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers, "n_peers must be in [1, 16]");
  LGX_REQUIRE(d > 0 && d <= 512, "d must be in [1, 512]");
  LGX_REQUIRE(!S_out || S_in, "S_out needs S_in");
  LGX_REQUIRE(!store_mean || S_out, "store_mean needs S_out");
  LGX_REQUIRE(g->n_partials == 0 || workspace, "graph has split rows: workspace required");
  LGX_REQUIRE(div != 0.0f && row_offset >= 0, "bad div / row_offset");
  PeerOut po{};
  for (int p = 0; p < n_peers; ++p) {
    LGX_REQUIRE(peers_host[p] != nullptr, "NULL peer pointer");
    po.ptr[p] = peers_host[p];
  }
  po.n = n_peers; po.store_mean = store_mean; po.row_offset = row_offset;
  return spmm_impl(g, X, S_in, nullptr, S_out, div, d, workspace, (cudaStream_t)stream, po);
}

int lgx_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(ptr && handle64 && bytes > 0, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  LGX_CHECK_CUDA(cudaMalloc(ptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) {
    cudaFree(*ptr);
    *ptr = nullptr;
    set_error(std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(e));
    return LGX_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  return LGX_OK;
}

int lgx_peer_open(const unsigned char* handle64, void** ptr) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(ptr && handle64, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  LGX_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return LGX_OK;
}

int lgx_peer_copy(void* const* peers_host, int32_t n_peers, int32_t self, size_t offset_bytes, size_t bytes,
                  lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(peers_host && n_peers >= 1 && n_peers <= kMaxPeers && self >= 0 && self < n_peers, "bad peer list");
  if (bytes == 0) return LGX_OK;
  const char* src = reinterpret_cast<const char*>(peers_host[self]) + offset_bytes;
  for (int k = 1; k < n_peers; ++k) {             // staggered order: rank r starts with r+1 (no hot receiver)
    const int p = (self + k) % n_peers;
    char* dst = reinterpret_cast<char*>(peers_host[p]) + offset_bytes;
    LGX_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));   // one plain async copy per peer
  }
  return LGX_OK;
}

int lgx_peer_barrier(void* const* flag_peers_host, int32_t n_peers, int32_t self, uint32_t epoch, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(flag_peers_host && n_peers >= 1 && n_peers <= kMaxPeers && self >= 0 && self < n_peers, "bad peer list");
  LGX_REQUIRE(epoch != 0, "epoch 0 is the initial state of the flags");
  PeerFlags pf{};
  for (int p = 0; p < n_peers; ++p) {
    LGX_REQUIRE(flag_peers_host[p] != nullptr, "NULL peer pointer");
    pf.ptr[p] = reinterpret_cast<uint32_t*>(flag_peers_host[p]);
  }
  k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(pf, n_peers, self, epoch);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_peer_close(void* ptr) {
  if (ptr) LGX_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return LGX_OK;
}

int lgx_peer_free(void* ptr) {
  if (ptr) LGX_CHECK_CUDA(cudaFree(ptr));
  return LGX_OK;
}

size_t lgx_propagate_workspace_bytes(const lgx_graph* g, int32_t d, int32_t n_layers) {
  if (!g || d <= 0) return 0;
  (void)n_layers;
  const size_t layer = align256((size_t)g->n_rows * (size_t)d * sizeof(float));
  return 2 * layer + lgx_spmm_workspace_bytes(g, d);
}

static int make_drop(const lgx_graph* g, float keep_prob, uint64_t seed, bool transpose, DropSpec* out);

static int propagate_fwd_impl(const lgx_graph* g, const float* E0, float* out_mean, float* layers_out, int32_t n_layers,
                              int32_t d, void* workspace, lgx_stream stream, const DropSpec& drop) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(g && E0 && out_mean, "NULL argument");
  LGX_REQUIRE(g->n_rows == g->n_cols, "propagate needs a square graph (use lgx_spmm for row shards)");
  LGX_REQUIRE(n_layers >= 0 && n_layers <= 64, "n_layers out of range");
  LGX_REQUIRE(d > 0 && d <= 512, "d must be in [1, 512]");
  LGX_REQUIRE(workspace || n_layers == 0, "workspace is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_el = (size_t)g->n_rows * d;
  if (n_layers == 0) {
    LGX_CHECK_CUDA(cudaMemcpyAsync(out_mean, E0, n_el * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return LGX_OK;
  }
  const size_t layer = align256(n_el * sizeof(float));
  float* buf[2] = {reinterpret_cast<float*>(workspace), reinterpret_cast<float*>((char*)workspace + layer)};
  void* spmm_ws = (char*)workspace + 2 * layer;
  const float* X = E0;
  const float* S_in = E0;
  for (int l = 1; l <= n_layers; ++l) {
    const bool last = l == n_layers;
    float* Y = layers_out ? layers_out + (size_t)(l - 1) * n_el : (last ? nullptr : buf[(l - 1) & 1]);
    const float div = last ? (float)(n_layers + 1) : 1.0f;  // torch.mean = sum / (L+1), PT/model.py:175
    int rc = spmm_impl(g, X, S_in, Y, out_mean, div, d, spmm_ws, st, PeerOut{}, drop);
    if (rc != LGX_OK) return rc;
    X = Y;
    S_in = out_mean;
  }
  return LGX_OK;
}

int lgx_propagate_fwd(const lgx_graph* g, const float* E0, float* out_mean, float* layers_out, int32_t n_layers,
                      int32_t d, void* workspace, lgx_stream stream) {
  return propagate_fwd_impl(g, E0, out_mean, layers_out, n_layers, d, workspace, stream, DropSpec{});
}

int lgx_propagate_fwd_dropout(const lgx_graph* g, const float* E0, float* out_mean, int32_t n_layers, int32_t d,
                              float keep_prob, uint64_t seed, void* workspace, lgx_stream stream) {
  LGX_REQUIRE(g, "graph is NULL");
  DropSpec drop{};
  int rc = make_drop(g, keep_prob, seed, false, &drop);
  if (rc != LGX_OK) return rc;
  return propagate_fwd_impl(g, E0, out_mean, nullptr, n_layers, d, workspace, stream, drop);
}

static int make_drop(const lgx_graph* g, float keep_prob, uint64_t seed, bool transpose, DropSpec* out) {
  LGX_REQUIRE(keep_prob > 0.0f && keep_prob <= 1.0f, "keep_prob must be in (0, 1]");
  LGX_REQUIRE(!transpose || g->tpos, "call lgx_graph_enable_dropout before the dropout backward");
  out->tpos = transpose ? g->tpos : nullptr;
  out->seed = seed; out->keep_prob = keep_prob; out->inv_keep = 1.0f / keep_prob; out->enabled = 1;
  return LGX_OK;
}

static int propagate_bwd_impl(const lgx_graph* g, const float* g_scaled, float* dE0, int32_t n_layers, int32_t d,
                              void* workspace, lgx_stream stream, const DropSpec& drop) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(g && g_scaled && dE0, "NULL argument");
  LGX_REQUIRE(g->n_rows == g->n_cols, "propagate needs a square graph");
  LGX_REQUIRE(n_layers >= 0 && n_layers <= 64, "n_layers out of range");
  LGX_REQUIRE(d > 0 && d <= 512, "d must be in [1, 512]");
  LGX_REQUIRE(workspace || n_layers == 0, "workspace is NULL");
  LGX_REQUIRE(g_scaled != dE0 || n_layers == 0, "dE0 must not alias g_scaled");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_el = (size_t)g->n_rows * d;
  if (n_layers == 0) {
    if (dE0 != g_scaled)
      LGX_CHECK_CUDA(cudaMemcpyAsync(dE0, g_scaled, n_el * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return LGX_OK;
  }
  const size_t layer = align256(n_el * sizeof(float));
  float* buf[2] = {reinterpret_cast<float*>(workspace), reinterpret_cast<float*>((char*)workspace + layer)};
  void* spmm_ws = (char*)workspace + 2 * layer;
  // Horner: t_L = g;  t_{k-1} = g + A t_k;  dE0 = t_0   (A_hat symmetric => A^T = A)
  const float* t = g_scaled;
  for (int l = n_layers; l >= 1; --l) {
    float* t_new = (l == 1) ? dE0 : buf[l & 1];
    int rc = spmm_impl(g, t, g_scaled, nullptr, t_new, 1.0f, d, spmm_ws, st, PeerOut{}, drop);
    if (rc != LGX_OK) return rc;
    t = t_new;
  }
  return LGX_OK;
}

int lgx_propagate_bwd(const lgx_graph* g, const float* g_scaled, float* dE0, int32_t n_layers, int32_t d,
                      void* workspace, lgx_stream stream) {
  return propagate_bwd_impl(g, g_scaled, dE0, n_layers, d, workspace, stream, DropSpec{});
}

// Backward through the DROPPED graph of the same step: dX = (A_drop)^T dY, and entry (i, j) of the
// transpose carries the keep decision of entry (j, i) -> hash of the mirrored position.
int lgx_propagate_bwd_dropout(const lgx_graph* g, const float* g_scaled, float* dE0, int32_t n_layers, int32_t d,
                              float keep_prob, uint64_t seed, void* workspace, lgx_stream stream) {
  LGX_REQUIRE(g, "graph is NULL");
  DropSpec drop{};
  int rc = make_drop(g, keep_prob, seed, true, &drop);
  if (rc != LGX_OK) return rc;
  return propagate_bwd_impl(g, g_scaled, dE0, n_layers, d, workspace, stream, drop);
}

}  // extern "C"
