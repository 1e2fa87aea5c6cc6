// Fused full-catalogue scoring + train mask + top-K on the 5th-gen tensor cores (sm_100a).
//
// Replaces getUsersRating (PT/model.py:179-184: SGEMM [B,d]x[d,M] + sigmoid), the train-item mask
// (PT/Procedure.py:129-134), torch.topk (:135) and the dead 36 MB/batch score-matrix D2H (:136).
// The [B, M] score matrix lives only in TMEM: it never reaches shared memory, L2 or HBM.
//
// One CTA = 128 users (UMMA M, one TMEM lane per user) x one split of the item catalogue.
//   warp 0      TMA producer: user tile once (resident in smem), then item K-blocks [256 x 64] bf16
//               through an mbarrier ring (cp.async.bulk.tensor, SWIZZLE_128B)
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::f16 (M128 N256 K16), fp32
//               accumulators double-buffered in TMEM (2 x 256 columns), tcgen05.commit -> mbarriers
//   warp 2      TMEM allocator / deallocator
//   warps 4-11  epilogue: each thread owns one user row (TMEM lane) and one 128-column half of the
//               tile: tcgen05.ld 32 columns at a time (double-buffered in registers), train items of
//               the row overwritten with -inf by a monotone cursor over the user's sorted CSR row,
//               3-input-max filter over 4 groups of 8 columns against the row's bound, survivors
//               appended to a shared-memory queue, queues flushed warp-convergently into a
//               register-resident sorted top-K list.  The bound is the best of: the thread's own
//               K-th score, the partner half's published bound, max_j min(a_j, b_{K-j}) over both
//               halves' quartile ranks, and (when the catalogue is split over CTAs) the row's bound
//               shared through global memory.  Operand rows past the end are NaN-filled by TMA, so
//               out-of-range columns and users need no code.
// bf16x3 mode runs the same kernel with K = 3d over split operands (hi.hi + hi.lo + lo.hi).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "lgx_common.cuh"
#include "lgx_score_plan.cuh"
#include "lgx_topk.cuh"
#include "lgx_tc_ptx.cuh"

namespace lgx {

int launch_merge_i32(const float* ws_val, const int32_t* ws_idx, int P, int B, int K, int64_t item_offset,
                     int64_t m_local, TrainMask mask, const int64_t* users, int64_t* out_idx, float* out_val,
                     cudaStream_t st);
TrainMask make_mask(const lgx_graph* g);
int score_topk_fp32(const lgx_graph* g, const float* U, const int64_t* users, int B, const float* I, int M, int d,
                    int K, int64_t item_offset, int64_t* out_idx, float* out_val, void* workspace, cudaStream_t st);
// the group-queue kernel (lgx_score_gq.cu): the default tcgen05 path
struct GqConfig { int k_blocks, stages, q_cap, union_bound; size_t smem; bool ok; };
GqConfig gq_config(int d, int K, int mode);
ScorePlan gq_plan(int B, int M, int sms);
int score_topk_gq(const lgx_graph* g, const void* U_op, const int64_t* users, int B, const void* I_op, int M, int d,
                  int K, int mode, int64_t item_offset, int64_t* out_idx, float* out_val, void* workspace,
                  cudaStream_t st);
// LGX_SCORE_KERNEL=2 selects the first tcgen05 kernel of this file (per-column candidates, mask cursor in the
// epilogue) for A/B runs; the default is the group-queue kernel.
static bool use_gq(int d, int K, int mode) {
  static const int v = [] { const char* e = std::getenv("LGX_SCORE_KERNEL"); return e ? std::atoi(e) : 3; }();
  return v != 2 && gq_config(d, K, mode).ok;
}

constexpr int TC_TILE_U = 128;                 // UMMA M
constexpr int TC_KBLK = 64;                    // bf16 elements per K-block = one 128-byte swizzle row
constexpr int TC_A_BLOCK_BYTES = TC_TILE_U * TC_KBLK * 2;   // 16 KB
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_TMEM_BUF = 256;               // column stride between the two accumulator buffers

// Tile geometry.  NSEG = column segments per tile = epilogue warps per SM sub-partition:
//   NSEG 2: N = 256, 8 epilogue warps, each thread owns 128 columns (4 chunks, double-buffered TMEM loads)
//   NSEG 3: N = 192, 12 epilogue warps, each thread owns 64 columns (2 chunks) -- three warps per
//           scheduler hide each other's TMEM-load and insert latencies; 128-register budget per thread.
template <int NSEG>
struct TcGeo {
  static constexpr int TILE_I = NSEG == 2 ? 256 : 192;          // UMMA N
  static constexpr int EPI = NSEG * TC_TILE_U;                   // epilogue threads
  static constexpr int THREADS = 128 + EPI;
  static constexpr int SEG_COLS = TILE_I / NSEG;
  static constexpr int STAGE_BYTES = TILE_I * TC_KBLK * 2;      // 32 KB / 24 KB
  // kind::f16 instruction descriptor: D=F32, A=B=BF16, both K-major, N=TILE_I, M=128.
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_I >> 3) << 17) |
                                    ((uint32_t)(TC_TILE_U >> 4) << 24);
};
constexpr int TC_SMEM_LIMIT = 232448;          // 227 KB opt-in limit per CTA


struct TcParams {
  int B, M, K;            // users in batch, local items, top-K
  int k_blocks, stages;   // 64-wide K blocks per tile, B-operand ring depth
  int q_cap;              // per-thread candidate queue capacity (multiple of 8)
  int union_bound;        // row threshold from both halves' quartile ranks; the quartile array exists only then
  int n_splits, tiles_per_split;
  int64_t item_offset;
  TrainMask mask;
  const int64_t* users;   // global user id per batch row (mask lookup) or NULL
  float* ws_val;          // [n_splits, B, K]
  int32_t* ws_idx;
  unsigned* row_bound;    // [B] ordered-uint encoding of each row's best known lower bound on its final K-th
                          // score, shared by every CTA that works on the row (NULL when n_splits == 1)
};

// float <-> unsigned with the same ordering (for atomicMax); 0 sorts below every float
__device__ __forceinline__ unsigned ord_encode(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_decode(unsigned u) {
  return u == 0u ? -CUDART_INF_F : __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Monotone cursor over one user's sorted train-item list (the user row of the bipartite CSR).
// Item tiles are visited in ascending order, so "is this column a train item" is a compare against
// the next pending train item instead of a binary search per candidate.
struct TrainCursor {
  const int32_t* idx;   // CSR column array
  int64_t cur, end;
  int32_t bias;         // n_users + item_offset: column value of local item 0
  int32_t next;         // local id of the next train item, INT32_MAX when exhausted
  __device__ __forceinline__ void init(const TrainMask& m, int64_t uid, int64_t item_offset, int first_local) {
    idx = m.indices; cur = end = 0; next = INT32_MAX; bias = 0;
    if (m.indptr == nullptr || uid < 0) return;
    bias = (int32_t)(m.n_users + item_offset);
    int64_t lo = m.indptr[uid], hi = m.indptr[uid + 1];
    end = hi;
    const int32_t key = bias + first_local;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(idx + mid) < key) lo = mid + 1; else hi = mid; }
    cur = lo;
    next = cur < end ? __ldg(idx + cur) - bias : INT32_MAX;
  }
  __device__ __forceinline__ void advance() {
    ++cur;
    next = cur < end ? __ldg(idx + cur) - bias : INT32_MAX;
  }
};

// Per-thread epilogue state.
//  * the row's top-K list lives in REGISTERS (KMAX slots, sorted best-first, right-aligned: slots
//    [0, KMAX-K) hold +inf sentinels so the K-th best is always the last slot).  One insert is a
//    branch-free sweep in which every slot is computed from the OLD neighbours -- no dependency chain,
//    no shared-memory latency (the shared-memory list version spent 30% of the epilogue there);
//  * candidates are only APPENDED (2 stores to a shared-memory queue) inside divergent code; the 32
//    lanes' queues are merged into the lists in a warp-convergent flush, so inserts run in lockstep.
template <int KMAX, int NSEG, bool SHARE>
struct EpiState {
  static constexpr int EPI = TcGeo<NSEG>::EPI;
  float lv[KMAX];
  int32_t li[KMAX];
  // Candidate queue: q_cap rows of [EPI scores | EPI item ids]; this thread owns one column of each, so an
  // append is two stores off one pointer (the id at a constant offset) and one add.
  float* qhead; float* qbase;   // next free slot / first slot of the score column
  float* thr_mine;              // shared memory: largest float below my row-threshold bound (published at every flush)
  const float* thr_other;       // the same from the thread(s) owning the other column segment(s) of this row
  const float* thr_other2;
  float4* quart_mine;           // NSEG 2, K == KMAX: my list's ranks K/4, K/2, 3K/4, K (published at every flush)
  const float4* quart_other;
  float tu;                     // what I last published to thr_mine
  bool use_union;
  unsigned* gbound;             // this row's slot of TcParams::row_bound, or NULL
  // Filter threshold: a score must beat my own K-th best, and must be >= the partner's K-th best -- the
  // partner already holds K items of this row at least that good, so anything below it cannot reach the
  // row's final top-K (equal scores are kept: the final merge breaks ties by item id).
  __device__ __forceinline__ float filter() const {
    if (NSEG == 2) return max3(lv[KMAX - 1], tu, *reinterpret_cast<const volatile float*>(thr_other));
    return max3(lv[KMAX - 1], *reinterpret_cast<const volatile float*>(thr_other),
                *reinterpret_cast<const volatile float*>(thr_other2));
  }
  __device__ __forceinline__ void init(int K) {
#pragma unroll
    for (int p = 0; p < KMAX; ++p) {
      lv[p] = p < KMAX - K ? CUDART_INF_F : -CUDART_INF_F;
      li[p] = INT32_MAX;
    }
    tu = -CUDART_INF_F;
  }
  static constexpr int ROW = EPI * 2;        // queue row stride in 4-byte words
  __device__ __forceinline__ void append(float s, int32_t j) {
    qhead[0] = s;
    reinterpret_cast<int32_t*>(qhead)[EPI] = j;
    qhead += ROW;
  }
  __device__ __forceinline__ bool fuller_than(int rows) const { return qhead > qbase + rows * ROW; }
  // Items reach a thread in ascending id order, so an equal score always loses the tie: strict '>'.
  __device__ __forceinline__ void insert(float x, int32_t xi) {
#pragma unroll
    for (int p = KMAX - 1; p >= 1; --p) {
      const bool gp = x > lv[p], gq = x > lv[p - 1];
      lv[p] = gp ? (gq ? lv[p - 1] : x) : lv[p];
      li[p] = gp ? (gq ? li[p - 1] : xi) : li[p];
    }
    const bool g0 = x > lv[0];
    lv[0] = g0 ? x : lv[0];
    li[0] = g0 ? xi : li[0];
  }
  // Row-level bound from both halves' lists (a = mine, b = partner's, ranks best-first): for any j the row
  // already holds j + (K - j) = K items scoring >= min(a_j, b_{K-j}), so the row's final K-th best is at least
  // max_j min(a_j, b_{K-j}) -- for two similar halves about the (K/2)-th best of each, far tighter than
  // max(a_K, b_K).  A torn read of the partner's quartiles is harmless: every component only ever rises and
  // each old value is itself a valid rank bound.
  __device__ __forceinline__ void flush() {
    // other CTAs' bound for this row: loaded first, consumed after the inserts
    const unsigned gb = (SHARE && gbound) ? *reinterpret_cast<const volatile unsigned*>(gbound) : 0u;
    for (const float* a = qbase; a < qhead; a += ROW) insert(a[0], reinterpret_cast<const int32_t*>(a)[EPI]);
    qhead = qbase;
    float t = lv[KMAX - 1];
    if (NSEG == 2 && use_union) {
      const float a1 = lv[KMAX / 4 - 1], a2 = lv[KMAX / 2 - 1], a3 = lv[3 * KMAX / 4 - 1];
      const volatile float4* o = quart_other;
      const float b1 = o->x, b2 = o->y, b3 = o->z, b4 = o->w;
      *quart_mine = make_float4(a1, a2, a3, t);
      t = fmaxf(max3(fminf(a1, b3), fminf(a2, b2), fminf(a3, b1)), fmaxf(t, b4));
    }
    if (SHARE && gbound) {
      const unsigned mine = (t != t) ? 0u : ord_encode(t);
      if (mine > gb) atomicMax(gbound, mine);
      t = fmaxf(t, ord_decode(gb));
    }
    const int tb = __float_as_int(t);                  // publish prev_float(bound); -inf stays -inf
    tu = (t == -CUDART_INF_F || t != t) ? -CUDART_INF_F : __int_as_float(tb > 0 ? tb - 1 : (tb == 0 ? (int)0x80000001 : tb + 1));
    *thr_mine = tu;
    __syncwarp();
  }
};

// Check one 32-column chunk against the row threshold.
//   rare path 1: a train item of this row falls in the chunk -> its score is overwritten with -inf
//   fast path  : max over 4 groups of 8 columns (3-input max), one compare against the threshold
//   rare path 2: groups whose max beats the threshold append their survivors to the queue
template <int KMAX, bool SMALLQ, int NSEG, bool SHARE>
__device__ __forceinline__ void epi_chunk(uint32_t (&v)[32], int j0, EpiState<KMAX, NSEG, SHARE>& st, int q_cap,
                                          TrainCursor& tc) {
  // Train items of this row inside the chunk (rare per lane, ~2%): overwrite their score with -inf.
  if (tc.next < j0 + 32) {
    uint32_t excl = 0;
    do {
      if (tc.next >= j0) excl |= 1u << (tc.next - j0);
      tc.advance();
    } while (tc.next < j0 + 32);
    // one 32-column predicated sweep: per-8-column-group sweeps (fewer executed instructions, more branches
    // and code) measured 2.45 vs 2.35 ms on B200
    if (excl) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((excl >> i) & 1u) v[i] = 0xff800000u;   // -inf
    }
  }
  // Columns past the end of the catalogue need no code: TMA fills out-of-bounds operand rows with NaN, their
  // scores are NaN, and max / '>' ignore NaN.  (A per-chunk "j >= M" pre-mask cost 8%: 2.27 vs 2.10 ms.)
  // 4 groups of 8 columns.  Measured on B200 (Amazon-Book shape): 8 groups of 4 -> 2.26 ms, 2 groups of 16 -> 2.14 ms,
  // 4 groups of 8 -> 2.06 ms (more groups = more branches, fewer groups = longer predicated bodies).
  constexpr int NG = 4;
  float gm[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const float a = max3(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]), __uint_as_float(v[8 * g + 2]));
    const float b = max3(__uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
    gm[g] = fmaxf(max3(a, b, __uint_as_float(v[8 * g + 6])), __uint_as_float(v[8 * g + 7]));
  }
  const float m = fmaxf(max3(gm[0], gm[1], gm[2]), gm[3]);
  constexpr int GW = 32 / NG;
  const float th = st.filter();
  if (__any_sync(0xffffffffu, m > th)) {               // warp-uniform
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      // warp-uniform branch (a lane-divergent one measured 2% slower); the per-column predicate guards the lanes
      if (__any_sync(0xffffffffu, gm[g] > th)) {
#pragma unroll
        for (int i = GW * g; i < GW * g + GW; ++i) {
          const float s = __uint_as_float(v[i]);
          if (s > th) st.append(s, j0 + i);
        }
      }
      if (SMALLQ) {                                     // small queues (big user tile in smem): check per group
        __syncwarp();
        if (__any_sync(0xffffffffu, st.fuller_than(q_cap - 8))) st.flush();
      }
    }
    if (!SMALLQ) {
      __syncwarp();
      if (__any_sync(0xffffffffu, st.fuller_than(q_cap - 32))) st.flush();   // warp-convergent
    }
  }
}

template <int KMAX, bool SMALLQ, int NSEG, bool SHARE>
__global__ void __launch_bounds__(TcGeo<NSEG>::THREADS, 1)
k_score_topk_tc(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_i,
                const TcParams p) {
  using G = TcGeo<NSEG>;
  constexpr int EPI = G::EPI;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gbase = smem_dyn + (base - raw);
  const uint32_t sA = base;
  const uint32_t sB = sA + (uint32_t)p.k_blocks * TC_A_BLOCK_BYTES;
  const uint32_t off_stages = (uint32_t)p.k_blocks * TC_A_BLOCK_BYTES;
  const uint32_t off_queue = off_stages + (uint32_t)p.stages * G::STAGE_BYTES;
  // The final lists are staged over memory that is idle by then: the B-operand ring for NSEG 2
  // (KMAX*256*8 <= 64 KB <= 2 stages), the drained candidate queues for NSEG 3 (q_cap >= KMAX).
  unsigned char* stage_area = gbase + (NSEG == 2 ? off_stages : off_queue);
  float* lval_all = reinterpret_cast<float*>(stage_area);
  int32_t* lidx_all = reinterpret_cast<int32_t*>(stage_area + (size_t)KMAX * EPI * 4);
  const uint32_t off_bar = off_queue + (uint32_t)p.q_cap * EPI * 8;
  const uint32_t bar_full = base + off_bar;                     // [stages]
  const uint32_t bar_empty = bar_full + 8 * TC_MAX_STAGES;      // [stages]
  const uint32_t bar_a = bar_empty + 8 * TC_MAX_STAGES;
  const uint32_t bar_tfull = bar_a + 8;                         // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + off_bar + 8 * (2 * TC_MAX_STAGES + 5));
  float* thr_all = reinterpret_cast<float*>(gbase + off_bar + 8 * (2 * TC_MAX_STAGES + 5) + 24);   // [EPI], 16-byte aligned
  float4* quart_all = reinterpret_cast<float4*>(gbase + off_bar + 8 * (2 * TC_MAX_STAGES + 5) + 24 + 4 * EPI);   // [EPI]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u_tile = blockIdx.x, split = blockIdx.y;
  const int n_tiles = (p.M + G::TILE_I - 1) / G::TILE_I;
  const int t_begin = split * p.tiles_per_split;
  const int n_my = max(0, min(n_tiles, t_begin + p.tiles_per_split) - t_begin);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_u)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_i)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_a, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < EPI) {
    thr_all[threadIdx.x] = -CUDART_INF_F;
    if (p.union_bound) quart_all[threadIdx.x] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && n_my > 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(bar_a, (uint32_t)p.k_blocks * TC_A_BLOCK_BYTES);
      for (int kb = 0; kb < p.k_blocks; ++kb)
        tma_load_2d(sA + kb * TC_A_BLOCK_BYTES, &tmap_u, bar_a, kb * TC_KBLK, u_tile * TC_TILE_U);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_my; ++it) {
        const int row0 = (t_begin + it) * G::TILE_I;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait<true>(bar_empty + 8 * stage, phase ^ 1);
          mbar_expect_tx(bar_full + 8 * stage, G::STAGE_BYTES);
          tma_load_2d(sB + stage * G::STAGE_BYTES, &tmap_i, bar_full + 8 * stage, kb * TC_KBLK, row0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_my > 0) {
      // ---------------------------------------------------------------- MMA issuer (one thread)
      mbar_wait<true>(bar_a, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_my; ++it) {
        const int buf = it & 1;
        mbar_wait<true>(bar_tempty + 8 * buf, (uint32_t)((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * TC_TMEM_BUF;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait<true>(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(sA + kb * TC_A_BLOCK_BYTES);
          const uint64_t bdesc = umma_desc_sw128(sB + stage * G::STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < TC_KBLK / 16; ++k)   // advance 16 bf16 = 32 B inside the swizzle atom: +2 in 16-byte units
            tc_mma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), G::IDESC,
                       (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(bar_empty + 8 * stage);          // frees the smem stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * buf);              // accumulator tile complete
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (EPI threads)
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int h = (warp - 4) >> 2;            // which column segment of the tile
    const int row = q * 32 + lane;            // user row inside the tile == TMEM lane
    const int col = h * TC_TILE_U + row;      // this thread's list column
    EpiState<KMAX, NSEG, SHARE> st;
    st.init(p.K);
    st.qbase = st.qhead = reinterpret_cast<float*>(gbase + off_queue) + col;
    st.thr_mine = thr_all + col;               // partners: same row, other segment(s)
    st.thr_other = thr_all + ((h + 1) % NSEG) * TC_TILE_U + row;
    st.thr_other2 = thr_all + ((h + 2) % NSEG) * TC_TILE_U + row;
    st.quart_mine = quart_all + col;
    st.quart_other = quart_all + ((h + 1) % NSEG) * TC_TILE_U + row;
    st.use_union = NSEG == 2 && p.K == KMAX && p.union_bound;
    st.gbound = (SHARE && u_tile * TC_TILE_U + row < p.B) ? p.row_bound + (u_tile * TC_TILE_U + row) : nullptr;
    if (SHARE) st.flush();                     // start from what earlier / concurrent units of this row already know
    const int u = u_tile * TC_TILE_U + row;
    const int64_t uid = (u < p.B) ? (p.users ? p.users[u] : (int64_t)u) : -1;
    TrainCursor tcur;
    tcur.init(p.mask, uid, p.item_offset, t_begin * G::TILE_I);
    for (int it = 0; it < n_my; ++it) {
      const int buf = it & 1;
      mbar_wait<false>(bar_tfull + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_TMEM_BUF + h * G::SEG_COLS);
      const int j_base = (t_begin + it) * G::TILE_I + h * G::SEG_COLS;
      if (NSEG == 2) {
        uint32_t va[32], vb[32];
        LGX_TMEM_LD32(va, taddr);
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {        // rolled: two chunk bodies in the instruction stream, not four
          LGX_TMEM_WAIT(va);
          LGX_TMEM_LD32(vb, taddr + (uint32_t)(c + 1) * 32);
          epi_chunk<KMAX, SMALLQ, NSEG, SHARE>(va, j_base + c * 32, st, p.q_cap, tcur);
          LGX_TMEM_WAIT(vb);
          if (c == 0) {
            LGX_TMEM_LD32(va, taddr + 64);
          } else {
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * buf);  // all four chunks are in registers: TMEM buffer may be overwritten
          }
          epi_chunk<KMAX, SMALLQ, NSEG, SHARE>(vb, j_base + (c + 1) * 32, st, p.q_cap, tcur);
        }
      } else {
        // two chunks per tile, one register buffer: the other two warps of this scheduler cover the load latency
        uint32_t va[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          LGX_TMEM_LD32(va, taddr + (uint32_t)c * 32);
          LGX_TMEM_WAIT(va);
          if (c == 1) {
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * buf);
          }
          epi_chunk<KMAX, SMALLQ, NSEG, SHARE>(va, j_base + c * 32, st, p.q_cap, tcur);
        }
      }
    }
    st.flush();
    if (NSEG != 2) asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");   // every queue drained before it is overwritten
    // stage the register lists, merge the column segments of every row and publish the split's partial list
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      lval_all[k * EPI + col] = st.lv[k];
      lidx_all[k * EPI + col] = st.li[k];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
    if (h == 0 && u < p.B) {
      int pos[NSEG];                              // lists are right-aligned in their KMAX slots
#pragma unroll
      for (int s = 0; s < NSEG; ++s) pos[s] = KMAX - p.K;
      const int64_t o = ((int64_t)split * p.B + u) * p.K;
      for (int k = 0; k < p.K; ++k) {
        float bv = lval_all[pos[0] * EPI + row];
        int32_t bi = lidx_all[pos[0] * EPI + row];
        int bs = 0;
#pragma unroll
        for (int s = 1; s < NSEG; ++s) {
          const float fv = lval_all[pos[s] * EPI + s * TC_TILE_U + row];
          const int32_t fi = lidx_all[pos[s] * EPI + s * TC_TILE_U + row];
          if (better(fv, fi, bv, bi)) { bv = fv; bi = fi; bs = s; }
        }
        p.ws_val[o + k] = bv; p.ws_idx[o + k] = bi;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) pos[s] += (bs == s);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// --------------------------------------------------------------------------------------- host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// [rows, ktot] bf16 row-major -> boxes of [box_rows x 64] with the 128-byte swizzle
int make_operand_map(CUtensorMap* map, const void* ptr, int rows, int ktot, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LGX_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return LGX_ERR_CUDA; }
  return LGX_OK;
}

struct TcConfig { int nseg, k_blocks, stages, q_cap, tile_items, union_bound; size_t smem; bool ok; };

static TcConfig tc_config_nseg(int d, int K, int mode, int nseg) {
  TcConfig c{};
  c.nseg = nseg;
  const int epi = nseg * TC_TILE_U;
  const size_t stage_bytes = nseg == 2 ? TcGeo<2>::STAGE_BYTES : TcGeo<3>::STAGE_BYTES;
  c.tile_items = nseg == 2 ? TcGeo<2>::TILE_I : TcGeo<3>::TILE_I;
  const int ktot = mode == LGX_SCORE_BF16X3 ? 3 * d : d;
  c.ok = (d % TC_KBLK == 0) && K >= 1 && K <= (nseg == 2 ? 32 : 24);
  c.k_blocks = ktot / TC_KBLK;
  // Queue depth vs B-ring depth, first fit: 40-row queues, else 36 rows when that buys the missing ring stage,
  // else 16-row queues (SMALLQ kernels, no union bound: its quartile array would cost them a stage).
  static const int forced_q = [] { const char* e = std::getenv("LGX_SCORE_QCAP"); return e ? std::atoi(e) : 0; }();
  static const int union_env = [] { const char* e = std::getenv("LGX_SCORE_UNION"); return e ? std::atoi(e) : 1; }();
  const int want = c.k_blocks == 1 ? 4 : 3;
  const int cands[3] = {40, 36, 16};
  size_t fixed = 0;
  bool found = false;
  for (int i = 0; i < 3 && !found; ++i) {
    const int qc = cands[i];
    if (forced_q && qc != forced_q) continue;              // LGX_SCORE_QCAP=40|36|16: A/B runs
    if (nseg == 3 && qc < 40) continue;                    // the 12-warp variant stages its lists over full-size queues
    const int uni = (qc >= 36 && nseg == 2 && union_env) ? 1 : 0;
    const size_t fx = 1024 + (size_t)c.k_blocks * TC_A_BLOCK_BYTES + (size_t)qc * epi * 8 +
                      8 * (2 * TC_MAX_STAGES + 5) + 24 + (size_t)(4 + 16 * uni) * epi;
    const int need = (qc == 16 || forced_q || nseg == 3) ? 2 : want;
    if (fx + (size_t)need * stage_bytes > TC_SMEM_LIMIT) continue;
    c.q_cap = qc; c.union_bound = uni; fixed = fx; found = true;
    c.stages = (int)std::min<size_t>(TC_MAX_STAGES, (TC_SMEM_LIMIT - fx) / stage_bytes);
  }
  if (!c.ok || !found) { c.ok = false; return c; }
  c.smem = fixed + (size_t)c.stages * stage_bytes;
  return c;
}

// Default: 8 epilogue warps over 256-column tiles.  LGX_SCORE_NSEG=3 selects the 12-warp / 192-column variant
// where it fits (k <= 24, full-size queues + >= 2 stages: d = 64, bf16 d = 128).  Measured on B200 at the
// Amazon-Book shape it is 2% SLOWER (2.42 vs 2.37 ms): a third list per row adds inserts and the single
// TMEM register buffer exposes load latency, which cancels the extra latency hiding -- kept for A/B runs.
static TcConfig tc_config(int d, int K, int mode) {
  static const int forced = [] { const char* e = std::getenv("LGX_SCORE_NSEG"); return e ? std::atoi(e) : 0; }();
  if (forced == 3) {
    const TcConfig c3 = tc_config_nseg(d, K, mode, 3);
    if (c3.ok) return c3;
  }
  return tc_config_nseg(d, K, mode, 2);
}

// Splits per user tile: wave-aware (lgx_score_plan.cuh).  LGX_SCORE_SPLITS=R forces R (A/B runs),
// LGX_SCORE_UNIT_OVERHEAD sets the chooser's per-unit overhead in item tiles of 256 columns.
static ScorePlan tc_plan(int B, int M, const TcConfig& cfg, int sms) {
  static const int forced = [] { const char* e = std::getenv("LGX_SCORE_SPLITS"); return e ? std::atoi(e) : 0; }();
  static const double c0 = [] { const char* e = std::getenv("LGX_SCORE_UNIT_OVERHEAD"); return e ? std::atof(e) : 14.0; }();
  ScorePlan p = plan_score_waves(B, M, TC_TILE_U, cfg.tile_items, sms, c0 * 256.0 / cfg.tile_items);
  if (forced > 0) {
    const int r = std::max(1, std::min(forced, std::min(p.n_item_tiles, kMaxSplits)));
    p.tiles_per_split = (p.n_item_tiles + r - 1) / r;
    p.n_splits = (p.n_item_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  }
  return p;
}

template <int KMAX, bool SMALLQ, int NSEG, bool SHARE>
static int launch_tc2(dim3 grid, const TcConfig& cfg, const CUtensorMap& tm_u, const CUtensorMap& tm_i,
                     const TcParams& p, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};       // the opt-in is a per-device function attribute
  const int dev = current_device();
  if (dev >= kMaxDevices || !configured[dev]) {
    LGX_CHECK_CUDA(cudaFuncSetAttribute(k_score_topk_tc<KMAX, SMALLQ, NSEG, SHARE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    if (dev < kMaxDevices) configured[dev] = true;
  }
  k_score_topk_tc<KMAX, SMALLQ, NSEG, SHARE><<<grid, TcGeo<NSEG>::THREADS, cfg.smem, st>>>(tm_u, tm_i, p);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

// SHARE (rows' bounds exchanged through TcParams::row_bound) is compiled in only when a row is split over several
// CTAs: carried unused it cost the single-split Amazon-Book pass 4 % (2.06 -> 2.15 ms).
template <int KMAX, bool SMALLQ, int NSEG>
static int launch_tc(dim3 grid, const TcConfig& cfg, const CUtensorMap& tm_u, const CUtensorMap& tm_i,
                     const TcParams& p, cudaStream_t st) {
  return p.row_bound ? launch_tc2<KMAX, SMALLQ, NSEG, true>(grid, cfg, tm_u, tm_i, p, st)
                     : launch_tc2<KMAX, SMALLQ, NSEG, false>(grid, cfg, tm_u, tm_i, p, st);
}

static int score_topk_tc(const lgx_graph* g, const void* U_op, const int64_t* users, int B, const void* I_op, int M,
                         int d, int K, int mode, int64_t item_offset, int64_t* out_idx, float* out_val,
                         void* workspace, cudaStream_t st) {
  // the group-queue kernel where its shared-memory budget fits (everything but K = 3d = 384)
  if (use_gq(d, K, mode)) return score_topk_gq(g, U_op, users, B, I_op, M, d, K, mode, item_offset, out_idx, out_val, workspace, st);
  const TcConfig cfg = tc_config(d, K, mode);
  if (!cfg.ok) {
    set_error("tcgen05 scoring needs d % 64 == 0, k <= 32 and the user tile + 2 item stages to fit in 227 KB "
              "(d<=256 for bf16, d<=128 for bf16x3); use LGX_SCORE_FP32 otherwise");
    return LGX_ERR_INVALID;
  }
  const int ktot = cfg.k_blocks * TC_KBLK;
  CUtensorMap tm_u, tm_i;
  int rc = make_operand_map(&tm_u, U_op, B, ktot, TC_TILE_U);
  if (rc != LGX_OK) return rc;
  rc = make_operand_map(&tm_i, I_op, M, ktot, cfg.tile_items);
  if (rc != LGX_OK) return rc;
  const ScorePlan plan = tc_plan(B, M, cfg, sm_count());
  TcParams p;
  p.B = B; p.M = M; p.K = K; p.k_blocks = cfg.k_blocks; p.stages = cfg.stages; p.q_cap = cfg.q_cap;
  p.n_splits = plan.n_splits; p.tiles_per_split = plan.tiles_per_split; p.item_offset = item_offset;
  p.mask = make_mask(g); p.users = users;
  p.union_bound = cfg.union_bound;
  p.ws_val = reinterpret_cast<float*>(workspace);
  p.ws_idx = reinterpret_cast<int32_t*>(p.ws_val + (size_t)plan.n_splits * B * K);
  p.row_bound = nullptr;
  if (plan.n_splits > 1) {                      // rows are worked on by several CTAs: share their bounds
    p.row_bound = reinterpret_cast<unsigned*>(p.ws_idx + (size_t)plan.n_splits * B * K);
    LGX_CHECK_CUDA(cudaMemsetAsync(p.row_bound, 0, (size_t)B * sizeof(unsigned), st));
  }
  dim3 grid(plan.n_user_tiles, plan.n_splits);
  const bool smallq = cfg.q_cap < 36;
  if (cfg.nseg == 3) {
    if (K <= 20) rc = launch_tc<20, false, 3>(grid, cfg, tm_u, tm_i, p, st);   // the reference's default topks=[20]
    else rc = launch_tc<24, false, 3>(grid, cfg, tm_u, tm_i, p, st);
  } else if (K <= 20 && !smallq) rc = launch_tc<20, false, 2>(grid, cfg, tm_u, tm_i, p, st);
  else if (K <= 24 && !smallq) rc = launch_tc<24, false, 2>(grid, cfg, tm_u, tm_i, p, st);
  else if (K <= 24) rc = launch_tc<24, true, 2>(grid, cfg, tm_u, tm_i, p, st);
  else if (!smallq) rc = launch_tc<32, false, 2>(grid, cfg, tm_u, tm_i, p, st);
  else rc = launch_tc<32, true, 2>(grid, cfg, tm_u, tm_i, p, st);
  if (rc != LGX_OK) return rc;
  return launch_merge_i32(p.ws_val, p.ws_idx, plan.n_splits, B, K, item_offset, M, p.mask, users, out_idx, out_val, st);
}

}  // namespace lgx

using namespace lgx;

extern "C" {

size_t lgx_score_topk_workspace_bytes(int32_t B, int32_t M, int32_t d, int32_t k, int32_t mode) {
  if (B <= 0 || M <= 0 || k <= 0) return 0;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) sms = sm_count(); else cudaGetLastError();
  if (mode == LGX_SCORE_FP32) return (size_t)plan_score(B, M, 64, 64, sms, 4).n_splits * B * k * 8 + 256;
  if (use_gq(d, k, mode)) return (size_t)gq_plan(B, M, sms).n_splits * B * k * 8 + (size_t)B * sizeof(unsigned) + 256;
  TcConfig cfg = tc_config(d, k, mode);
  if (!cfg.ok) cfg.tile_items = TcGeo<2>::TILE_I;
  const ScorePlan plan = tc_plan(B, M, cfg, sms);
  return (size_t)plan.n_splits * B * k * 8 + (size_t)B * sizeof(unsigned) + 256;
}

int lgx_score_plan(int32_t B, int32_t M, int32_t d, int32_t k, int32_t mode, int32_t sms, int32_t* plan4) {
  LGX_REQUIRE(plan4 != nullptr, "NULL argument");
  LGX_REQUIRE(B > 0 && M > 0 && d > 0 && k > 0, "B, M, d, k must be positive");
  LGX_REQUIRE(mode == LGX_SCORE_FP32 || mode == LGX_SCORE_BF16 || mode == LGX_SCORE_BF16X3, "unknown mode");
  if (sms <= 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) sms = sm_count(); else { cudaGetLastError(); sms = 148; }
  }
  ScorePlan plan;
  if (mode == LGX_SCORE_FP32) {
    plan = plan_score(B, M, 64, 64, sms, 4);
  } else if (use_gq(d, k, mode)) {
    plan = gq_plan(B, M, sms);
  } else {
    TcConfig cfg = tc_config(d, k, mode);
    if (!cfg.ok) cfg.tile_items = TcGeo<2>::TILE_I;
    plan = tc_plan(B, M, cfg, sms);
  }
  plan4[0] = plan.n_user_tiles; plan4[1] = plan.n_item_tiles; plan4[2] = plan.n_splits; plan4[3] = plan.tiles_per_split;
  return LGX_OK;
}

int lgx_score_topk(const lgx_graph* g, const void* U_op, const int64_t* users, int32_t B, const void* I_op, int32_t M,
                   int32_t d, int32_t k, int32_t mode, int64_t item_offset, int64_t* out_idx, float* out_val,
                   void* workspace, size_t workspace_bytes, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(U_op && I_op && out_idx && out_val && workspace, "NULL argument");
  LGX_REQUIRE(B > 0 && M > 0 && d > 0, "B, M, d must be positive");
  LGX_REQUIRE(k > 0 && k <= M && k <= 255, "k must be in [1, min(M, 255)]");
  LGX_REQUIRE(g == nullptr || users != nullptr || B <= g->n_users, "mask needs user ids");
  LGX_REQUIRE(mode == LGX_SCORE_FP32 || mode == LGX_SCORE_BF16 || mode == LGX_SCORE_BF16X3, "unknown mode");
  if (workspace_bytes < lgx_score_topk_workspace_bytes(B, M, d, k, mode)) {
    set_error("workspace too small for lgx_score_topk");
    return LGX_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == LGX_SCORE_FP32)
    return score_topk_fp32(g, reinterpret_cast<const float*>(U_op), users, B, reinterpret_cast<const float*>(I_op), M,
                           d, k, item_offset, out_idx, out_val, workspace, st);
  LGX_REQUIRE((reinterpret_cast<uintptr_t>(U_op) & 15) == 0 && (reinterpret_cast<uintptr_t>(I_op) & 15) == 0,
              "packed operands must be 16-byte aligned");
  return score_topk_tc(g, U_op, users, B, I_op, M, d, k, mode, item_offset, out_idx, out_val, workspace, st);
}

}  // extern "C"
