// Fused full-catalogue scoring + train mask + top-K on the 5th-gen tensor cores (sm_100a), "group queue" kernel.
//
// Replaces getUsersRating (PT/model.py:179-184), the train-item mask (PT/Procedure.py:129-134) and torch.topk
// (:135).  The [B, M] score matrix lives only in TMEM.
//
// What changed against the first tcgen05 kernel (lgx_score_tc.cu, kept for A/B runs): that kernel's epilogue spent
// 5 issue slots per score, ~85 % of them outside the max filter: per-column candidate appends (a warp holds 32
// independent rows, so "rare per row" is "always" per warp), the train-mask sweep and the mask cursor's dependent
// global loads.  Here
//   * the TRAIN MASK IS APPLIED BY THE TENSOR CORE: two helper warps take the tile's pre-bucketed train entries
//     (k_mask_buckets) and give every row that has train items in the tile one K-slot of a small mask operand pair in
//     shared memory: A_mask[row, slot] = -2^100, B_mask[col, slot] = 1 for the row's train columns (or one slot per
//     COLUMN when the tile holds a popular item: more than 32 dirty rows, fewer dirty columns).  One or two extra
//     K = 16 MMA steps add -2^100 to exactly the train (row, col) scores, so the epilogue has no mask code at all;
//   * the epilogue keeps the top-K GROUPS of 8 columns by group maximum (the 3-input max tree it needs anyway):
//     a candidate is one predicated (max, group id) append per group instead of an 8-column scan.  The K-th best
//     group maximum is a valid lower bound on the row's K-th best score (K distinct items score at least that),
//     so every item of the final top-K lies in one of the K kept groups.  Queued candidates are inserted into the
//     register-resident sorted list at most two per lane per check (every two 32-column chunks): one lane's long
//     queue must not make its warp -- and through the accumulator hand-back the whole pipeline -- wait;
//   * a second small kernel (k_rescore_topk, one warp per row) recomputes the 8K scores of the kept groups from the
//     same bf16 operands in fp32, re-applies the mask exactly and emits the sorted top-K (ties by item id).
//
//   warp 0      TMA producer (user tile resident, item K-blocks [256 x 64] bf16 through an mbarrier ring)
//   warp 1      MMA issuer: tcgen05.mma.kind::f16 M128 N256 K16 into double-buffered fp32 TMEM accumulators,
//               the tile's mask steps, then the real K steps; the next tile's barriers are checked behind three queued
//               MMAs of the current one
//   warp 2      TMEM alloc / dealloc; mask builder of the odd tiles (mask buffer 1)
//   warp 3      mask builder of the even tiles (mask buffer 0): scatters the tile's pre-bucketed train entries into
//               its 32-slot buffer
//   warps 4-11  epilogue: one thread = one user row x one 128-column half of the tile
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "lgx_common.cuh"
#include "lgx_score_plan.cuh"
#include "lgx_topk.cuh"
#include <mutex>
#include "lgx_tc_ptx.cuh"

namespace lgx {

TrainMask make_mask(const lgx_graph* g);
int make_operand_map(CUtensorMap* map, const void* ptr, int rows, int ktot, int box_rows);

constexpr int GQ_TILE_U = 128;                  // UMMA M: one TMEM lane per user
constexpr int GQ_TILE_I = 256;                  // UMMA N
constexpr int GQ_KBLK = 64;                     // bf16 per K block = one 128-byte swizzle row
constexpr int GQ_A_BLOCK = GQ_TILE_U * GQ_KBLK * 2;   // 16 KB
constexpr int GQ_STAGE = GQ_TILE_I * GQ_KBLK * 2;     // 32 KB
constexpr int GQ_EPI = 256;                     // epilogue threads
constexpr int GQ_THREADS = 128 + GQ_EPI;
constexpr int GQ_MAX_STAGES = 6;
constexpr int GQ_TMEM_BUF = 256;
#ifndef LGX_GQ_GROUP
#define LGX_GQ_GROUP 8
#endif
#ifndef LGX_GQ_EARLY_CHUNK
#define LGX_GQ_EARLY_CHUNK 0
#endif
#ifndef LGX_GQ_READY_AFTER
#define LGX_GQ_READY_AFTER 3   // the next tile's readiness checks run behind this many queued MMAs of the last K block
#endif
#ifndef LGX_GQ_TH_REFRESH
#define LGX_GQ_TH_REFRESH 0
#endif
#ifndef LGX_GQ_EVERY
#define LGX_GQ_EVERY 2
#endif
#ifndef LGX_GQ_SMALLQ_BELOW
#define LGX_GQ_SMALLQ_BELOW 32      // 24: one check per tile for 24-row queues measured 15 % slower (fewer, longer drains)
#endif
#ifndef LGX_GQ_DRAIN
#define LGX_GQ_DRAIN 2         // queue entries inserted per lane per check once a lane holds LGX_GQ_LOW (0: only full flushes)
#endif
#ifndef LGX_GQ_LOW
#define LGX_GQ_LOW 4
#endif
// columns per candidate group: 8.  (4 halves the rescoring work -- 63 us less under ncu -- for twice the predicated
// appends in the epilogue: measured 22 us slower per call before the bounded drains and 31 % slower with them; the
// 4-column code paths are kept for A/B builds with -DLGX_GQ_GROUP=4.)
constexpr int GQ_GROUP = LGX_GQ_GROUP;
static_assert(GQ_GROUP == 8 || GQ_GROUP == 4, "groups of 4 or 8 columns");
constexpr int GQ_SMEM_LIMIT = 232448;
constexpr uint32_t GQ_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GQ_TILE_I >> 3) << 17) |
                              ((uint32_t)(GQ_TILE_U >> 4) << 24);
constexpr unsigned short GQ_BF16_ONE = 0x3F80;       // 1.0
constexpr unsigned short GQ_BF16_NEG_BIG = 0xF180;   // -2^100: finite (0 * it stays 0), below any real score

struct GqParams {
  int B, M, K;
  int k_blocks, stages;
  int q_cap;
  int union_bound;
  int n_splits, tiles_per_split;
  int rotate;             // 1: every CTA starts its item-tile walk at its own offset
  int cl;                 // CTAs per cluster sharing every B tile by TMA multicast (1 = no cluster)
  int has_mask;
  int dbg;                // LGX_GQ_DEBUG (experiments): 1 = TMEM loads only, 2 = no TMEM loads either, 4 = no mask,
                          // 8 = masks built but no mask MMA issued, 32 = no pre-bucketing (in-kernel list walk)
  int64_t item_offset;
  TrainMask mask;
  const int64_t* users;
  float* ws_val;          // [n_splits, B, K] group maxima, best first
  int32_t* ws_idx;        // [n_splits, B, K] group ids (local item id / 8)
  unsigned* row_bound;    // [B] or NULL (n_splits == 1)
  // pre-bucketed train entries (k_mask_buckets); mk_region == NULL: every user tile walks its lists in the kernel
  const int64_t* mk_region;
  const int32_t* mk_ptr;
  const uint16_t* mk_entries;
};

__device__ __forceinline__ unsigned gq_ord_encode(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float gq_ord_decode(unsigned u) {
  return u == 0u ? -CUDART_INF_F : __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// byte offset of element (row, k) in a K-major [rows x 64] bf16 tile with the 128-byte swizzle
__device__ __forceinline__ uint32_t gq_swz(int row, int k) {
  return (uint32_t)row * 128u + (uint32_t)((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1));
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, unsigned short v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// ------------------------------------------------------------------------------------ mask buckets
// The in-kernel mask builder must publish one mask per ~700 cycles.  Walking 128 sorted train lists tile by tile
// (cursor, dirty ballots, slot ranks, undo records) took ~460 instructions per tile -- ncu showed the builder warps
// 94 % busy and the epilogue starved -- so that irregular part runs ahead of time, massively parallel, in a small
// pre-kernel: one CTA per user tile buckets the tile's train entries by item tile (count, scan, fill) into
//     ptr[u_tile][n_tiles + 1], entries[region[u_tile] + ptr ..] = (row << 8 | column), 16 bits each, any order.
// The builder then only scatters ~20 pre-bucketed entries per tile.  The entry area is a bump allocator sized for
// the usual case; a user tile that does not fit gets region = -1 and the main kernel walks its lists itself.
struct GqBucketParams {
  TrainMask mask;
  const int64_t* users;
  int B, M, n_tiles;
  int64_t item_offset;
  unsigned long long* cursor;     // bump allocator state (zeroed by the caller)
  unsigned long long cap;         // entries available
  int64_t* region;                // [n_user_tiles]
  int32_t* ptr;                   // [n_user_tiles][n_tiles + 1]
  uint16_t* entries;
};

constexpr int MB_THREADS = 1024;      // latency-bound (one dependent global load per entry): many threads per user tile
__global__ void __launch_bounds__(MB_THREADS)
k_mask_buckets(const GqBucketParams bp) {
  extern __shared__ int s_cnt[];                 // [n_tiles + 1]
  __shared__ int64_t s_lo[GQ_TILE_U];            // first in-range CSR entry of every row
  __shared__ int s_pre[GQ_TILE_U + 1];           // exclusive prefix of the rows' in-range lengths
  __shared__ int s_warp[MB_THREADS / 32];
  __shared__ long long s_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ut = blockIdx.x;
  const int n1 = bp.n_tiles + 1;
  const int32_t bias = (int32_t)(bp.mask.n_users + bp.item_offset);
  const bool whole = bp.mask.m_items_hint > 0 && bp.item_offset == 0 && bp.M >= bp.mask.m_items_hint;   // the whole catalogue: no searches
  for (int i = tid; i < n1; i += MB_THREADS) s_cnt[i] = 0;
  if (tid < GQ_TILE_U) {
    const int u = ut * GQ_TILE_U + tid;
    int64_t lo = 0, hi = 0;
    if (u < bp.B) {
      const int64_t uid = bp.users ? bp.users[u] : (int64_t)u;
      const int64_t r0 = bp.mask.indptr[uid], r1 = bp.mask.indptr[uid + 1];
      lo = r0; hi = r1;
      if (!whole) {
        int64_t a = r0, b = r1;                     // first entry >= bias
        while (a < b) { const int64_t m = (a + b) >> 1; if (__ldg(bp.mask.indices + m) < bias) a = m + 1; else b = m; }
        lo = a;
        b = r1;                                     // first entry >= bias + M
        const int64_t key = (int64_t)bias + bp.M;
        while (a < b) { const int64_t m = (a + b) >> 1; if ((int64_t)__ldg(bp.mask.indices + m) < key) a = m + 1; else b = m; }
        hi = a;
      }
    }
    s_lo[tid] = lo;
    // inclusive scan of the lengths over the 128 rows (4 warps)
    int incl = (int)(hi - lo);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    s_pre[tid + 1] = incl;                          // warp-local for now
  }
  __syncthreads();
  if (tid < GQ_TILE_U) {
    int add = 0;
    for (int w2 = 0; w2 < warp; ++w2) add += s_warp[w2];
    s_pre[tid + 1] += add;
  }
  if (tid == 0) s_pre[0] = 0;
  __syncthreads();
  const int total = s_pre[GQ_TILE_U];
  // the tile's entries as one flat range: thread-strided, every load independent of the others
  auto locate = [&](int idx, int& row) -> int64_t {
    int a = 0, b = GQ_TILE_U;                       // last row with s_pre[row] <= idx
    while (b - a > 1) { const int m = (a + b) >> 1; if (s_pre[m] <= idx) a = m; else b = m; }
    row = a;
    return s_lo[a] + (idx - s_pre[a]);
  };
  for (int idx = tid; idx < total; idx += MB_THREADS) {
    int row;
    const int64_t e = locate(idx, row);
    atomicAdd(&s_cnt[(__ldg(bp.mask.indices + e) - bias) >> 8], 1);
  }
  __syncthreads();
  // exclusive scan of s_cnt[0 .. n_tiles): thread t owns a contiguous chunk
  const int chunk = (n1 + MB_THREADS - 1) / MB_THREADS;
  const int c0 = min(tid * chunk, n1), c1 = min(c0 + chunk, n1);
  int sum = 0;
  for (int i = c0; i < c1; ++i) sum += s_cnt[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __syncthreads();                                  // s_warp is reused
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int base = incl - sum;
  for (int w2 = 0; w2 < warp; ++w2) base += s_warp[w2];
  for (int i = c0; i < c1; ++i) { const int v = s_cnt[i]; s_cnt[i] = base; base += v; }
  __syncthreads();
  if (tid == 0) {
    const unsigned long long tot = (unsigned long long)s_cnt[bp.n_tiles];
    const unsigned long long at = atomicAdd(bp.cursor, tot);
    s_base = (at + tot <= bp.cap) ? (long long)at : -1;
    bp.region[ut] = s_base;
  }
  __syncthreads();
  if (s_base < 0) return;
  for (int i = tid; i < n1; i += MB_THREADS) bp.ptr[(int64_t)ut * n1 + i] = s_cnt[i];
  __syncthreads();
  uint16_t* out = bp.entries + s_base;
  for (int idx = tid; idx < total; idx += MB_THREADS) {
    int row;
    const int64_t e = locate(idx, row);
    const int c = __ldg(bp.mask.indices + e) - bias;
    out[atomicAdd(&s_cnt[c >> 8], 1)] = (uint16_t)((row << 8) | (c & 255));
  }
}

// rank of `key` among the set bits of an N x 32-bit map (bits below it)
template <int N>
__device__ __forceinline__ int gq_bits_rank(const unsigned (&m)[N], int key) {
  const int kw = key >> 5;
  const unsigned low = (1u << (key & 31)) - 1u;
  int r = 0;
#pragma unroll
  for (int w = 0; w < N; ++w) r += w < kw ? __popc(m[w]) : (w == kw ? __popc(m[w] & low) : 0);
  return r;
}

// Fallback cursor (a user tile whose entries did not fit the bucket area): one dependent load per train item.
struct GqCursor {
  const int32_t* idx;
  int64_t cur, end;
  int32_t bias, next;
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

// Per-thread epilogue state: the top-K groups of this thread's half of the row, sorted best-first in registers
// (right-aligned in KMAX slots), a shared-memory candidate queue that only the warp-convergent flush drains, and
// the row-threshold exchange with the thread that owns the other half (see lgx_score_tc.cu for the derivation).
template <int KMAX, bool SHARE>
struct GqEpi {
  float lv[KMAX];
  int32_t li[KMAX];
  uint32_t qb;            // shared-memory address of this thread's queue slot 0; slot n at qb + n * ROWB
  int qn, qh;             // queue slots [qh, qn) hold candidates not yet inserted (FIFO; both reset when it empties)
  unsigned pubs;          // bounded drains since the start (SHARE: throttles the global bound exchange)
  float gseen;            // SHARE: best row bound read from the other splits so far
  static constexpr uint32_t ROWB = GQ_EPI * 8;
  float* thr_mine; const float* thr_other;
  float4* quart_mine; const float4* quart_other;
  float tu;
  bool use_union;
  unsigned* gbound;
  static constexpr int ROW = GQ_EPI * 2;
  __device__ __forceinline__ float filter() const {
    return max3(lv[KMAX - 1], tu, *reinterpret_cast<const volatile float*>(thr_other));
  }
  __device__ __forceinline__ void init(int K) {
#pragma unroll
    for (int p = 0; p < KMAX; ++p) {
      lv[p] = p < KMAX - K ? CUDART_INF_F : -CUDART_INF_F;
      li[p] = INT32_MAX;
    }
    tu = -CUDART_INF_F;
    pubs = 0u;
    gseen = -CUDART_INF_F;
  }
  // One 8-byte store per candidate at an address derived from the COUNT: the store's address register is never
  // overwritten (a pointer bumped after every store made each predicated bump wait for the store to read it).
  __device__ __forceinline__ void append(float s, int32_t j) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(qb + (uint32_t)qn * ROWB), "f"(s), "r"(j) : "memory");
    ++qn;
  }
  __device__ __forceinline__ bool fuller_than(int rows) const { return qn > rows; }
  __device__ __forceinline__ bool pending_at_least(int n) const { return qn - qh >= n; }
  // groups reach a thread in ascending id order, so an equal maximum loses the tie: strict '>'
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
  __device__ __forceinline__ void pop_insert() {
    float x;
    int32_t xi;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(x), "=r"(xi) : "r"(qb + (uint32_t)qh * ROWB) : "memory");
    insert(x, xi);
    ++qh;
  }
  // everything queued (warp-convergent call; lanes run their own counts)
  __device__ __forceinline__ void flush() {
    while (qh < qn) pop_insert();
    qh = qn = 0;
    publish(true);
  }
  // at most `iters` candidates per lane: one lane's long queue no longer makes its warp (and through the accumulator
  // hand-back, the whole pipeline) wait for a 16-deep insertion burst
  __device__ __forceinline__ void drain(int iters) {
    for (int r = 0; r < iters; ++r) {
      if (qh < qn) pop_insert();
      if (!__any_sync(0xffffffffu, qh < qn)) break;
    }
    if (qh == qn) qh = qn = 0;
    publish(iters > LGX_GQ_DRAIN);
  }
  // `global`: also exchange the row bound with the other item splits of this row (SHARE).  That is a global load and
  // an atomic, ~1 000 cycles of latency in the middle of a drain: done on full flushes and on every 8th bounded drain
  // only (with it on every drain a split unit cost as much as ~80 item tiles of scoring).
  __device__ __forceinline__ void publish(bool global) {
    const bool glob = SHARE && gbound && (global || (++pubs & 7u) == 0u);
    const unsigned gb = glob ? *reinterpret_cast<const volatile unsigned*>(gbound) : 0u;
    float t = lv[KMAX - 1];
    if (use_union) {
      const float a1 = lv[KMAX / 4 - 1], a2 = lv[KMAX / 2 - 1], a3 = lv[3 * KMAX / 4 - 1];
      const volatile float4* o = quart_other;
      const float b1 = o->x, b2 = o->y, b3 = o->z, b4 = o->w;
      *quart_mine = make_float4(a1, a2, a3, t);
      t = fmaxf(max3(fminf(a1, b3), fminf(a2, b2), fminf(a3, b1)), fmaxf(t, b4));
    }
    if (glob) {
      const unsigned mine = (t != t) ? 0u : gq_ord_encode(t);
      if (mine > gb) atomicMax(gbound, mine);
      gseen = fmaxf(gseen, gq_ord_decode(gb));
    }
    if (SHARE) t = fmaxf(t, gseen);          // the last bound seen from the other splits stays valid
    const int tb = __float_as_int(t);                  // publish prev_float(bound); -inf stays -inf
    tu = (t == -CUDART_INF_F || t != t) ? -CUDART_INF_F
                                       : __int_as_float(tb > 0 ? tb - 1 : (tb == 0 ? (int)0x80000001 : tb + 1));
    *thr_mine = tu;
    __syncwarp();
  }
};

// 32 columns = 4 groups: 3-input max tree per group, one predicated append per group.  NaN scores (operand rows
// past the end, filled by TMA) are ignored by max and fail '>'.
template <int KMAX, bool SHARE>
__device__ __forceinline__ void gq_chunk(const uint32_t (&v)[32], int gid0, GqEpi<KMAX, SHARE>& st, float th) {
  if (GQ_GROUP == 8) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float a = max3(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1]), __uint_as_float(v[8 * g + 2]));
      const float b = max3(__uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
      const float m = max3(a, b, fmaxf(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
      if (m > th) st.append(m, gid0 + g);
    }
  } else {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float m = fmaxf(max3(__uint_as_float(v[4 * g + 0]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2])),
                            __uint_as_float(v[4 * g + 3]));
      if (m > th) st.append(m, gid0 + g);
    }
  }
}

#ifdef LGX_GQ_PROF
// Diagnostic build only (-DLGX_GQ_PROF, never shipped): cycle counters summed over all CTAs and tiles.
//  0 mma: wait3 cycles   1 mma: mask-part cycles   2 mma: real-issue cycles   3 mma: tempty not ready at first poll
//  4 mma: mfull not ready   5 mma: full not ready   6 tiles
//  8 epi(warp 0): wait tfull   9 epi: tfull -> tempty arrive   10 epi: arrive -> tile done
//  12 builder: reclaim wait   13 builder: rest   16 tma: wait empty
__device__ unsigned long long gq_prof[32];
// timeline of one CTA (blockIdx.x == 7), tiles [GQ_TR0, GQ_TR0 + 64): [role 0..11][tile][event 0..5] = clock64
constexpr int GQ_TR0 = 120;
__device__ long long gq_trace[12 * 64 * 6];
// per CTA (blockIdx.x < 1024): smid, globaltimer at kernel entry / MMA loop start / MMA loop end / last epilogue warp done / exit
__device__ long long gq_cta[1024 * 8];
__device__ __forceinline__ long long gq_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PROF_T(x) const long long x = clock64()
#define PROF_ADD(i, v) (prof_acc[i] += (unsigned long long)(v))
#define TRACE(role, it, ev, t)                                                                     \
  do {                                                                                             \
    if (blockIdx.x == 7 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (it) >= GQ_TR0 && (it) < GQ_TR0 + 64) \
      gq_trace[((role) * 64 + ((it) - GQ_TR0)) * 6 + (ev)] = (t);                                   \
  } while (0)
#else
#define PROF_T(x)
#define PROF_ADD(i, v)
#define TRACE(role, it, ev, t)
#endif

template <int KMAX, bool SMALLQ, bool SHARE>
__global__ void __launch_bounds__(GQ_THREADS, 1)
k_score_topk_gq(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_i,
                const GqParams p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - raw);
  const uint32_t sA = base;
  const uint32_t off_am = (uint32_t)p.k_blocks * GQ_A_BLOCK;          // A_mask [128 x 64] bf16
  const uint32_t off_bm = off_am + GQ_A_BLOCK;                         // B_mask [256 x 64] bf16
  const uint32_t off_stages = off_bm + GQ_STAGE;
  const uint32_t off_queue = off_stages + (uint32_t)p.stages * GQ_STAGE;
  const uint32_t sAm = base + off_am, sBm = base + off_bm, sB = base + off_stages;
  // final lists are staged over the idle B ring (KMAX * 256 * 8 <= 64 KB <= 2 stages)
  float* lval_all = reinterpret_cast<float*>(gbase + off_stages);
  int32_t* lidx_all = reinterpret_cast<int32_t*>(gbase + off_stages + (size_t)KMAX * GQ_EPI * 4);
  const uint32_t off_bar = off_queue + (uint32_t)p.q_cap * GQ_EPI * 8;
  const uint32_t bar_full = base + off_bar;                       // [stages]
  const uint32_t bar_empty = bar_full + 8 * GQ_MAX_STAGES;        // [stages]
  const uint32_t bar_a = bar_empty + 8 * GQ_MAX_STAGES;
  const uint32_t bar_tfull = bar_a + 8;                           // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                     // [2]
  // mask buffers: the 64-slot mask operand is two 32-slot buffers (K steps {0,1} and {2,3}); tile `it` uses buffer it & 1
  const uint32_t bar_mfull = bar_tempty + 16;                     // [b] the builder published buffer b
  const uint32_t bar_mfree = bar_mfull + 16;                      // [b] the MMAs that read buffer b retired
  constexpr uint32_t kBarBytes = 8 * (2 * GQ_MAX_STAGES + 9);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + off_bar + kBarBytes);
  volatile int* mcnt = reinterpret_cast<volatile int*>(gbase + off_bar + kBarBytes + 8);    // [b] slots used | more << 8
  constexpr uint32_t kMiscBytes = 24;                                                      // tmem slot + mcnt
  float* thr_all = reinterpret_cast<float*>(gbase + off_bar + kBarBytes + kMiscBytes);     // [EPI]
  float4* quart_all = reinterpret_cast<float4*>(gbase + off_bar + kBarBytes + kMiscBytes + 4 * GQ_EPI);   // [EPI], 16-byte aligned

  // Roles: warps 0-7 epilogue, warp 8 TMA, warp 9 MMA, warp 10 TMEM alloc + mask builder, warp 11 mask builder.  The single-lane
  // helper roles get the HIGHEST warp ids: the issue arbiter prefers higher warp ids, and the helpers (one per SM
  // sub-partition, each sharing it with two epilogue warps) are the serial resources of the pipeline.
  const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = hw_warp < 8 ? hw_warp + 4 : hw_warp - 8;      // role index: 0 TMA, 1 MMA, 2 alloc, 3 builder, 4-11 epilogue
  const int u_tile = blockIdx.x, split = blockIdx.y;
  const int n_tiles = (p.M + GQ_TILE_I - 1) / GQ_TILE_I;
  const int t_begin = split * p.tiles_per_split;
  const int n_my = max(0, min(n_tiles, t_begin + p.tiles_per_split) - t_begin);
  // Every CTA walks its item tiles from a different starting point (and wraps): the CTAs of a wave start together
  // and otherwise ask the L2 for the same B tile at the same moment, tile after tile.  The list-walking mask
  // fallback needs ascending tiles, so a user tile without pre-bucketed entries keeps the plain order.
  int rot = 0;
  // (the CTAs of a multicast cluster share every B tile and therefore one order: no rotation there)
  if (p.rotate && !(p.cl > 1 && !(p.dbg & 128)) && n_my > 1 && !(p.has_mask && (p.mk_region == nullptr || p.mk_region[min(u_tile, (p.B - 1) / GQ_TILE_U)] < 0)))
    rot = (int)(((long long)u_tile * 40503LL) % n_my);
  auto tile_of = [&](int it) { const int t = it + rot; return t_begin + (t >= n_my ? t - n_my : t); };
  // Cluster mode (p.cl = 2 or 4 CTAs along x: neighbouring user tiles, the same item tiles): CTA r loads rows
  // [r, r + 1) * 256 / cl of every B tile and TMA multicasts them into all cl CTAs, so a tile leaves the L2 once per
  // cluster instead of once per CTA -- the B stream out of the L2 (32 KB per 512 tensor cycles per SM) is what the
  // pipeline waits for at d = 64.  A ring stage is free again when ALL cl consumers have retired it: every MMA warp's
  // commit arrives on the stage's empty barrier of every CTA.  `live` is false for the CTAs that pad an uneven
  // user-tile count: they load their share and release stages, nothing else.
  const bool mc = p.cl > 1 && !(p.dbg & 128);              // experiment 128: cluster launch, unicast loads
  const uint32_t cta_rank = mc ? cluster_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << p.cl) - 1u);
  const uint32_t mc_rows = (uint32_t)GQ_TILE_I / (uint32_t)(mc ? p.cl : 1);
  const bool live = u_tile * GQ_TILE_U < p.B;
#ifdef LGX_GQ_PROF
  unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long pstart = clock64();
  if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1024) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    gq_cta[blockIdx.x * 8 + 0] = smid;
    gq_cta[blockIdx.x * 8 + 1] = gq_gtime();
    gq_cta[blockIdx.x * 8 + 6] = clock64();
  }
#endif

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_u)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_i)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, mc ? p.cl : 1);
    }
    mbar_init(bar_a, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, GQ_EPI / 32);     // one arrival per epilogue warp
      mbar_init(bar_mfull + 8 * b, 1);
      mbar_init(bar_mfree + 8 * b, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < GQ_EPI) {
    thr_all[threadIdx.x] = -CUDART_INF_F;
    if (p.union_bound) quart_all[threadIdx.x] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  }
  if (p.has_mask) {
    // mask operands start as zeros: (GQ_A_BLOCK + GQ_STAGE) / 16 = 3072 16-byte stores over 384 threads
    for (uint32_t o = threadIdx.x * 16u; o < (uint32_t)(GQ_A_BLOCK + GQ_STAGE); o += GQ_THREADS * 16u)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sAm + o), "r"(0u) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (mc) cluster_sync_all();           // the peers' barriers exist before anything of ours can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (n_my > 0 && (live || mc)) {
      // ---------------------------------------------------------------- TMA producer (converged warp, one lane issues)
      if (live && elect_one()) {
        mbar_expect_tx(bar_a, (uint32_t)p.k_blocks * GQ_A_BLOCK);
        for (int kb = 0; kb < p.k_blocks; ++kb)
          tma_load_2d(sA + kb * GQ_A_BLOCK, &tmap_u, bar_a, kb * GQ_KBLK, u_tile * GQ_TILE_U);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_my; ++it) {
        const int row0 = tile_of(it) * GQ_TILE_I;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          PROF_T(tw0);
          mbar_wait<false>(bar_empty + 8 * stage, phase ^ 1);
          PROF_T(tw1);
          PROF_ADD(0, tw1 - tw0);
          if (elect_one()) {
            mbar_expect_tx(bar_full + 8 * stage, GQ_STAGE);         // our rows + the peers' (which may land first)
            if (mc)
              tma_load_2d_mc(sB + stage * GQ_STAGE + cta_rank * mc_rows * 128u, &tmap_i, bar_full + 8 * stage,
                             kb * GQ_KBLK, row0 + (int)(cta_rank * mc_rows), mc_mask);
            else
              tma_load_2d(sB + stage * GQ_STAGE, &tmap_i, bar_full + 8 * stage, kb * GQ_KBLK, row0);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
#ifdef LGX_GQ_PROF
      if (lane == 0) atomicAdd(&gq_prof[16], prof_acc[0]);
#endif
    }
  } else if (warp == 1) {
    if (n_my > 0 && !live) {
      if (mc) {
        // padding CTA: the stages it receives are released unread (a commit with no MMA before it arrives at once)
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_my * p.k_blocks; ++j) {
          mbar_wait<false>(bar_full + 8 * stage, phase);
          if (elect_one()) tc_commit_mc(bar_empty + 8 * stage, mc_mask);
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (n_my > 0) {
      // ---------------------------------------------------------------- MMA issuer (converged warp, one lane issues)
      // This warp is the serial resource of the pipeline: everything it needs per tile is kept in registers and
      // advanced incrementally (barrier addresses, the B descriptor, phases), and the tile loop body exists twice so
      // the TMEM buffer is static.
      mbar_wait<false>(bar_a, 0);
      tc_fence_after();
      const uint64_t adesc_m = umma_desc_sw128(sAm), bdesc_m = umma_desc_sw128(sBm);
      const uint64_t adesc0 = umma_desc_sw128(sA);
      const uint64_t bdesc0 = umma_desc_sw128(sB);
      int stage = 0;
      uint32_t phase = 0;               // parity of the B ring
      uint32_t mr0 = 0u, mr1 = 0u;      // mask rounds consumed per buffer (tile `it` uses buffer it & 1 = its TMEM buffer)
      // ready(it): everything tile `it` needs before its first MMA -- its TMEM buffer drained (tempty), its mask round
      // published (mfull; returns the round's count word) and its first B block landed (full).  All three are
      // normally complete long before they are looked at, but a completed try_wait still costs ~85 cycles and the
      // tensor pipe queues only about two MMAs: checked between tiles, the pipe ran dry while this warp polled
      // (~350 of a tile's ~1 450 cycles).  So tile it + 1 is checked in the MIDDLE of tile it's last K block, behind
      // two queued MMAs -- unless the ring is so short that the next tile's first block cannot have been requested yet.
      const bool pipelined = p.stages >= p.k_blocks + 2;
      auto ready = [&](int it, const int buf, int st, uint32_t ph) -> int {
        PROF_T(m0);
        mbar_wait<false>(bar_tempty + 8 * buf, (uint32_t)((it >> 1) & 1) ^ 1u);
        PROF_T(w1);
        int v = 0;
        if (p.has_mask) {
          mbar_wait<false>(bar_mfull + 8 * buf, (buf ? mr1 : mr0) & 1u);
          v = mcnt[buf];
        }
        PROF_T(w2);
        mbar_wait<false>(bar_full + 8 * st, ph);
        tc_fence_after();
        PROF_T(w3);
        PROF_ADD(0, w3 - m0); PROF_ADD(3, w1 - m0); PROF_ADD(4, w2 - w1); PROF_ADD(5, w3 - w2); PROF_ADD(6, 1);
        TRACE(0, it, 0, m0); TRACE(0, it, 1, w1); TRACE(0, it, 2, w2); TRACE(0, it, 3, w3);
        return v;
      };
      auto tile = [&](int it, const int buf, int v) -> int {      // returns ready(it + 1) when pipelined
        const uint32_t tmem_d = tmem_base + (uint32_t)buf * GQ_TMEM_BUF;
        uint32_t acc = 0u;
        if (!pipelined) v = ready(it, buf, stage, phase);
        PROF_T(m2);
        if (p.has_mask) {
          // The tile's train mask first: one K = 16 step per 16 published slots puts -2^100 into the (row, train
          // column) accumulators; the real K steps then accumulate on top.  Mask first, so a buffer is released as
          // soon as its own MMAs retire instead of behind the tile's real MMAs.
          const uint32_t b = (uint32_t)buf;
          uint32_t mr = buf ? mr1 : mr0;
          for (;;) {
            ++mr;
            const int n = v & 255;
            if (elect_one()) {
              if (n > 0) tc_mma_f16(tmem_d, adesc_m + (uint64_t)(4 * b), bdesc_m + (uint64_t)(4 * b), GQ_IDESC, acc);
              if (n > 16) tc_mma_f16(tmem_d, adesc_m + (uint64_t)(4 * b + 2), bdesc_m + (uint64_t)(4 * b + 2), GQ_IDESC, 1u);
              tc_commit(bar_mfree + 8 * b);
            }
            __syncwarp();
            if (n > 0) acc = 1u;
            if (!(v >> 8)) break;
            mbar_wait<false>(bar_mfull + 8 * b, mr & 1u);     // a tile with more than 32 dirty rows: further rounds
            tc_fence_after();
            v = mcnt[b];
          }
          if (buf) mr1 = mr; else mr0 = mr;
        }
        int v_next = 0;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (kb > 0) {
            mbar_wait<false>(bar_full + 8 * stage, phase);
            tc_fence_after();
          }
          const bool last = kb == p.k_blocks - 1;
          const uint64_t adesc = adesc0 + (uint64_t)(kb * (GQ_A_BLOCK >> 4));
          const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (GQ_STAGE >> 4));
          if (elect_one()) {
            tc_mma_f16(tmem_d, adesc, bdesc, GQ_IDESC, acc);
            if (LGX_GQ_READY_AFTER >= 2) tc_mma_f16(tmem_d, adesc + 2, bdesc + 2, GQ_IDESC, 1u);
            if (LGX_GQ_READY_AFTER >= 3) tc_mma_f16(tmem_d, adesc + 4, bdesc + 4, GQ_IDESC, 1u);
          }
          __syncwarp();
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == p.stages) { nstage = 0; nphase ^= 1; }
          if (pipelined && last && it + 1 < n_my) v_next = ready(it + 1, buf ^ 1, nstage, nphase);
          if (elect_one()) {
            if (LGX_GQ_READY_AFTER < 2) tc_mma_f16(tmem_d, adesc + 2, bdesc + 2, GQ_IDESC, 1u);
            if (LGX_GQ_READY_AFTER < 3) tc_mma_f16(tmem_d, adesc + 4, bdesc + 4, GQ_IDESC, 1u);
            tc_mma_f16(tmem_d, adesc + 6, bdesc + 6, GQ_IDESC, 1u);
            if (mc) tc_commit_mc(bar_empty + 8 * stage, mc_mask);
            else tc_commit(bar_empty + 8 * stage);
            if (last) tc_commit(bar_tfull + 8 * buf);   // accumulator tile complete
          }
          __syncwarp();
          acc = 1u;
          stage = nstage;
          phase = nphase;
        }
        PROF_T(m3);
        PROF_ADD(2, m3 - m2);
        TRACE(0, it, 4, m2); TRACE(0, it, 5, m3);
        return v_next;
      };
      PROF_T(mstart);
#ifdef LGX_GQ_PROF
      if (lane == 0 && blockIdx.y == 0 && blockIdx.x < 1024) gq_cta[blockIdx.x * 8 + 2] = gq_gtime();
#endif
      int v = pipelined ? ready(0, 0, 0, 0u) : 0;
      for (int it = 0; it < n_my; it += 2) {
        v = tile(it, 0, v);
        if (it + 1 < n_my) v = tile(it + 1, 1, v);
      }
#ifdef LGX_GQ_PROF
      PROF_T(mend);
      if (lane == 0 && blockIdx.y == 0 && blockIdx.x < 1024) gq_cta[blockIdx.x * 8 + 3] = gq_gtime();
      if (lane == 0) {
        atomicAdd(&gq_prof[0], prof_acc[0]); atomicAdd(&gq_prof[1], (unsigned long long)(mend - mstart)); atomicAdd(&gq_prof[2], prof_acc[2]);
        atomicAdd(&gq_prof[3], prof_acc[3]); atomicAdd(&gq_prof[4], prof_acc[4]); atomicAdd(&gq_prof[5], prof_acc[5]);
        atomicAdd(&gq_prof[6], prof_acc[6]);
      }
#endif
    }
  } else if (warp < 4) {
    // -------------------------------------------------------------------- mask builders (warps 3 and 2)
    // Tile `it` uses mask buffer it & 1.  With pre-bucketed entries the two warps take alternate tiles, so each owns
    // one buffer and has two tile periods for a tile's work (one warp needed ~1 600 cycles per tile against a
    // 2 000-cycle tile period and was the longest chain of the pipeline); the list-walking fallback keeps one
    // cursor set and runs on warp 3 alone over both buffers.
    if (p.has_mask && n_my > 0 && live) {
      const unsigned lt = (1u << lane) - 1u;
      const int bw = warp == 3 ? 0 : 1;   // builder index
      uint32_t rc0 = 0, rc1 = 0;          // rounds published per buffer
#ifdef LGX_GQ_PROF
      int tr_it = bw;                     // trace only: the tile being built (one round per tile assumed)
#endif
      // what this lane wrote into buffer b the last time (shared-memory addresses, 0 = nothing), or the whole
      // buffer was written by a multi-pass round and is cleared wholesale
      uint32_t uA0 = 0, uB0 = 0, uA1 = 0, uB1 = 0;     // first entry a lane wrote
      uint32_t uC0 = 0, uD0 = 0, uC1 = 0, uD1 = 0;     // second entry (tiles with 33-64 entries)
      bool wide0 = false, wide1 = false;
      // zero what the previous use of buffer b wrote; then every lane may write again
      auto reclaim = [&](uint32_t b) {
        PROF_T(b0);
        mbar_wait<false>(bar_mfree + 8 * b, ((b ? rc1 : rc0) & 1u) ^ 1u);   // the MMAs that read this buffer have retired
        PROF_T(b1);
        PROF_ADD(0, b1 - b0);
        TRACE(1 + bw, tr_it, 0, b0); TRACE(1 + bw, tr_it, 1, b1);
        if (b) ++rc1; else ++rc0;
        const bool wide = b ? wide1 : wide0;
        if (wide) {
          // both K steps of the buffer, all 384 operand rows: 2 x 2 16-byte chunks per row
          for (int i = lane; i < (GQ_TILE_U + GQ_TILE_I) * 4; i += 32) {
            const int rrow = i >> 2, ch = (int)(4 * b) + (i & 3);
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};"
                         ::"r"(sAm + (uint32_t)rrow * 128u + (uint32_t)((ch ^ (rrow & 7)) << 4)), "r"(0u) : "memory");
          }
        } else {
          const uint32_t a = b ? uA1 : uA0, c = b ? uB1 : uB0;
          if (c) { st_shared_u16(a, 0); st_shared_u16(c, 0); }
          const uint32_t a2 = b ? uC1 : uC0, c2 = b ? uD1 : uD0;
          if (c2) { st_shared_u16(a2, 0); st_shared_u16(c2, 0); }
        }
        if (b) { uA1 = uB1 = uC1 = uD1 = 0; wide1 = false; } else { uA0 = uB0 = uC0 = uD0 = 0; wide0 = false; }
        __syncwarp();     // a slot changes owner between rounds: all zeroing before any new write
      };
      auto publish = [&](uint32_t b, int n, bool more) {
        if (lane == 0) mcnt[b] = ((p.dbg & 8) ? 0 : n) | (more ? 256 : 0);
        if (!(p.dbg & 64))                 // experiment 64: no proxy fence (timing only, results are wrong)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_mfull + 8 * b);
#ifdef LGX_GQ_PROF
        TRACE(1 + bw, tr_it, 2, clock64());
        tr_it += 2;
#endif
      };
      const int64_t region = p.mk_region ? p.mk_region[u_tile] : -1;
      if (p.dbg & 16) {                   // experiment 16: handshake only, nothing is built
        for (int it = bw; it < n_my; it += 2) {
          const uint32_t b = (uint32_t)bw;
          reclaim(b);
          publish(b, 0, false);
        }
      } else if (region >= 0) {
        // ---- fast path: the tile's entries are pre-bucketed; entries and offsets are fetched two tiles ahead
        const int32_t* ptrw = p.mk_ptr + (int64_t)u_tile * (n_tiles + 1);
        const uint16_t* ents = p.mk_entries + region;
        int wbase = -(1 << 30), pw = 0;
        auto getptr = [&](int t) {        // warp-uniform; a 32-tile window of offsets lives in one register per lane
          if (t - wbase >= 32 || t < wbase) { wbase = t; pw = __ldg(ptrw + min(t + lane, n_tiles)); }
          return __shfl_sync(0xffffffffu, pw, t - wbase);
        };
        // Three prefetch slots used round-robin: a slot's entry load is issued three of this warp's tiles before its
        // use and never moved between registers (a rotating copy made every iteration wait for its own load).
        struct Pre { int e0, n; uint32_t ent, ent2; };
        Pre s0{0, 0, 0u, 0u}, s1{0, 0, 0u, 0u}, s2{0, 0, 0u, 0u};
        auto fetch = [&](Pre& q, int it2) {
          q.e0 = 0; q.n = 0; q.ent = 0u; q.ent2 = 0u;
          if (it2 < n_my) {
            const int t = tile_of(it2);
            q.e0 = getptr(t);
            q.n = getptr(t + 1) - q.e0;
            if (lane < q.n) q.ent = __ldg(ents + q.e0 + lane);
            if (lane + 32 < q.n) q.ent2 = __ldg(ents + q.e0 + 32 + lane);
          }
        };
        const uint32_t b = (uint32_t)bw;              // this warp's tiles all use its own buffer
        // A slot is a star of the tile's (row, train column) graph: centred on a ROW (the row's -2^100 in A_mask, 1.0 in
        // B_mask for each of its train columns) or on a COLUMN (1.0 in B_mask, -2^100 in A_mask for each row that
        // trained on it).  Row stars unless the tile has more than 32 dirty rows and fewer dirty columns: such tiles
        // hold a popular item (Amazon-Book shape: 2.8 % of the tiles, 46 dirty rows but 17 dirty columns on average)
        // and cost two builder <-> MMA round trips (~6 000 cycles, ten times per CTA) when ranked by row.
        auto process = [&](const Pre& q) {
          const int n_e = q.n, e0 = q.e0;
          if (n_e <= 64) {
            // up to two entries per lane, all in registers
            const bool v0 = lane < n_e, v1 = lane + 32 < n_e;
            const int row0 = (int)(q.ent >> 8), col0 = (int)(q.ent & 255u);
            const int row1 = (int)(q.ent2 >> 8), col1 = (int)(q.ent2 & 255u);
            unsigned m[4];
#pragma unroll
            for (int w = 0; w < 4; ++w)
              m[w] = __reduce_or_sync(0xffffffffu, ((v0 && (row0 >> 5) == w) ? 1u << (row0 & 31) : 0u) |
                                                       ((v1 && (row1 >> 5) == w) ? 1u << (row1 & 31) : 0u));
            int n_dirty = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
            int rank0 = gq_bits_rank<4>(m, row0), rank1 = gq_bits_rank<4>(m, row1);
            if (n_dirty > 32) {
              unsigned cm[8];
              int n_cols = 0;
#pragma unroll
              for (int w = 0; w < 8; ++w) {
                cm[w] = __reduce_or_sync(0xffffffffu, ((v0 && (col0 >> 5) == w) ? 1u << (col0 & 31) : 0u) |
                                                          ((v1 && (col1 >> 5) == w) ? 1u << (col1 & 31) : 0u));
                n_cols += __popc(cm[w]);
              }
              if (n_cols < n_dirty) {
                n_dirty = n_cols;
                rank0 = gq_bits_rank<8>(cm, col0);
                rank1 = gq_bits_rank<8>(cm, col1);
              }
            }
            const int nrounds = n_dirty > 32 ? (n_dirty + 31) >> 5 : 1;
            for (int r = 0; r < nrounds; ++r) {
              reclaim(b);
              if (v0 && (rank0 >> 5) == r) {
                const int slot = 32 * (int)b + (rank0 & 31);
                const uint32_t a = sAm + gq_swz(row0, slot), c = sBm + gq_swz(col0, slot);
                st_shared_u16(a, GQ_BF16_NEG_BIG);      // entries of one star write the centre's value several times
                st_shared_u16(c, GQ_BF16_ONE);
                if (b) { uA1 = a; uB1 = c; } else { uA0 = a; uB0 = c; }
              }
              if (v1 && (rank1 >> 5) == r) {
                const int slot = 32 * (int)b + (rank1 & 31);
                const uint32_t a = sAm + gq_swz(row1, slot), c = sBm + gq_swz(col1, slot);
                st_shared_u16(a, GQ_BF16_NEG_BIG);
                st_shared_u16(c, GQ_BF16_ONE);
                if (b) { uC1 = a; uD1 = c; } else { uC0 = a; uD0 = c; }
              }
              publish(b, min(32, n_dirty - 32 * r), r + 1 < nrounds);
            }
          } else {
            // more than 64 train entries in one 128 x 256 tile: pass 1 collects the dirty rows and columns, every
            // round re-reads the entries it needs; the buffer is cleared wholesale afterwards
            unsigned m[4] = {0u, 0u, 0u, 0u}, cm[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            for (int cbase = 0; cbase < n_e; cbase += 32) {
              const bool valid = cbase + lane < n_e;
              const uint32_t e = valid ? (uint32_t)__ldg(ents + e0 + cbase + lane) : 0u;
              const int row = (int)(e >> 8), col = (int)(e & 255u);
#pragma unroll
              for (int w = 0; w < 4; ++w)
                m[w] |= __reduce_or_sync(0xffffffffu, (valid && (row >> 5) == w) ? 1u << (row & 31) : 0u);
#pragma unroll
              for (int w = 0; w < 8; ++w)
                cm[w] |= __reduce_or_sync(0xffffffffu, (valid && (col >> 5) == w) ? 1u << (col & 31) : 0u);
            }
            int n_dirty = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]), n_cols = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) n_cols += __popc(cm[w]);
            const bool by_col = n_dirty > 32 && n_cols < n_dirty;
            if (by_col) n_dirty = n_cols;
            const int nrounds = n_dirty > 32 ? (n_dirty + 31) >> 5 : 1;
            for (int r = 0; r < nrounds; ++r) {
              reclaim(b);
              for (int cbase = 0; cbase < n_e; cbase += 32) {
                if (cbase + lane < n_e) {
                  const uint32_t e = (uint32_t)__ldg(ents + e0 + cbase + lane);
                  const int row = (int)(e >> 8), col = (int)(e & 255u);
                  const int rank = by_col ? gq_bits_rank<8>(cm, col) : gq_bits_rank<4>(m, row);
                  if ((rank >> 5) == r) {
                    const int slot = 32 * (int)b + (rank & 31);
                    st_shared_u16(sAm + gq_swz(row, slot), GQ_BF16_NEG_BIG);
                    st_shared_u16(sBm + gq_swz(col, slot), GQ_BF16_ONE);
                  }
                }
              }
              if (b) wide1 = true; else wide0 = true;
              publish(b, min(32, n_dirty - 32 * r), r + 1 < nrounds);
            }
          }
        };
        // ONE copy of process() (a 3x unrolled loop tripled ~700 instructions of rarely executed code and the warp
        // ran out of the instruction cache): the slot to use is selected with warp-uniform selects, and the slot just
        // consumed is refilled by loads straight into its own registers, three of this warp's tiles ahead.
        fetch(s0, bw);
        fetch(s1, bw + 2);
        fetch(s2, bw + 4);
        int k = 0;
#pragma unroll 1
        for (int it = bw; it < n_my; it += 2) {
          Pre cur;
          cur.e0 = k == 0 ? s0.e0 : (k == 1 ? s1.e0 : s2.e0);
          cur.n = k == 0 ? s0.n : (k == 1 ? s1.n : s2.n);
          cur.ent = k == 0 ? s0.ent : (k == 1 ? s1.ent : s2.ent);
          cur.ent2 = k == 0 ? s0.ent2 : (k == 1 ? s1.ent2 : s2.ent2);
          if (k == 0) fetch(s0, it + 6);
          else if (k == 1) fetch(s1, it + 6);
          else fetch(s2, it + 6);
          k = k == 2 ? 0 : k + 1;
          process(cur);
        }
      } else if (bw == 0) {
        // ---- fallback: walk the 128 sorted train lists here (4 rows per lane, one dependent load per train item)
        GqCursor cur[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int u = u_tile * GQ_TILE_U + x * 32 + lane;
          const int64_t uid = (u < p.B) ? (p.users ? p.users[u] : (int64_t)u) : -1;
          cur[x].init(p.mask, uid, p.item_offset, t_begin * GQ_TILE_I);
        }
        for (int it = 0; it < n_my; ++it) {
          const int tile_lo = (t_begin + it) * GQ_TILE_I, tile_hi = tile_lo + GQ_TILE_I;
          int rank[4], n = 0;
          bool dirty[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            dirty[x] = cur[x].next < tile_hi;
            const unsigned bal = __ballot_sync(0xffffffffu, dirty[x]);
            rank[x] = n + __popc(bal & lt);
            n += __popc(bal);
          }
          const int nrounds = n > 32 ? (n + 31) >> 5 : 1;
          const uint32_t b = (uint32_t)(it & 1);
          for (int r = 0; r < nrounds; ++r) {
            reclaim(b);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              if (dirty[x] && (rank[x] >> 5) == r) {
                const int slot = 32 * (int)b + (rank[x] & 31);
                st_shared_u16(sAm + gq_swz(x * 32 + lane, slot), GQ_BF16_NEG_BIG);
                while (cur[x].next < tile_hi) {
                  st_shared_u16(sBm + gq_swz(cur[x].next - tile_lo, slot), GQ_BF16_ONE);
                  cur[x].advance();
                }
              }
            }
            if (b) wide1 = true; else wide0 = true;
            publish(b, min(32, n - 32 * r), r + 1 < nrounds);
          }
        }
      }
    }
#ifdef LGX_GQ_PROF
    if (p.has_mask && n_my > 0 && live && lane == 0) {
      atomicAdd(&gq_prof[12], prof_acc[0]);
      atomicAdd(&gq_prof[13], (unsigned long long)(clock64() - pstart));
    }
#endif
  } else if (live) {
    // ------------------------------------------------------------------ epilogue (256 threads)
    const int q = hw_warp & 3;                // TMEM lane quarter this warp may access (hardware warp id % 4)
    const int h = hw_warp >> 2;               // which 128-column half of the tile
    const int row = q * 32 + lane;
    const int col = h * GQ_TILE_U + row;
    GqEpi<KMAX, SHARE> st;
    st.init(p.K);
    st.qb = base + off_queue + (uint32_t)col * 8u;
    st.qn = st.qh = 0;
    st.thr_mine = thr_all + col;
    st.thr_other = thr_all + (h ^ 1) * GQ_TILE_U + row;
    st.quart_mine = quart_all + col;
    st.quart_other = quart_all + (h ^ 1) * GQ_TILE_U + row;
    st.use_union = p.K == KMAX && p.union_bound;
    st.gbound = (SHARE && u_tile * GQ_TILE_U + row < p.B) ? p.row_bound + (u_tile * GQ_TILE_U + row) : nullptr;
    if (SHARE) st.flush();
    const int u = u_tile * GQ_TILE_U + row;
    for (int it = 0; it < n_my; ++it) {
      const int buf = it & 1;
      PROF_T(e0);
      mbar_wait<false>(bar_tfull + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      PROF_T(e1);
      PROF_ADD(0, e1 - e0);
      TRACE(4 + hw_warp, it, 0, e0); TRACE(4 + hw_warp, it, 1, e1);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * GQ_TMEM_BUF + h * 128);
      const int gid_base = (tile_of(it) * GQ_TILE_I + h * 128) / GQ_GROUP;
      if (p.dbg & 2) {                        // experiment: barrier handshake only (TMA / MMA / builder floor)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        continue;
      }
      float th = (p.dbg & 1) ? CUDART_INF_F : st.filter();   // experiment 1: loads + max tree, nothing appended
      // Three register buffers: the loads of the first three chunks are issued back to back and the fourth as soon as
      // the first chunk is consumed, so the TMEM buffer goes back to the MMA warp after ONE chunk of epilogue work
      // instead of three.  The accumulator round trip (epilogue hold time + MMA latency) over two TMEM buffers is what
      // bounds this kernel: a tile is only 512 tensor cycles of work (K = d = 64).  LGX_KEEP pins a buffer's values
      // across the issue of later loads (otherwise the compiler folds the buffers into one set of registers).
      uint32_t va[32], vb[32], vc[32];
      LGX_TMEM_LD32(va, taddr);
      LGX_TMEM_LD32(vb, taddr + 32);
      LGX_TMEM_LD32(vc, taddr + 64);
      LGX_TMEM_WAIT(va);
      LGX_KEEP(vb);
      LGX_KEEP(vc);
      // flush checks: a lane appends at most APC = 32 / GQ_GROUP candidates per chunk; small queues are checked with
      // a margin of 8 appends (every 2 chunks at 8-column groups), deep ones once per tile with a margin of 16
      constexpr int APC = 32 / GQ_GROUP;
      // chunks between two checks; an 8-row queue must survive the appends in between (margin <= 8)
      constexpr int EVERY = SMALLQ ? (LGX_GQ_EVERY * APC <= 8 ? LGX_GQ_EVERY : 8 / APC) : 4;
      constexpr int MARGIN = EVERY * APC;
      static_assert(!LGX_GQ_EARLY_CHUNK || EVERY >= 2, "two chunks are processed before the first check");
      auto check = [&](int chunks_done) {
        if (chunks_done % EVERY == 0) {
          __syncwarp();
          // one insertion loop per check site (the drain loop, unbounded when a queue needs room for the next
          // appends): every jump over an inlined loop body costs an instruction-fetch bubble on the common path
          const bool room = __any_sync(0xffffffffu, st.fuller_than(p.q_cap - MARGIN));
          if (room || (LGX_GQ_DRAIN > 0 && __any_sync(0xffffffffu, st.pending_at_least(LGX_GQ_LOW)))) {
            st.drain(room ? (1 << 20) : LGX_GQ_DRAIN);
            if (LGX_GQ_TH_REFRESH && !(p.dbg & 1)) th = st.filter();     // the rest of the tile filters with the new bound
          }
        }
      };
      gq_chunk<KMAX, SHARE>(va, gid_base, st, th);
      LGX_TMEM_LD32(va, taddr + 96);
#if LGX_GQ_EARLY_CHUNK
      // the second chunk is processed while the fourth load is in flight (the wait right behind the load exposed
      // its whole latency once per tile)
      LGX_KEEP(vb);
      gq_chunk<KMAX, SHARE>(vb, gid_base + 1 * (32 / GQ_GROUP), st, th);
#endif
      LGX_TMEM_WAIT(va);
      tc_fence_before();
      __syncwarp();
      // ONE arrival per warp: 256 per-thread arrivals are 256 serialised shared-memory atomics on the path that
      // hands the accumulator buffer back to the MMA warp (the round trip that bounds this kernel)
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);      // the whole tile half is in registers
      PROF_T(e2);
      PROF_ADD(1, e2 - e1);
      TRACE(4 + hw_warp, it, 2, e2);
      check(1);
      LGX_KEEP(vc);
#if !LGX_GQ_EARLY_CHUNK
      LGX_KEEP(vb);
      gq_chunk<KMAX, SHARE>(vb, gid_base + 1 * APC, st, th);
#endif
      check(2);
      gq_chunk<KMAX, SHARE>(vc, gid_base + 2 * APC, st, th);
      check(3);
      gq_chunk<KMAX, SHARE>(va, gid_base + 3 * APC, st, th);
      check(4);
      PROF_T(e3);
      PROF_ADD(2, e3 - e2);
      TRACE(4 + hw_warp, it, 3, e3);
    }
#ifdef LGX_GQ_PROF
    if (lane == 0 && hw_warp == 0 && blockIdx.y == 0 && blockIdx.x < 1024) gq_cta[blockIdx.x * 8 + 4] = gq_gtime();
    if (lane == 0) {      // one lane of every epilogue warp: divide by 8 x tiles
      atomicAdd(&gq_prof[8], prof_acc[0]); atomicAdd(&gq_prof[9], prof_acc[1]); atomicAdd(&gq_prof[10], prof_acc[2]);
    }
#endif
    st.flush();
    // stage the register lists, merge the two halves of every row and publish the split's K best groups
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      lval_all[k * GQ_EPI + col] = st.lv[k];
      lidx_all[k * GQ_EPI + col] = st.li[k];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(GQ_EPI) : "memory");
    if (h == 0 && u < p.B) {
      int pa = KMAX - p.K, pb = KMAX - p.K;
      const int64_t o = ((int64_t)split * p.B + u) * p.K;
      for (int k = 0; k < p.K; ++k) {
        const float av = lval_all[pa * GQ_EPI + row], bv = lval_all[pb * GQ_EPI + GQ_TILE_U + row];
        const int32_t ai = lidx_all[pa * GQ_EPI + row], bi = lidx_all[pb * GQ_EPI + GQ_TILE_U + row];
        const bool take_b = better(bv, bi, av, ai);
        p.ws_val[o + k] = take_b ? bv : av;
        p.ws_idx[o + k] = take_b ? bi : ai;
        pa += take_b ? 0 : 1;
        pb += take_b ? 1 : 0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (mc) cluster_sync_all();           // no CTA leaves while a peer may still multicast into it or signal its barriers
#ifdef LGX_GQ_PROF
  if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1024) {
    gq_cta[blockIdx.x * 8 + 5] = gq_gtime();
    gq_cta[blockIdx.x * 8 + 7] = clock64();
  }
#endif
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ rescoring
// One warp per batch row.  The row's candidate lists hold (group maximum, group id); T = the largest K-th entry
// over the row's lists is a lower bound on the row's K-th best score, so only groups with maximum >= T can hold
// a top-K item (for a single list: all K of them).  Their 8 items each are scored again from the same bf16 operands
// with fp32 accumulation -- as a [16 items x k] x [k x 8] mma.sync per pair of groups, the user vector replicated
// over the 8 columns: an fp32 FMA loop with bf16 unpacking measured 330 us for the Amazon-Book pass, 75 % of it
// integer unpack instructions.  Train items are dropped by binary search in the row's sorted train list, items
// within `margin` of T are collected and ranked by counting.
struct GqRescoreParams {
  const float* ws_val; const int32_t* ws_idx;
  int P, B, K, M, ktot;
  const __nv_bfloat16* U_op; const __nv_bfloat16* I_op;
  TrainMask mask; const int64_t* users;
  int64_t item_offset;
  int64_t* out_idx; float* out_val;
};

constexpr int RS_WARPS = 8;
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 3          // 85 registers, no spills (4 blocks = 64 registers spilled 72 bytes per thread; same speed)
#endif
constexpr int RS_CAP = 160;      // candidate buffer per warp: compressed to the best K whenever > RS_CAP - 64 are held
constexpr int RS_KMAX = 32;
constexpr int RS_KTOT_MAX = 384;
constexpr int RS_GLIST = 64;     // selected groups staged per warp

// rank-by-counting compression of buf[0..n) to its best min(n, K) entries, sorted best-first (score desc, id asc)
__device__ __forceinline__ int rs_compress(float* bv, int32_t* bi, float* tv, int32_t* ti, int n, int K, int lane) {
  for (int e = lane; e < n; e += 32) {
    const float v = bv[e];
    const int32_t i = bi[e];
    int rank = 0;
    for (int f = 0; f < n; ++f) rank += better(bv[f], bi[f], v, i) ? 1 : 0;
    if (rank < K) { tv[rank] = v; ti[rank] = i; }
  }
  __syncwarp();
  const int m = min(n, K);
  for (int e = lane; e < m; e += 32) { bv[e] = tv[e]; bi[e] = ti[e]; }
  __syncwarp();
  return m;
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct RsRow {
  float* bv; int32_t* bi; float* tv; int32_t* ti;
  int n;
};

// Collect the items that can still reach the top-K.  Train items are NOT filtered here: a binary search is six dependent global loads, and doing one per tile
// (10 per row) was most of the first version's 250 us; rs_drop_masked checks all collected items in one pass.
__device__ __forceinline__ void rs_push(const GqRescoreParams& rp, float sc, int64_t j, bool ok, float T_lo, RsRow& row,
                                        int lane) {
  const bool keep = ok && j < rp.M && sc >= T_lo;
  const unsigned kbits = __ballot_sync(0xffffffffu, keep);
  if (keep) {
    const int pos = row.n + __popc(kbits & ((1u << lane) - 1u));
    row.bv[pos] = sc;
    row.bi[pos] = (int32_t)j;
  }
  row.n += __popc(kbits);
}

// remove the row's train items from the buffer (one binary search per buffered item, all in parallel)
__device__ __forceinline__ void rs_drop_masked(const GqRescoreParams& rp, int64_t uid, RsRow& row, int lane) {
  if (rp.mask.indptr == nullptr) return;
  int kept = 0;
  for (int e0 = 0; e0 < row.n; e0 += 32) {
    const int e = e0 + lane;
    float v = 0.f;
    int32_t i = 0;
    bool keep = false;
    if (e < row.n) {
      v = row.bv[e];
      i = row.bi[e];
      keep = !rp.mask.contains(uid, rp.item_offset + i);
    }
    const unsigned kb = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) {
      const int pos = kept + __popc(kb & ((1u << lane) - 1u));     // pos <= e: in-place compaction is safe per 32-batch
      row.bv[pos] = v;
      row.bi[pos] = i;
    }
    kept += __popc(kb);
    __syncwarp();
  }
  row.n = kept;
}

// mma.sync m16n8k16 with the ITEMS as the B operand (n = 8 items of one group, so one 16-byte load per lane is
// exactly two steps' {b0, b1} -- no register shuffling) and the user vector replicated over the 16 rows of A.
// Lane (g = lane / 4, t = lane % 4) loads the 8 consecutive k [32 kb + 8t, +8) of item g; they feed logical k
// {2t, 2t+1, 2t+8, 2t+9} of two k16 steps, and the user fragment uses the same permutation, so the dot product is
// unchanged.  D[row][2t], D[row][2t+1] = scores of items 2t, 2t+1 of the group in every lane of every quad.
__device__ __forceinline__ void rs_process(const GqRescoreParams& rp, const int32_t* glist, int cnt, const uint4* us,
                                           int64_t uid, float T_lo, RsRow& row, int lane) {
  constexpr int GPS = 8 / GQ_GROUP;         // groups per MMA slot (the 8 items of one n = 8 B operand)
  const int g = lane >> 2, t = lane & 3;
  const int kb32 = rp.ktot >> 5;            // even: ktot is a multiple of 64
  for (int m = 0; m < cnt; m += 4 * GPS) {  // four slots per pass
    const uint4* pr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gi = m + q * GPS + g / GQ_GROUP;
      const int32_t gid = glist[gi < cnt ? gi : m];
      // rows past the end of the catalogue are clamped for the load and dropped in rs_push
      const int64_t j = min((int64_t)gid * GQ_GROUP + (g % GQ_GROUP), (int64_t)rp.M - 1);
      pr[q] = reinterpret_cast<const uint4*>(rp.I_op + j * rp.ktot) + t;
    }
    float c[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f;
    for (int kb = 0; kb < kb32; kb += 2) {
      uint4 x[2][4];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int q = 0; q < 4; ++q) x[jj][q] = __ldg(pr[q] + 4 * (kb + jj));
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const uint4 xu = us[4 * (kb + jj) + t];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          mma_bf16_16816(c[q], xu.x, xu.x, xu.y, xu.y, x[jj][q].x, x[jj][q].y);
          mma_bf16_16816(c[q], xu.z, xu.z, xu.w, xu.w, x[jj][q].z, x[jj][q].w);
        }
      }
    }
    // quad g < 4 publishes slot g: its lane t holds the slot's items 2t and 2t+1
    const float sa = g == 0 ? c[0][0] : (g == 1 ? c[1][0] : (g == 2 ? c[2][0] : c[3][0]));
    const float sb = g == 0 ? c[0][1] : (g == 1 ? c[1][1] : (g == 2 ? c[2][1] : c[3][1]));
    const int gi = m + g * GPS + (2 * t) / GQ_GROUP;
    const bool ok = g < 4 && gi < cnt;
    const int64_t j0 = (int64_t)glist[ok ? gi : m] * GQ_GROUP + (2 * t) % GQ_GROUP;
    rs_push(rp, sa, j0, ok, T_lo, row, lane);
    rs_push(rp, sb, j0 + 1, ok, T_lo, row, lane);
    __syncwarp();
    if (row.n > RS_CAP - 64) {
      rs_drop_masked(rp, uid, row, lane);
      row.n = rs_compress(row.bv, row.bi, row.tv, row.ti, row.n, rp.K, lane);
    }
  }
}

__global__ void __launch_bounds__(RS_WARPS * 32, RS_MIN_BLOCKS)
k_rescore_topk(const GqRescoreParams rp) {
  __shared__ __align__(16) __nv_bfloat16 s_user[RS_WARPS][RS_KTOT_MAX];
  __shared__ float s_bv[RS_WARPS][RS_CAP];
  __shared__ int32_t s_bi[RS_WARPS][RS_CAP];
  __shared__ float s_tv[RS_WARPS][RS_KMAX];
  __shared__ int32_t s_ti[RS_WARPS][RS_KMAX];
  __shared__ int32_t s_gl[RS_WARPS][RS_GLIST];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * RS_WARPS + wib;
  if (u >= rp.B) return;
  const int K = rp.K, P = rp.P, ktot = rp.ktot;
  RsRow row{s_bv[wib], s_bi[wib], s_tv[wib], s_ti[wib], 0};
  {
    const uint4* ur = reinterpret_cast<const uint4*>(rp.U_op + (int64_t)u * ktot);
    uint4* dst = reinterpret_cast<uint4*>(s_user[wib]);
    for (int c = lane; c < ktot / 8; c += 32) dst[c] = __ldg(ur + c);
  }
  // T = max over lists of their K-th entry; vtop = the row's best group maximum
  float T = -CUDART_INF_F, vtop = -CUDART_INF_F;
  for (int pp = lane; pp < P; pp += 32) {
    const int64_t o = ((int64_t)pp * rp.B + u) * K;
    T = fmaxf(T, rp.ws_val[o + K - 1]);
    vtop = fmaxf(vtop, rp.ws_val[o]);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    T = fmaxf(T, __shfl_xor_sync(0xffffffffu, T, s));
    vtop = fmaxf(vtop, __shfl_xor_sync(0xffffffffu, vtop, s));
  }
  // tensor-core and mma.sync fp32 sums of the same products differ by O(ktot * 2^-24) relative to sum |products|
  const float margin = (T == -CUDART_INF_F) ? 0.f : 1e-4f * fmaxf(fabsf(T), fabsf(vtop)) + 1e-30f;
  const float T_lo = T - margin;
  const int64_t uid = rp.users ? rp.users[u] : (int64_t)u;
  __syncwarp();
  const uint4* us = reinterpret_cast<const uint4*>(s_user[wib]);
  int32_t* glist = s_gl[wib];
  int cnt = 0;
  const int total = P * K;
  for (int e0 = 0; e0 < total; e0 += 32) {
    const int e = e0 + lane;
    float gv = -CUDART_INF_F;
    int32_t gid = INT32_MAX;
    if (e < total) {
      const int pp = e / K, k = e - pp * K;
      const int64_t o = ((int64_t)pp * rp.B + u) * K + k;
      gv = rp.ws_val[o];
      gid = rp.ws_idx[o];
    }
    const bool sel = gid != INT32_MAX && gv >= T;
    const unsigned sb = __ballot_sync(0xffffffffu, sel);
    if (sel) glist[cnt + __popc(sb & ((1u << lane) - 1u))] = gid;
    cnt += __popc(sb);
    __syncwarp();
    if (cnt > RS_GLIST - 32) {
      rs_process(rp, glist, cnt, us, uid, T_lo, row, lane);
      cnt = 0;
      __syncwarp();
    }
  }
  rs_process(rp, glist, cnt, us, uid, T_lo, row, lane);
  rs_drop_masked(rp, uid, row, lane);
  const int n = rs_compress(row.bv, row.bi, row.tv, row.ti, row.n, K, lane);
  for (int k = lane; k < n; k += 32) {
    rp.out_val[(int64_t)u * K + k] = row.bv[k];
    rp.out_idx[(int64_t)u * K + k] = (int64_t)row.bi[k] + rp.item_offset;
  }
  if (n < K && lane == 0) {
    // fewer than K unmasked items: the reference's index_put_(-1024) + topk returns train items next
    // (PT/Procedure.py:134-135); a shard with fewer than K items pads with (-inf, -1), merged away later
    int got = n;
    if (rp.mask.indptr != nullptr) {
      for (int64_t qq = rp.mask.indptr[uid]; qq < rp.mask.indptr[uid + 1] && got < K; ++qq) {
        const int64_t it = (int64_t)rp.mask.indices[qq] - rp.mask.n_users;
        if (it >= rp.item_offset && it < rp.item_offset + rp.M) {
          rp.out_val[(int64_t)u * K + got] = kMaskValue;
          rp.out_idx[(int64_t)u * K + got] = it;
          ++got;
        }
      }
    }
    for (; got < K; ++got) {
      rp.out_val[(int64_t)u * K + got] = -CUDART_INF_F;
      rp.out_idx[(int64_t)u * K + got] = -1;
    }
  }
}

// --------------------------------------------------------------------------------------------- host
struct GqConfig { int k_blocks, stages, q_cap, union_bound; size_t smem; bool ok; };   // also declared in lgx_score_tc.cu

GqConfig gq_config(int d, int K, int mode) {
  GqConfig c{};
  const int ktot = mode == LGX_SCORE_BF16X3 ? 3 * d : d;
  c.ok = (d % GQ_KBLK == 0) && K >= 1 && K <= 32 && ktot <= RS_KTOT_MAX;
  c.k_blocks = ktot / GQ_KBLK;
  static const int forced_q = [] { const char* e = std::getenv("LGX_SCORE_QCAP"); return e ? std::atoi(e) : 0; }();
  static const int union_env = [] { const char* e = std::getenv("LGX_SCORE_UNION"); return e ? std::atoi(e) : 1; }();
  // Queue depth vs B-ring depth.  The ring matters more: with fewer than 3 stages in flight the TMA latency shows
  // (bf16x3 d = 64, three K blocks per tile: 2.21 ms with 24-row queues + 2 stages, 1.95 ms with 8-row queues + 3
  // stages), so take the deepest queue that still leaves 3 stages, else the configuration with the most stages.
  const int cands[4] = {32, 24, 16, 8};
  bool found = false;
  int best_stages = 0;
  for (int pass = 0; pass < 2 && !found; ++pass) {
    for (int i = 0; i < 4 && !found; ++i) {
      const int qc = cands[i];
      if (forced_q && qc != forced_q) continue;
      size_t fx = 1024 + (size_t)c.k_blocks * GQ_A_BLOCK + GQ_A_BLOCK + GQ_STAGE + (size_t)qc * GQ_EPI * 8 +
                  8 * (2 * GQ_MAX_STAGES + 9) + 24 + (size_t)(4 + 16) * GQ_EPI + 16;
      int uni = union_env ? 1 : 0;
      if (uni && qc == 8 && fx + (size_t)2 * GQ_STAGE > GQ_SMEM_LIMIT) {    // last resort: drop the 4 KB quartile array
        uni = 0;
        fx -= (size_t)16 * GQ_EPI;
      }
      if (fx + (size_t)2 * GQ_STAGE > GQ_SMEM_LIMIT) continue;
      const int stages = (int)std::min<size_t>(GQ_MAX_STAGES, (GQ_SMEM_LIMIT - fx) / GQ_STAGE);
      if (pass == 0) {
        best_stages = std::max(best_stages, stages);
        if (stages < 3 && !forced_q) continue;
      } else if (stages < best_stages) {
        continue;
      }
      c.q_cap = qc;
      c.union_bound = uni;
      c.stages = stages;
      c.smem = fx + (size_t)c.stages * GQ_STAGE;
      found = true;
    }
  }
  if (!found) c.ok = false;
  return c;
}

// per-unit overhead of a (user tile, item split) unit in item tiles of 256, for the wave-aware split planner
// Re-fitted for the round-2 kernel (profiles/r2_score_split_sweep.txt): a tile now costs half of what it did, a unit's
// start-up (cold thresholds: every group is a candidate until the lists fill) does not, so the same overhead is worth
// 45 - 180 tiles; Amazon-Book catalogue, 103 user tiles: 1 split 0.381 ms, 4 splits 0.460 ms; 52 tiles: 2 splits
// 0.263 ms, 5 splits 0.293 ms; 32 tiles: 4 splits 0.210 ms.
constexpr double kGqUnitOverheadTiles = 60.0;
constexpr int kGqCluster = 1;           // default CTAs per cluster (LGX_SCORE_CLUSTER = 1 / 2 / 4)
ScorePlan gq_plan(int B, int M, int sms) {
  // both knobs are re-read on every call (experiments sweep them inside one process)
  const char* es = std::getenv("LGX_SCORE_SPLITS");
  const char* eo = std::getenv("LGX_SCORE_UNIT_OVERHEAD");
  const int forced = es ? std::atoi(es) : 0;
  const double c0 = eo ? std::atof(eo) : kGqUnitOverheadTiles;
  ScorePlan p = plan_score_waves(B, M, GQ_TILE_U, GQ_TILE_I, sms, c0);
  if (forced > 0) {
    const int r = std::max(1, std::min(forced, std::min(p.n_item_tiles, kMaxSplits)));
    p.tiles_per_split = (p.n_item_tiles + r - 1) / r;
    p.n_splits = (p.n_item_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  }
  return p;
}

template <int KMAX, bool SMALLQ, bool SHARE>
static int gq_launch2(dim3 grid, const GqConfig& cfg, const CUtensorMap& tm_u, const CUtensorMap& tm_i,
                      const GqParams& p, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};       // the opt-in is a per-device function attribute
  const int dev = current_device();
  if (dev >= kMaxDevices || !configured[dev]) {
    LGX_CHECK_CUDA(cudaFuncSetAttribute(k_score_topk_gq<KMAX, SMALLQ, SHARE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, GQ_SMEM_LIMIT));
    if (dev < kMaxDevices) configured[dev] = true;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid;
  lc.blockDim = dim3(GQ_THREADS);
  lc.dynamicSmemBytes = cfg.smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = p.cl > 1 ? 1 : 0;
  LGX_CHECK_CUDA(cudaLaunchKernelEx(&lc, k_score_topk_gq<KMAX, SMALLQ, SHARE>, tm_u, tm_i, p));
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

template <int KMAX, bool SMALLQ>
static int gq_launch(dim3 grid, const GqConfig& cfg, const CUtensorMap& tm_u, const CUtensorMap& tm_i,
                     const GqParams& p, cudaStream_t st) {
  return p.row_bound ? gq_launch2<KMAX, SMALLQ, true>(grid, cfg, tm_u, tm_i, p, st)
                     : gq_launch2<KMAX, SMALLQ, false>(grid, cfg, tm_u, tm_i, p, st);
}

int score_topk_gq(const lgx_graph* g, const void* U_op, const int64_t* users, int B, const void* I_op, int M, int d,
                  int K, int mode, int64_t item_offset, int64_t* out_idx, float* out_val, void* workspace,
                  cudaStream_t st) {
  const GqConfig cfg = gq_config(d, K, mode);
  if (!cfg.ok) {
    set_error("tcgen05 scoring needs d % 64 == 0, k <= 32 and 3*d <= 384 (bf16x3) / d <= 256 (bf16); "
              "use LGX_SCORE_FP32 otherwise");
    return LGX_ERR_INVALID;
  }
  const int ktot = cfg.k_blocks * GQ_KBLK;
  CUtensorMap tm_u, tm_i;
  int rc = make_operand_map(&tm_u, U_op, B, ktot, GQ_TILE_U);
  if (rc != LGX_OK) return rc;
  GqParams p;
  {
    const char* e = std::getenv("LGX_GQ_DEBUG");
    p.dbg = e ? std::atoi(e) : 0;
    // CTAs per cluster sharing the B stream (re-read per call: experiments sweep it inside one process)
    const char* ec = std::getenv("LGX_SCORE_CLUSTER");
    const int cl = ec ? std::atoi(ec) : kGqCluster;
    p.cl = (cl == 2 || cl == 4) ? cl : 1;
    const char* er = std::getenv("LGX_SCORE_ROTATE");
    p.rotate = er ? std::atoi(er) : 0;      // measured: 3 % faster without mask / epilogue work, 2 % slower with
  }
  rc = make_operand_map(&tm_i, I_op, M, ktot, (p.cl > 1 && !(p.dbg & 128)) ? GQ_TILE_I / p.cl : GQ_TILE_I);
  if (rc != LGX_OK) return rc;
  const ScorePlan plan = gq_plan(B, M, sm_count());
  p.B = B; p.M = M; p.K = K; p.k_blocks = cfg.k_blocks; p.stages = cfg.stages; p.q_cap = cfg.q_cap;
  p.union_bound = cfg.union_bound;
  p.n_splits = plan.n_splits; p.tiles_per_split = plan.tiles_per_split; p.item_offset = item_offset;
  p.mask = make_mask(g); p.users = users;
  p.has_mask = (g != nullptr && !(p.dbg & 4)) ? 1 : 0;
  p.ws_val = reinterpret_cast<float*>(workspace);
  p.ws_idx = reinterpret_cast<int32_t*>(p.ws_val + (size_t)plan.n_splits * B * K);
  p.row_bound = nullptr;
  if (plan.n_splits > 1) {
    p.row_bound = reinterpret_cast<unsigned*>(p.ws_idx + (size_t)plan.n_splits * B * K);
    LGX_CHECK_CUDA(cudaMemsetAsync(p.row_bound, 0, (size_t)B * sizeof(unsigned), st));
  }
  // ---- train-mask buckets (stream-ordered scratch from the device's memory pool: no sync, reused across calls)
  p.mk_region = nullptr; p.mk_ptr = nullptr; p.mk_entries = nullptr;
  void* mk_scratch = nullptr;
  const int n_tiles = plan.n_item_tiles;
  const size_t bucket_smem = (size_t)(n_tiles + 1) * sizeof(int);
  if (p.has_mask && bucket_smem <= 200 * 1024 && !(p.dbg & 32)) {
    static bool pool_ready[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < kMaxDevices && !pool_ready[dev]) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;                   // keep freed blocks in the pool instead of returning them to the OS
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      pool_ready[dev] = true;
    }
    // entries: every train item of the batch rows once.  Sized for 4x the graph's mean user degree (never more than
    // the whole graph); a user tile that does not fit falls back to walking its lists inside the scoring kernel.
    const double avg = (double)g->nnz * 0.5 / (double)std::max(1, g->n_users);
    const unsigned long long cap = (unsigned long long)std::min<double>((double)g->nnz, 4.0 * avg * B + (1 << 20));
    const size_t off_region = 256;
    const size_t off_ptr = off_region + ((sizeof(int64_t) * plan.n_user_tiles + 255) & ~(size_t)255);
    const size_t off_ent = off_ptr + ((sizeof(int32_t) * (size_t)plan.n_user_tiles * (n_tiles + 1) + 255) & ~(size_t)255);
    const size_t total = off_ent + sizeof(uint16_t) * cap + 256;
    // identity batch (users == NULL): the buckets depend on the graph only and are kept with it
    const char* ecache = std::getenv("LGX_SCORE_MASK_CACHE");          // re-read per call, like the other switches
    const bool cache_on = !ecache || std::atoi(ecache) != 0;
    const bool cacheable = cache_on && users == nullptr;
    bool cached = false;
    unsigned char* base = nullptr;
    static std::mutex mu;                    // held until this call's launches are queued
    std::unique_lock<std::mutex> lock(mu, std::defer_lock);
    if (cacheable) {
      lock.lock();
      if (g->mk_cache && g->mk_key[0] == B && g->mk_key[1] == M && g->mk_key[2] == item_offset) {
        cached = true;
        LGX_CHECK_CUDA(cudaStreamWaitEvent(st, reinterpret_cast<cudaEvent_t>(g->mk_ready), 0));
      } else {
        if (g->mk_cache) { LGX_CHECK_CUDA(cudaDeviceSynchronize()); cudaFree(g->mk_cache); g->mk_cache = nullptr; }
        LGX_CHECK_CUDA(cudaMalloc(&g->mk_cache, total));
        if (!g->mk_ready) {
          cudaEvent_t ev;
          LGX_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
          g->mk_ready = ev;
        }
        g->mk_key[0] = B; g->mk_key[1] = M; g->mk_key[2] = item_offset;
      }
      base = reinterpret_cast<unsigned char*>(g->mk_cache);
    } else {
      LGX_CHECK_CUDA(cudaMallocAsync(&mk_scratch, total, st));
      base = reinterpret_cast<unsigned char*>(mk_scratch);
    }
    if (!cached) LGX_CHECK_CUDA(cudaMemsetAsync(base, 0, 256, st));
    GqBucketParams bp;
    bp.mask = p.mask; bp.mask.m_items_hint = g->m_items; bp.users = users; bp.B = B; bp.M = M; bp.n_tiles = n_tiles; bp.item_offset = item_offset;
    bp.cursor = reinterpret_cast<unsigned long long*>(base);
    bp.cap = cap;
    bp.region = reinterpret_cast<int64_t*>(base + off_region);
    bp.ptr = reinterpret_cast<int32_t*>(base + off_ptr);
    bp.entries = reinterpret_cast<uint16_t*>(base + off_ent);
    if (bucket_smem > 40 * 1024) {
      static size_t configured[kMaxDevices] = {};
      if (dev >= kMaxDevices || bucket_smem > configured[dev]) {
        LGX_CHECK_CUDA(cudaFuncSetAttribute(k_mask_buckets, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev < kMaxDevices) configured[dev] = 200 * 1024;
      }
    }
    if (!cached) {
      k_mask_buckets<<<plan.n_user_tiles, MB_THREADS, bucket_smem, st>>>(bp);
      LGX_CHECK_LAUNCH();
      if (cacheable) LGX_CHECK_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(g->mk_ready), st));
    }
    p.mk_region = bp.region; p.mk_ptr = bp.ptr; p.mk_entries = bp.entries;
  }
  dim3 grid((plan.n_user_tiles + p.cl - 1) / p.cl * p.cl, plan.n_splits);    // whole clusters: padding CTAs are not `live`
  const bool smallq = cfg.q_cap < LGX_GQ_SMALLQ_BELOW;     // deep queues: one check per tile with a 16-append margin
  if (K <= 20 && !smallq) rc = gq_launch<20, false>(grid, cfg, tm_u, tm_i, p, st);
  else if (K <= 20) rc = gq_launch<20, true>(grid, cfg, tm_u, tm_i, p, st);
  else if (K <= 24 && !smallq) rc = gq_launch<24, false>(grid, cfg, tm_u, tm_i, p, st);
  else if (K <= 24) rc = gq_launch<24, true>(grid, cfg, tm_u, tm_i, p, st);
  else if (!smallq) rc = gq_launch<32, false>(grid, cfg, tm_u, tm_i, p, st);
  else rc = gq_launch<32, true>(grid, cfg, tm_u, tm_i, p, st);
  if (mk_scratch) LGX_CHECK_CUDA(cudaFreeAsync(mk_scratch, st));     // stream-ordered: after the scoring kernel
  if (rc != LGX_OK) return rc;
  GqRescoreParams rp;
  rp.ws_val = p.ws_val; rp.ws_idx = p.ws_idx; rp.P = plan.n_splits; rp.B = B; rp.K = K; rp.M = M; rp.ktot = ktot;
  rp.U_op = reinterpret_cast<const __nv_bfloat16*>(U_op); rp.I_op = reinterpret_cast<const __nv_bfloat16*>(I_op);
  rp.mask = p.mask; rp.users = users; rp.item_offset = item_offset; rp.out_idx = out_idx; rp.out_val = out_val;
  k_rescore_topk<<<(B + RS_WARPS - 1) / RS_WARPS, RS_WARPS * 32, 0, st>>>(rp);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

#ifdef LGX_GQ_PROF
extern "C" __attribute__((visibility("default"))) int lgx_debug_gq_cta(long long* out) {
  return cudaMemcpyFromSymbol(out, gq_cta, sizeof(long long) * 1024 * 8) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int lgx_debug_gq_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, gq_trace, sizeof(long long) * 12 * 64 * 6) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int lgx_debug_gq_prof(unsigned long long* out, int reset) {
  unsigned long long z[32] = {};
  if (cudaMemcpyFromSymbol(out, gq_prof, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(gq_prof, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
#endif

}  // namespace lgx
