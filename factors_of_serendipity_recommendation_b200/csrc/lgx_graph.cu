// Device-side graph build: interactions -> canonical CSR of D^-1/2 A D^-1/2 + degree-sorted schedule.
//
// Replaces PT/dataloader.py:288-293 (UserItemNet, degrees), :349-364 (lil/dok assembly, rowsum,
// d^-1/2, D A D, tocsr) and :331-337 (COO conversion).  The reference assembles the adjacency with
// Python-level scipy lil slice assignment; here both directed copies of every interaction are
// sorted as 64-bit (row, col) keys, run-length encoded (duplicates -> multiplicity) and turned into
// CSR with one pass each.  Device-wide sort / scan / run-length primitives come from CUB (CUDA
// toolkit); every per-element kernel is written here.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include <mutex>
#include <vector>

#include "lgx_common.cuh"

namespace lgx {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

int device_ok() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return LGX_ERR_DEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("liblgx is built for sm_100a (B200) only; found sm_" + std::to_string(major) + std::to_string(minor) +
              " -- there is no fallback path");
    return LGX_ERR_DEVICE;
  }
  return LGX_OK;
}

// Per device: a process may drive several GPUs (one per stream / thread), so nothing here is cached process-wide.
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return dev;
}

int sm_count() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 148;
  int v = cached[dev];             // racing threads compute the same value
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    cached[dev] = v;
  }
  return v;
}

// ------------------------------------------------------------------------------------ kernels
__global__ void k_make_keys(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t E,
                            int32_t n_users, int32_t m_items, uint64_t* __restrict__ keys, int* __restrict__ bad) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    int32_t u = users[e], i = items[e];
    if (u < 0 || u >= n_users || i < 0 || i >= m_items) {
      *bad = 1;
      u = 0;
      i = 0;
    }
    uint64_t ur = (uint64_t)(uint32_t)u, ic = (uint64_t)(uint32_t)(n_users + i);
    keys[e] = (ur << 32) | ic;        // user row, item column
    keys[E + e] = (ic << 32) | ur;    // item row, user column
  }
}

// indptr[r] = first position k with row(uniq[k]) >= r   (r in [0, n_rows])
__global__ void k_indptr_from_keys(const uint64_t* __restrict__ uniq, int64_t nnz, int64_t n_rows,
                                   int64_t* __restrict__ indptr) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)(uniq[mid] >> 32) < r) lo = mid + 1; else hi = mid;
  }
  indptr[r] = lo;
}

__global__ void k_split_keys(const uint64_t* __restrict__ uniq, const int32_t* __restrict__ mult, int64_t nnz,
                             int32_t* __restrict__ indices, int32_t* __restrict__ degree, int has_dups) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    uint64_t key = uniq[k];
    indices[k] = (int32_t)(uint32_t)(key & 0xffffffffull);
    if (has_dups) atomicAdd(&degree[(int64_t)(key >> 32)], mult[k]);
  }
}

__global__ void k_degree_from_indptr(const int64_t* __restrict__ indptr, int64_t n_rows, int32_t* __restrict__ degree) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) degree[r] = (int32_t)(indptr[r + 1] - indptr[r]);
}

// dinv = correctly rounded fp32 of deg^-1/2: IEEE double sqrt and divide, then one rounding to fp32
// (bit-identical to the adjacency shipped with the reference); isolated -> 0
// like d_inv[np.isinf(d_inv)] = 0 at PT/dataloader.py:359.
__global__ void k_dinv(const int32_t* __restrict__ degree, int64_t n_rows, float* __restrict__ dinv) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) {
    int32_t d = degree[r];
    dinv[r] = d > 0 ? (float)(1.0 / sqrt((double)d)) : 0.0f;
  }
}

// values[k] = fl(fl(dinv[row] * mult) * dinv[col])   -- d_mat.dot(adj).dot(d_mat), PT/dataloader.py:362-363
__global__ void k_values(const uint64_t* __restrict__ uniq, const int32_t* __restrict__ mult, int64_t nnz,
                         const float* __restrict__ dinv, float* __restrict__ values) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    uint64_t key = uniq[k];
    float dr = dinv[(int64_t)(key >> 32)], dc = dinv[(int64_t)(key & 0xffffffffull)];
    values[k] = __fmul_rn(__fmul_rn(dr, (float)mult[k]), dc);
  }
}

__global__ void k_row_len(const int64_t* __restrict__ indptr, int64_t n_rows, uint32_t* __restrict__ len,
                          int32_t* __restrict__ iota) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) {
    len[r] = (uint32_t)(indptr[r + 1] - indptr[r]);
    iota[r] = (int32_t)r;
  }
}

// per sorted position: how many work units / partial slots / long-row entries the row needs
__global__ void k_unit_counts(const uint32_t* __restrict__ len_sorted, int64_t n_rows, int32_t chunk,
                              int64_t* __restrict__ n_units, int64_t* __restrict__ n_part, int64_t* __restrict__ is_long) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_rows) {
    int64_t len = len_sorted[s];
    int64_t u = len <= chunk ? 1 : (len + chunk - 1) / chunk;
    n_units[s] = u;
    n_part[s] = u > 1 ? u : 0;
    is_long[s] = u > 1 ? 1 : 0;
  }
}

// One warp per sorted row writes that row's units (lanes stride over the units of a long row).
__global__ void k_fill_work(const int32_t* __restrict__ order, const int64_t* __restrict__ indptr, int64_t n_rows,
                            int32_t chunk, const int64_t* __restrict__ unit_off, const int64_t* __restrict__ part_off,
                            const int64_t* __restrict__ long_off, WorkItem* __restrict__ work,
                            LongRow* __restrict__ long_rows) {
  int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= n_rows) return;
  int32_t row = order[s];
  int64_t beg = indptr[row], len = indptr[row + 1] - beg;
  int64_t u0 = unit_off[s];
  if (len <= chunk) {
    if (lane == 0) work[u0] = WorkItem{beg, row, (int32_t)len};
    return;
  }
  int64_t nu = (len + chunk - 1) / chunk, p0 = part_off[s];
  for (int64_t j = lane; j < nu; j += 32) {
    int64_t st = beg + j * chunk;
    int64_t l = (j == nu - 1) ? (len - j * chunk) : chunk;
    work[u0 + j] = WorkItem{st, row, (int32_t)l};   // u0 == p0: split units lead the schedule
  }
  if (lane == 0) long_rows[long_off[s]] = LongRow{row, (int32_t)p0, (int32_t)nu, 0};
}

// tpos[k] = position of entry (col, row) for entry k = (row, col); one warp per row
__global__ void k_transpose_pos(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n_rows,
                                int32_t* __restrict__ tpos, int* __restrict__ bad) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  for (int64_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) {
    const int32_t c = indices[k];
    int64_t lo = indptr[c], hi = indptr[c + 1];
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (indices[mid] < (int32_t)r) lo = mid + 1; else hi = mid; }
    if (lo < indptr[c + 1] && indices[lo] == (int32_t)r) tpos[k] = (int32_t)lo; else { tpos[k] = (int32_t)k; *bad = 1; }
  }
}

__global__ void k_dropout_mask(int64_t nnz, const int32_t* __restrict__ tpos, uint64_t seed, float keep_prob,
                               uint8_t* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
    out[k] = lgx_keep(seed, tpos ? (int64_t)tpos[k] : k, keep_prob) ? 1 : 0;
}

static inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 32;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, cap));
}
static inline int grid_exact(int64_t n, int block) { return (int)std::max<int64_t>(1, (n + block - 1) / block); }

static void free_graph(lgx_graph* g) {
  if (!g) return;
  cudaFree(g->indptr); cudaFree(g->indices); cudaFree(g->values); cudaFree(g->degree);
  cudaFree(g->dinv); cudaFree(g->row_order); cudaFree(g->work); cudaFree(g->long_rows); cudaFree(g->tpos);
  cudaFree(g->hot_ids); cudaFree(g->hot_idx); cudaFree(g->hot_val); cudaFree(g->hot_work);
  cudaFree(g->mk_cache);
  if (g->mk_ready) cudaEventDestroy(reinterpret_cast<cudaEvent_t>(g->mk_ready));
  delete g;
}

// Build row_order + work items for a graph whose indptr is final.
static int build_schedule(lgx_graph* g, int32_t chunk_nnz, cudaStream_t st) {
  const int64_t n = g->n_rows;
  g->chunk_nnz = chunk_nnz > 0 ? chunk_nnz : 256;
  uint32_t *len = nullptr, *len_sorted = nullptr;
  int32_t* iota = nullptr;
  int64_t *n_units = nullptr, *n_part = nullptr, *is_long = nullptr;
  void* tmp = nullptr;
  int rc = LGX_OK;
  auto cleanup = [&]() {
    cudaFree(len); cudaFree(len_sorted); cudaFree(iota); cudaFree(n_units); cudaFree(n_part); cudaFree(is_long);
    cudaFree(tmp);
  };
#define SCHED_CUDA(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                 \
      cleanup();                                                                            \
      return LGX_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)
  SCHED_CUDA(cudaMalloc(&len, sizeof(uint32_t) * (n + 1)));
  SCHED_CUDA(cudaMalloc(&len_sorted, sizeof(uint32_t) * (n + 1)));
  SCHED_CUDA(cudaMalloc(&iota, sizeof(int32_t) * (n + 1)));
  SCHED_CUDA(cudaMalloc(&g->row_order, sizeof(int32_t) * (n + 1)));
  SCHED_CUDA(cudaMalloc(&n_units, sizeof(int64_t) * (n + 1)));
  SCHED_CUDA(cudaMalloc(&n_part, sizeof(int64_t) * (n + 1)));
  SCHED_CUDA(cudaMalloc(&is_long, sizeof(int64_t) * (n + 1)));
  k_row_len<<<grid_exact(n, 256), 256, 0, st>>>(g->indptr, n, len, iota);
  // stable descending radix sort: equal-length rows keep ascending row id
  size_t tb = 0, tb2 = 0;
  SCHED_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, len, len_sorted, iota, g->row_order, n, 0, 32, st));
  SCHED_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, n_units, n_units, n + 1, st));
  tb = std::max(tb, tb2);
  SCHED_CUDA(cudaMalloc(&tmp, tb + 16));
  SCHED_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, tb, len, len_sorted, iota, g->row_order, n, 0, 32, st));
  SCHED_CUDA(cudaMemsetAsync(n_units + n, 0, sizeof(int64_t), st));
  SCHED_CUDA(cudaMemsetAsync(n_part + n, 0, sizeof(int64_t), st));
  SCHED_CUDA(cudaMemsetAsync(is_long + n, 0, sizeof(int64_t), st));
  k_unit_counts<<<grid_exact(n, 256), 256, 0, st>>>(len_sorted, n, g->chunk_nnz, n_units, n_part, is_long);
  SCHED_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, n_units, n_units, n + 1, st));
  SCHED_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, n_part, n_part, n + 1, st));
  SCHED_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, is_long, is_long, n + 1, st));
  int64_t totals[3] = {0, 0, 0};
  uint32_t max_len = 0;
  SCHED_CUDA(cudaMemcpyAsync(&totals[0], n_units + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SCHED_CUDA(cudaMemcpyAsync(&totals[1], n_part + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SCHED_CUDA(cudaMemcpyAsync(&totals[2], is_long + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (n > 0) SCHED_CUDA(cudaMemcpyAsync(&max_len, len_sorted, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  // hot-degree thresholds for the L2 hints: the row lengths at the ranks where the hottest rows fill ~72 MB.  A row
  // shard (n_rows < n_cols, degree-cyclic deal) sees every n_cols / n_rows-th row of the global order.
  uint32_t hot_len[5] = {0, 0, 0, 0, 0};
  if (n > 0) {
    const int widths[5] = {32, 64, 128, 256, 512};
    for (int j = 0; j < 5; ++j) {
      static const double budget = [] { const char* e = std::getenv("LGX_SPMM_L2_MB"); return (e ? std::atof(e) : 72.0) * 1e6; }();
      const double h_global = budget / (widths[j] * 4.0);
      const int64_t rank = std::min<int64_t>(n - 1, (int64_t)(h_global * (double)n / (double)std::max<int64_t>(1, g->n_cols)));
      SCHED_CUDA(cudaMemcpyAsync(&hot_len[j], len_sorted + rank, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
  }
  SCHED_CUDA(cudaStreamSynchronize(st));
  for (int j = 0; j < 5; ++j) g->hot_deg[j] = (int32_t)hot_len[j];
  g->n_work = totals[0];
  g->n_partials = totals[1];
  g->n_long = totals[2];
  g->max_row_nnz = max_len;
  SCHED_CUDA(cudaMalloc(&g->work, sizeof(WorkItem) * std::max<int64_t>(1, g->n_work)));
  SCHED_CUDA(cudaMalloc(&g->long_rows, sizeof(LongRow) * std::max<int64_t>(1, g->n_long)));
  if (n > 0) {
    k_fill_work<<<grid_exact(n * 32, 256), 256, 0, st>>>(g->row_order, g->indptr, n, g->chunk_nnz, n_units, n_part,
                                                         is_long, g->work, g->long_rows);
    SCHED_CUDA(cudaGetLastError());
  }
  SCHED_CUDA(cudaStreamSynchronize(st));
#undef SCHED_CUDA
  cleanup();
  return rc;
}

// ---------------------------------------------------------------------------------- hot columns
// The SpMM gathers X[col] once per stored entry, so "hot" = most frequent column.  For the square symmetric A_hat
// that is the row degree, but a row shard (n_rows < n_cols) has its own column statistics: count them.
__global__ void k_col_hist(const int32_t* __restrict__ indices, int64_t nnz, uint32_t* __restrict__ cnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) atomicAdd(&cnt[indices[k]], 1u);
}
__global__ void k_iota(int32_t* __restrict__ a, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}
__global__ void k_hot_rank(const int32_t* __restrict__ hot_ids, int h, int32_t* __restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < h) rank[hot_ids[i]] = i;
}
// One warp per work unit: entries whose column has a table slot go first (as the slot), the rest keep their column id;
// both parts keep their original relative order.
__global__ void __launch_bounds__(256)
k_hot_reorder(const WorkItem* __restrict__ work, int64_t n_work, const int32_t* __restrict__ indices,
              const float* __restrict__ values, const int32_t* __restrict__ rank, int32_t* __restrict__ hot_idx,
              float* __restrict__ hot_val, WorkItem* __restrict__ hot_work) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_work) return;
  const WorkItem it = work[w];
  const unsigned lt = (1u << lane) - 1u;
  int n_hot = 0;
  for (int base = 0; base < it.len; base += 32) {
    const bool hot = base + lane < it.len && rank[indices[it.start + base + lane]] >= 0;
    n_hot += __popc(__ballot_sync(0xffffffffu, hot));
  }
  int ph = 0, pc = n_hot;
  for (int base = 0; base < it.len; base += 32) {
    const bool in = base + lane < it.len;
    int32_t c = 0, r = -1;
    float v = 0.f;
    if (in) { c = indices[it.start + base + lane]; v = values[it.start + base + lane]; r = rank[c]; }
    const unsigned hm = __ballot_sync(0xffffffffu, in && r >= 0), cm = __ballot_sync(0xffffffffu, in && r < 0);
    if (in) {
      const int pos = r >= 0 ? ph + __popc(hm & lt) : pc + __popc(cm & lt);
      hot_idx[it.start + pos] = r >= 0 ? r : c;
      hot_val[it.start + pos] = v;
    }
    ph += __popc(hm);
    pc += __popc(cm);
  }
  if (lane == 0) {
    WorkItem o = it;
    o.len = it.len | (n_hot << 16);
    hot_work[w] = o;
  }
}

int build_hot_columns(lgx_graph* g, cudaStream_t st) {
  // opt-in (see hot_table_rows in lgx_spmm.cu); not for graphs where the re-ordered copies would cost GBs
  static const int enabled = [] { const char* e = std::getenv("LGX_SPMM_HOT"); return e ? std::atoi(e) : 0; }();
  if (enabled != 1) return LGX_OK;
  if (g->nnz == 0 || g->nnz > ((int64_t)1 << 28) || g->n_cols < 2 * kHotMax) return LGX_OK;
  const int64_t n = g->n_cols;
  uint32_t *cnt = nullptr, *cnt_sorted = nullptr;
  int32_t *iota = nullptr, *ids_sorted = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(cnt); cudaFree(cnt_sorted); cudaFree(iota); cudaFree(ids_sorted); cudaFree(tmp); };
#define HOT_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                 \
      cleanup();                                                                            \
      return LGX_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)
  HOT_CUDA(cudaMalloc(&cnt, sizeof(uint32_t) * n));
  HOT_CUDA(cudaMalloc(&cnt_sorted, sizeof(uint32_t) * n));
  HOT_CUDA(cudaMalloc(&iota, sizeof(int32_t) * n));
  HOT_CUDA(cudaMalloc(&ids_sorted, sizeof(int32_t) * n));
  HOT_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * n, st));
  k_col_hist<<<grid_for(g->nnz, 256), 256, 0, st>>>(g->indices, g->nnz, cnt);
  k_iota<<<grid_exact(n, 256), 256, 0, st>>>(iota, n);
  size_t tb = 0;
  HOT_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, cnt, cnt_sorted, iota, ids_sorted, n, 0, 32, st));
  HOT_CUDA(cudaMalloc(&tmp, tb + 16));
  HOT_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, tb, cnt, cnt_sorted, iota, ids_sorted, n, 0, 32, st));
  const int h = (int)std::min<int64_t>(kHotMax, n);
  HOT_CUDA(cudaMalloc(&g->hot_ids, sizeof(int32_t) * h));
  HOT_CUDA(cudaMemcpyAsync(g->hot_ids, ids_sorted, sizeof(int32_t) * h, cudaMemcpyDeviceToDevice, st));
  std::vector<uint32_t> top(h);
  HOT_CUDA(cudaMemcpyAsync(top.data(), cnt_sorted, sizeof(uint32_t) * h, cudaMemcpyDeviceToHost, st));
  HOT_CUDA(cudaStreamSynchronize(st));
#undef HOT_CUDA
  g->n_hot = h;
  int64_t acc = 0;
  int j = 0;
  for (int i = 0; i < h; ++i) {
    acc += top[i];
    while (j < 6 && i + 1 == kHotSteps[j]) g->hot_cover[j++] = acc;
  }
  for (; j < 6; ++j) g->hot_cover[j] = acc;
  cleanup();
  return LGX_OK;
}

// hot_idx for a table of the h hottest columns.  Serialised by a mutex: the first SpMM of a given width builds it.
int ensure_hot_index(const lgx_graph* g, int h, cudaStream_t st) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (g->hot_idx != nullptr && g->hot_h == h) return LGX_OK;
  if (g->hot_ids == nullptr || h > g->n_hot) { set_error("hot columns were not built for this graph"); return LGX_ERR_INVALID; }
  int32_t* rank = nullptr;
  LGX_CHECK_CUDA(cudaMalloc(&rank, sizeof(int32_t) * g->n_cols));
  if (g->hot_idx == nullptr) {
    cudaError_t e = cudaMalloc(&g->hot_idx, sizeof(int32_t) * g->nnz);
    if (e == cudaSuccess) e = cudaMalloc(&g->hot_val, sizeof(float) * g->nnz);
    if (e == cudaSuccess) e = cudaMalloc(&g->hot_work, sizeof(WorkItem) * std::max<int64_t>(1, g->n_work));
    if (e != cudaSuccess) { cudaFree(rank); set_error("cudaMalloc(hot arrays) failed"); return LGX_ERR_CUDA; }
  }
  cudaMemsetAsync(rank, 0xff, sizeof(int32_t) * g->n_cols, st);
  k_hot_rank<<<grid_exact(h, 256), 256, 0, st>>>(g->hot_ids, h, rank);
  if (g->n_work > 0)
    k_hot_reorder<<<grid_exact(g->n_work * 32, 256), 256, 0, st>>>(g->work, g->n_work, g->indices, g->values, rank,
                                                                  g->hot_idx, g->hot_val, g->hot_work);
  cudaError_t e = cudaStreamSynchronize(st);      // one-time set-up; a width change mid-stream must not race the old index
  cudaFree(rank);
  if (e != cudaSuccess) { set_error(std::string("hot index build failed: ") + cudaGetErrorString(e)); return LGX_ERR_CUDA; }
  g->hot_h = h;
  return LGX_OK;
}

}  // namespace lgx

using namespace lgx;

extern "C" {

const char* lgx_last_error(void) { return g_last_error.c_str(); }
int lgx_version(void) { return 100; }

int lgx_device_check(int* sm, int64_t* l2_bytes) {
  LGX_CHECK_DEVICE();
  int dev = 0;
  cudaGetDevice(&dev);
  if (sm) *sm = sm_count();
  if (l2_bytes) {
    int l2 = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
    *l2_bytes = l2;
  }
  return LGX_OK;
}

int lgx_graph_build(int32_t n_users, int32_t m_items, int64_t n_edges, const int32_t* users, const int32_t* items,
                    int32_t chunk_nnz, lgx_stream stream, lgx_graph** out) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(out != nullptr, "out is NULL");
  LGX_REQUIRE(n_users > 0 && m_items > 0, "n_users and m_items must be positive");
  LGX_REQUIRE(n_edges >= 0 && n_edges <= 1073000000, "n_edges out of range (max 1.073e9 interactions per handle: 2E must fit int32 for the run-length pass)");
  LGX_REQUIRE(n_edges == 0 || (users && items), "users/items NULL");
  LGX_REQUIRE((int64_t)n_users + m_items < ((int64_t)1 << 31), "n_users + m_items must fit int32");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = (int64_t)n_users + m_items, E2 = 2 * n_edges;

  lgx_graph* g = new lgx_graph();
  g->n_rows = g->n_cols = N;
  g->n_users = n_users;
  g->m_items = m_items;
  uint64_t *keys = nullptr, *keys_sorted = nullptr, *uniq = nullptr;
  int32_t* mult = nullptr;
  int64_t* d_runs = nullptr;
  int* d_bad = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(keys); cudaFree(keys_sorted); cudaFree(uniq); cudaFree(mult); cudaFree(d_runs); cudaFree(d_bad);
    cudaFree(tmp);
  };
#define BUILD_CUDA(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                 \
      cleanup();                                                                            \
      free_graph(g);                                                                        \
      return LGX_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)
  const int64_t cap = std::max<int64_t>(1, E2);
  BUILD_CUDA(cudaMalloc(&keys, sizeof(uint64_t) * cap));
  BUILD_CUDA(cudaMalloc(&keys_sorted, sizeof(uint64_t) * cap));
  BUILD_CUDA(cudaMalloc(&d_runs, sizeof(int64_t)));
  BUILD_CUDA(cudaMalloc(&d_bad, sizeof(int)));
  BUILD_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  BUILD_CUDA(cudaMemsetAsync(d_runs, 0, sizeof(int64_t), st));
  int64_t nnz = 0;
  if (n_edges > 0) {
    k_make_keys<<<grid_for(n_edges, 256), 256, 0, st>>>(users, items, n_edges, n_users, m_items, keys, d_bad);
    int bits_n = 1;
    while (((int64_t)1 << bits_n) < N) ++bits_n;
    size_t tb = 0, tb2 = 0;
    BUILD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, keys_sorted, E2, 0, 32 + bits_n, st));
    // run-length encode into `keys` (reused as the unique-key array) + multiplicities
    BUILD_CUDA(cudaMalloc(&mult, sizeof(int32_t) * cap));
    BUILD_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb2, keys_sorted, keys, mult, d_runs, E2, st));
    tb = std::max(tb, tb2);
    BUILD_CUDA(cudaMalloc(&tmp, tb + 16));
    BUILD_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, keys_sorted, E2, 0, 32 + bits_n, st));
    BUILD_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, tb, keys_sorted, keys, mult, d_runs, E2, st));
    int bad = 0;
    BUILD_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    BUILD_CUDA(cudaMemcpyAsync(&nnz, d_runs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BUILD_CUDA(cudaStreamSynchronize(st));
    if (bad) {
      set_error("invalid argument: a user or item id is outside [0, n_users) / [0, m_items)");
      cleanup();
      free_graph(g);
      return LGX_ERR_INVALID;
    }
  }
  uniq = keys;       // alias: unique keys live in `keys` now
  keys = nullptr;
  g->nnz = nnz;
  const int has_dups = nnz != E2;
  g->values_are_dinv_products = !has_dups;
  BUILD_CUDA(cudaMalloc(&g->indptr, sizeof(int64_t) * (N + 1)));
  BUILD_CUDA(cudaMalloc(&g->indices, sizeof(int32_t) * std::max<int64_t>(1, nnz)));
  BUILD_CUDA(cudaMalloc(&g->values, sizeof(float) * std::max<int64_t>(1, nnz)));
  BUILD_CUDA(cudaMalloc(&g->degree, sizeof(int32_t) * N));
  BUILD_CUDA(cudaMalloc(&g->dinv, sizeof(float) * N));
  BUILD_CUDA(cudaMemsetAsync(g->degree, 0, sizeof(int32_t) * N, st));
  k_indptr_from_keys<<<grid_exact(N + 1, 256), 256, 0, st>>>(uniq, nnz, N, g->indptr);
  if (nnz > 0) k_split_keys<<<grid_for(nnz, 256), 256, 0, st>>>(uniq, mult, nnz, g->indices, g->degree, has_dups);
  if (!has_dups) k_degree_from_indptr<<<grid_exact(N, 256), 256, 0, st>>>(g->indptr, N, g->degree);
  k_dinv<<<grid_exact(N, 256), 256, 0, st>>>(g->degree, N, g->dinv);
  if (nnz > 0) k_values<<<grid_for(nnz, 256), 256, 0, st>>>(uniq, mult, nnz, g->dinv, g->values);
  BUILD_CUDA(cudaGetLastError());
  BUILD_CUDA(cudaStreamSynchronize(st));
#undef BUILD_CUDA
  cleanup();
  int rc = build_schedule(g, chunk_nnz, st);
  if (rc == LGX_OK) rc = build_hot_columns(g, st);
  if (rc != LGX_OK) {
    free_graph(g);
    return rc;
  }
  *out = g;
  return LGX_OK;
}

int lgx_graph_build_host(int32_t n_users, int32_t m_items, int64_t n_edges, const int32_t* users_host,
                         const int32_t* items_host, int32_t chunk_nnz, lgx_stream stream, lgx_graph** out) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(n_edges >= 0, "n_edges negative");
  LGX_REQUIRE(n_edges == 0 || (users_host && items_host), "users/items NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t *du = nullptr, *di = nullptr;
  const size_t bytes = sizeof(int32_t) * std::max<int64_t>(1, n_edges);
  LGX_CHECK_CUDA(cudaMalloc(&du, bytes));
  if (cudaMalloc(&di, bytes) != cudaSuccess) {
    cudaFree(du);
    set_error("cudaMalloc failed for the item array");
    return LGX_ERR_CUDA;
  }
  cudaMemcpyAsync(du, users_host, sizeof(int32_t) * n_edges, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(di, items_host, sizeof(int32_t) * n_edges, cudaMemcpyHostToDevice, st);
  int rc = lgx_graph_build(n_users, m_items, n_edges, du, di, chunk_nnz, stream, out);
  cudaFree(du);
  cudaFree(di);
  return rc;
}

int lgx_graph_from_csr(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* indptr, const int32_t* indices,
                       const float* values, int32_t n_users, int32_t m_items, int32_t chunk_nnz, lgx_stream stream,
                       lgx_graph** out) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(out != nullptr, "out is NULL");
  LGX_REQUIRE(n_rows > 0 && n_cols > 0 && nnz >= 0, "bad CSR shape");
  LGX_REQUIRE(n_rows < ((int64_t)1 << 31) && n_cols < ((int64_t)1 << 31), "rows/cols must fit int32");
  LGX_REQUIRE(indptr && (nnz == 0 || (indices && values)), "CSR arrays NULL");
  cudaStream_t st = (cudaStream_t)stream;
  lgx_graph* g = new lgx_graph();
  g->n_rows = n_rows; g->n_cols = n_cols; g->nnz = nnz; g->n_users = n_users; g->m_items = m_items;
#define CSR_CUDA(expr)                                                       \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e)); \
      free_graph(g);                                                         \
      return LGX_ERR_CUDA;                                                   \
    }                                                                        \
  } while (0)
  CSR_CUDA(cudaMalloc(&g->indptr, sizeof(int64_t) * (n_rows + 1)));
  CSR_CUDA(cudaMalloc(&g->indices, sizeof(int32_t) * std::max<int64_t>(1, nnz)));
  CSR_CUDA(cudaMalloc(&g->values, sizeof(float) * std::max<int64_t>(1, nnz)));
  CSR_CUDA(cudaMalloc(&g->degree, sizeof(int32_t) * n_rows));
  CSR_CUDA(cudaMalloc(&g->dinv, sizeof(float) * n_rows));
  CSR_CUDA(cudaMemcpyAsync(g->indptr, indptr, sizeof(int64_t) * (n_rows + 1), cudaMemcpyDeviceToDevice, st));
  if (nnz > 0) {
    CSR_CUDA(cudaMemcpyAsync(g->indices, indices, sizeof(int32_t) * nnz, cudaMemcpyDeviceToDevice, st));
    CSR_CUDA(cudaMemcpyAsync(g->values, values, sizeof(float) * nnz, cudaMemcpyDeviceToDevice, st));
  }
  k_degree_from_indptr<<<grid_exact(n_rows, 256), 256, 0, st>>>(g->indptr, n_rows, g->degree);
  k_dinv<<<grid_exact(n_rows, 256), 256, 0, st>>>(g->degree, n_rows, g->dinv);
  CSR_CUDA(cudaGetLastError());
#undef CSR_CUDA
  int rc = build_schedule(g, chunk_nnz, st);
  if (rc == LGX_OK) rc = build_hot_columns(g, st);
  if (rc != LGX_OK) {
    free_graph(g);
    return rc;
  }
  *out = g;
  return LGX_OK;
}

int lgx_graph_get_flags(const lgx_graph* g) {
  return g && g->values_are_dinv_products ? LGX_GRAPH_NORMALIZED : 0;
}

int lgx_graph_set_flags(lgx_graph* g, int32_t flags) {
  LGX_REQUIRE(g != nullptr, "graph is NULL");
  g->values_are_dinv_products = (flags & LGX_GRAPH_NORMALIZED) != 0;
  return LGX_OK;
}

int lgx_graph_enable_dropout(lgx_graph* g, lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(g, "graph is NULL");
  if (g->tpos) return LGX_OK;
  LGX_REQUIRE(g->n_rows == g->n_cols, "dropout needs the square (symmetric-structure) graph");
  LGX_REQUIRE(g->nnz < ((int64_t)1 << 31), "dropout positions are int32: nnz must be < 2^31");
  cudaStream_t st = (cudaStream_t)stream;
  int* d_bad = nullptr;
  LGX_CHECK_CUDA(cudaMalloc(&g->tpos, sizeof(int32_t) * std::max<int64_t>(1, g->nnz)));
  LGX_CHECK_CUDA(cudaMalloc(&d_bad, sizeof(int)));
  LGX_CHECK_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  k_transpose_pos<<<grid_exact(g->n_rows * 32, 256), 256, 0, st>>>(g->indptr, g->indices, g->n_rows, g->tpos, d_bad);
  int bad = 0;
  LGX_CHECK_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  LGX_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_bad);
  if (bad) {
    cudaFree(g->tpos);
    g->tpos = nullptr;
    set_error("invalid argument: the graph's sparsity pattern is not symmetric (dropout backward needs mirrored entries)");
    return LGX_ERR_INVALID;
  }
  return LGX_OK;
}

int lgx_dropout_mask(const lgx_graph* g, float keep_prob, uint64_t seed, int32_t transpose, uint8_t* mask,
                     lgx_stream stream) {
  LGX_CHECK_DEVICE();
  LGX_REQUIRE(g && mask, "NULL argument");
  LGX_REQUIRE(!transpose || g->tpos, "call lgx_graph_enable_dropout first");
  if (g->nnz > 0)
    k_dropout_mask<<<grid_for(g->nnz, 256), 256, 0, (cudaStream_t)stream>>>(g->nnz, transpose ? g->tpos : nullptr, seed,
                                                                           keep_prob, mask);
  LGX_CHECK_LAUNCH();
  return LGX_OK;
}

int lgx_graph_info(const lgx_graph* g, int64_t* info) {
  LGX_REQUIRE(g && info, "NULL argument");
  info[0] = g->n_rows; info[1] = g->n_cols; info[2] = g->nnz; info[3] = g->n_users; info[4] = g->m_items;
  info[5] = g->n_work; info[6] = g->n_long; info[7] = g->max_row_nnz; info[8] = g->n_partials;
  info[9] = g->chunk_nnz;
  return LGX_OK;
}

int lgx_graph_export(const lgx_graph* g, int64_t* indptr, int32_t* indices, float* values, int32_t* degree,
                     float* dinv, int32_t* row_order, lgx_stream stream) {
  LGX_REQUIRE(g, "graph is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const auto D2D = cudaMemcpyDeviceToDevice;
  if (indptr) LGX_CHECK_CUDA(cudaMemcpyAsync(indptr, g->indptr, sizeof(int64_t) * (g->n_rows + 1), D2D, st));
  if (indices && g->nnz) LGX_CHECK_CUDA(cudaMemcpyAsync(indices, g->indices, sizeof(int32_t) * g->nnz, D2D, st));
  if (values && g->nnz) LGX_CHECK_CUDA(cudaMemcpyAsync(values, g->values, sizeof(float) * g->nnz, D2D, st));
  if (degree) LGX_CHECK_CUDA(cudaMemcpyAsync(degree, g->degree, sizeof(int32_t) * g->n_rows, D2D, st));
  if (dinv) LGX_CHECK_CUDA(cudaMemcpyAsync(dinv, g->dinv, sizeof(float) * g->n_rows, D2D, st));
  if (row_order) LGX_CHECK_CUDA(cudaMemcpyAsync(row_order, g->row_order, sizeof(int32_t) * g->n_rows, D2D, st));
  return LGX_OK;
}

int lgx_graph_pointers(const lgx_graph* g, const int64_t** indptr, const int32_t** indices, const float** values) {
  LGX_REQUIRE(g, "graph is NULL");
  if (indptr) *indptr = g->indptr;
  if (indices) *indices = g->indices;
  if (values) *values = g->values;
  return LGX_OK;
}

int lgx_graph_destroy(lgx_graph* g) {
  free_graph(g);
  return LGX_OK;
}

}  // extern "C"
