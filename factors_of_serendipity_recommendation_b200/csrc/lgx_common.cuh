// Shared helpers for liblgx: error slot, CUDA checks, the graph handle layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/lgx.h"

namespace lgx {

void set_error(const std::string& msg);
int device_ok();   // LGX_OK iff current device is sm_100; sets the error otherwise
int sm_count();          // of the current device
int current_device();
constexpr int kMaxDevices = 64;

#define LGX_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::lgx::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                       __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
      return LGX_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define LGX_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      ::lgx::set_error(std::string("invalid argument: ") + msg); \
      return LGX_ERR_INVALID;                                   \
    }                                                           \
  } while (0)

#define LGX_CHECK_DEVICE()               \
  do {                                   \
    int _d = ::lgx::device_ok();         \
    if (_d != LGX_OK) return _d;         \
  } while (0)

#define LGX_CHECK_LAUNCH() LGX_CHECK_CUDA(cudaGetLastError())

// One schedulable unit of SpMM work: a run of <= chunk_nnz non-zeros of one row.
// Rows longer than chunk_nnz come first in the degree-descending schedule, so the units of split
// rows are exactly units [0, n_partials) and a split unit's partial slot is its own index.
struct WorkItem {
  int64_t start;    // offset into indices / values
  int32_t row;      // output row
  int32_t len;      // number of non-zeros in this unit
};
static_assert(sizeof(WorkItem) == 16, "WorkItem is read as one int4");

// Edge dropout (LightGCN.__dropout_x, PT/model.py:125-143): entry k of A_hat is kept iff
// hash(seed, pos(k)) < keep_prob and then scaled by 1/keep_prob.  pos(k) = k for the forward pass and
// the mirrored entry's position for the backward pass (the dropped graph is not symmetric any more).
struct DropSpec {
  const int32_t* tpos;   // NULL: pos(k) = k
  uint64_t seed;
  float keep_prob;
  float inv_keep;
  int enabled;
};
__host__ __device__ __forceinline__ uint64_t lgx_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ bool lgx_keep(uint64_t seed, int64_t pos, float keep_prob) {
  const uint32_t r = (uint32_t)(lgx_mix64(seed ^ lgx_mix64((uint64_t)pos)) >> 40);   // 24 random bits
  return (float)r * (1.0f / 16777216.0f) < keep_prob;
}

// A row that was split into several units; reduced in fixed order by the long-row kernel.
struct LongRow {
  int32_t row;
  int32_t first_partial;
  int32_t n_partials;
  int32_t pad;
};

}  // namespace lgx

struct lgx_graph {
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  int32_t n_users = 0, m_items = 0;
  int32_t chunk_nnz = 0;
  bool values_are_dinv_products = false;  // built by lgx_graph_build from unique pairs: value == dinv[r]*dinv[c]
  int64_t n_work = 0, n_long = 0, n_partials = 0, max_row_nnz = 0;
  int64_t* indptr = nullptr;     // [n_rows + 1]
  int32_t* indices = nullptr;    // [nnz]
  float* values = nullptr;       // [nnz]
  int32_t* degree = nullptr;     // [n_rows] sum of multiplicities (0 for from_csr graphs: stored nnz)
  float* dinv = nullptr;         // [n_rows]
  int32_t* row_order = nullptr;  // [n_rows] stable degree-descending
  lgx::WorkItem* work = nullptr; // [n_work]
  lgx::LongRow* long_rows = nullptr;  // [n_long]
  int32_t* tpos = nullptr;            // [nnz] position of the mirrored entry (built on demand for dropout)
  // Hot columns (the embedding rows gathered most often), for the shared-memory staging of the SpMM:
  //   hot_ids   [n_hot] column ids by descending gather count (built with the graph; NULL = not built: huge graphs)
  //   hot_count host prefix sums: hot_cover[j] = gathers that hit the top kHotSteps[j] columns
  //   hot_idx / hot_val / hot_work: every work unit's entries re-ordered hot-first for a table of the hot_h hottest
  //             columns -- [n_hot table slots][cold column ids], values alongside, work items with len | n_hot << 16
  //             (built on first use for the table size that fits the embedding width; rebuilt if another width asks
  //             for another size)
  // degree above which a column counts as hot for the SpMM's L2 residency hints, per embedding width class
  // (d <= 32, 64, 128, 256, 512): the rows that qualify fill about 72 MB of L2
  int32_t hot_deg[5] = {0, 0, 0, 0, 0};
  int32_t* hot_ids = nullptr;
  int32_t n_hot = 0;
  int64_t hot_cover[6] = {0, 0, 0, 0, 0, 0};
  mutable int32_t* hot_idx = nullptr;
  mutable float* hot_val = nullptr;
  mutable lgx::WorkItem* hot_work = nullptr;
  mutable int32_t hot_h = 0;
  // Train mask of the identity batch (users == NULL: batch row u is user u) bucketed by (user tile, item tile) for
  // the tcgen05 scoring kernel -- like the reference's allPos (PT/dataloader.py builds it once per dataset), it
  // depends on the interactions only, so it is built by the first scoring call that needs it and kept.
  mutable void* mk_cache = nullptr;
  mutable int64_t mk_key[3] = {-1, -1, -1};      // B, M, item_offset it was built for
  mutable void* mk_ready = nullptr;              // cudaEvent_t recorded behind the build
};
namespace lgx {
constexpr int kHotMax = 3072;                                           // 192 KB of fp32 rows at d = 16
constexpr int kHotSteps[6] = {96, 192, 384, 768, 1536, 3072};           // table sizes for d = 512 ... 16
int build_hot_columns(lgx_graph* g, cudaStream_t st);                   // lgx_graph.cu
int ensure_hot_index(const lgx_graph* g, int h, cudaStream_t st);       // lgx_graph.cu; builds g->hot_idx for table size h
}
