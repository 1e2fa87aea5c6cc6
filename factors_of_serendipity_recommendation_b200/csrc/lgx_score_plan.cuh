// Host-side decomposition of a scoring call into (user tile, item split) units.
#pragma once
#include <algorithm>

namespace lgx {

constexpr int kMaxSplits = 160;

struct ScorePlan {
  int n_user_tiles;
  int n_item_tiles;
  int n_splits;         // item-catalogue splits per user tile (each yields a partial top-K list)
  int tiles_per_split;  // item tiles per split
};

// Pick the number of item splits so that units = n_user_tiles * n_splits fill the SMs when the
// batch alone does not.  Splits are not free: every split restarts the running top-K threshold,
// and the number of list inserts per row grows like K*ln(items per split / K) PER SPLIT (5 splits of
// the Amazon catalogue cost 3.9x the inserts of one pass -- measured, profiles/), so a batch with at
// least one user tile per SM is never split.
inline ScorePlan plan_score(int B, int M, int tile_users, int tile_items, int sms, int units_per_sm) {
  ScorePlan p;
  p.n_user_tiles = (B + tile_users - 1) / tile_users;
  p.n_item_tiles = (M + tile_items - 1) / tile_items;
  const long target = (long)sms * units_per_sm;
  int splits;
  if (p.n_user_tiles >= target) {
    splits = 1;
  } else {
    splits = (int)((target + p.n_user_tiles - 1) / p.n_user_tiles);
  }
  splits = std::max(1, std::min(splits, std::min(p.n_item_tiles, kMaxSplits)));
  p.tiles_per_split = (p.n_item_tiles + splits - 1) / splits;
  p.n_splits = (p.n_item_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

// Wave-aware variant for the tcgen05 kernel, whose CTAs share every row's running bound through global memory
// (TcParams::row_bound): a split no longer restarts the threshold from scratch, so splits are cheap enough to be
// used for making units / SMs nearly integral.  Cost model (tile units): waves(R) * (item_tiles / R + c0), with
// c0 = the measured per-unit overhead (A tile reload, pipeline fill, list staging, the unit's early inserts while
// its thresholds are cold) in item tiles, passed in by the caller: 14 for the round-1 kernel; 60 for the group-queue
// kernel, whose tiles cost half as much while a unit's start-up does not (lgx_score_gq.cu, kGqUnitOverheadTiles;
// profiles/r2_score_split_sweep.txt).  The Amazon-Book pass (412 user tiles x 358 item tiles) stays unsplit either way.
inline ScorePlan plan_score_waves(int B, int M, int tile_users, int tile_items, int sms, double unit_overhead_tiles) {
  ScorePlan p;
  p.n_user_tiles = (B + tile_users - 1) / tile_users;
  p.n_item_tiles = (M + tile_items - 1) / tile_items;
  const int max_r = std::max(1, std::min(std::min(p.n_item_tiles / 8, kMaxSplits), 64));
  int best_r = 1;
  double best_cost = 1e300;
  for (int r = 1; r <= max_r; ++r) {
    const long units = (long)r * p.n_user_tiles;
    const long waves = (units + sms - 1) / sms;
    const double cost = (double)waves * ((double)p.n_item_tiles / r + unit_overhead_tiles);
    if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best_r = r; }
  }
  p.tiles_per_split = (p.n_item_tiles + best_r - 1) / best_r;
  p.n_splits = (p.n_item_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

}  // namespace lgx
