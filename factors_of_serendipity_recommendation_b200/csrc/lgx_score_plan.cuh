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

// Pick the number of item splits so that units = n_user_tiles * n_splits fill the SMs
// (>= units_per_sm * sms units when the batch is small; otherwise the split count in 1..8 with the
// least last-wave waste).
inline ScorePlan plan_score(int B, int M, int tile_users, int tile_items, int sms, int units_per_sm) {
  ScorePlan p;
  p.n_user_tiles = (B + tile_users - 1) / tile_users;
  p.n_item_tiles = (M + tile_items - 1) / tile_items;
  const long target = (long)sms * units_per_sm;
  int splits;
  if (p.n_user_tiles >= target) {
    splits = 1;
    double best = 1e30;
    for (int s = 1; s <= 8; ++s) {
      const long units = (long)p.n_user_tiles * s;
      const double waste = (double)((units + sms - 1) / sms * sms) / (double)units;
      if (waste < best - 0.02) { best = waste; splits = s; }
    }
  } else {
    splits = (int)((target + p.n_user_tiles - 1) / p.n_user_tiles);
  }
  splits = std::max(1, std::min(splits, std::min(p.n_item_tiles, kMaxSplits)));
  p.tiles_per_split = (p.n_item_tiles + splits - 1) / splits;
  p.n_splits = (p.n_item_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

}  // namespace lgx
