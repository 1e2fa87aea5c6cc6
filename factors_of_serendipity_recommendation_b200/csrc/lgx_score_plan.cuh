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

}  // namespace lgx
