// Host-side schedule builder (see plan.h).  Pure C++; no CUDA calls here.
#include "plan.h"

#include <algorithm>
#include <string>

namespace accel {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

namespace {
struct PendingTile {
  int32_t g, win;
  TileSrc src;
};

// One attempt with a fixed number of block-rows per group.  Returns false when a single group does not fit
// the per-launch parameter tables (the caller retries with fewer rows per group).
bool schedule(const int32_t* row_ptr, const int32_t* col_idx, int32_t nbr, int32_t nbc, int32_t rows_per, Plan* p) {
  p->groups.clear(); p->batches.clear(); p->ops.clear(); p->tile_src.clear(); p->tile_dbg.clear();
  const int32_t n_groups = nbr > 0 ? (nbr + rows_per - 1) / rows_per : 0;
  std::vector<int32_t> cursor(kMaxGroupRows);
  std::vector<PendingTile> tiles;
  for (int32_t gi = 0; gi < n_groups; ++gi) {
    GroupRec G{};
    const int32_t br0 = gi * rows_per;
    const int32_t n_rows = std::min(rows_per, nbr - br0);
    G.br0_rows = static_cast<uint32_t>(br0) | (static_cast<uint32_t>(n_rows) << 16);
    G.batch_begin = static_cast<uint32_t>(p->batches.size());
    G.op_begin = static_cast<uint32_t>(p->ops.size());
    G.blob_off16 = static_cast<uint32_t>(p->tile_src.size() * (kBTileBytes / 16));
    for (int g = 0; g < n_rows; ++g) cursor[g] = row_ptr[br0 + g];
    for (int32_t kc = 0; kc < p->n_chunks; ++kc) {
      const int32_t t_lo = kc * kChunkTiles, t_hi = std::min(nbc, t_lo + kChunkTiles);
      tiles.clear();
      for (int g = 0; g < n_rows; ++g) {
        const int32_t end = row_ptr[br0 + g + 1];
        int32_t j = cursor[g];
        while (j < end && col_idx[j] < t_hi) {
          const int32_t t = col_idx[j] - t_lo;  // tile inside the chunk
          PendingTile pt{g, 0, {-1, -1}};
          if (j + 1 < end && col_idx[j + 1] == col_idx[j] + 1 && col_idx[j + 1] < t_hi) {
            pt.src.blk_lo = j; pt.src.blk_hi = j + 1; pt.win = t; j += 2;   // adjacent pair: one K=32 slice
          } else if (t + 1 < kChunkTiles) {
            pt.src.blk_lo = j; pt.win = t; j += 1;                          // lone block in K slot 0
          } else {
            pt.src.blk_hi = j; pt.win = t - 1; j += 1;                      // last tile of the chunk: K slot 1
          }
          tiles.push_back(pt);
        }
        cursor[g] = j;
      }
      if (tiles.empty()) continue;
      // (window, block-row) order: adjacent block-rows that share a window become one wide MMA
      std::stable_sort(tiles.begin(), tiles.end(), [](const PendingTile& a, const PendingTile& b) {
        return a.win != b.win ? a.win < b.win : a.g < b.g;
      });
      const size_t first_batch = p->batches.size();
      for (size_t i0 = 0; i0 < tiles.size(); i0 += kTilesPerBatch) {
        const size_t n = std::min<size_t>(kTilesPerBatch, tiles.size() - i0);
        uint32_t runs = 0;
        for (size_t i = 0; i < n;) {
          size_t len = 1;
          while (i + len < n && tiles[i0 + i + len].win == tiles[i0 + i].win &&
                 tiles[i0 + i + len].g == tiles[i0 + i].g + static_cast<int32_t>(len))
            ++len;
          OpRec op;
          op.d_n = static_cast<uint32_t>(tiles[i0 + i].g * kTile) | (static_cast<uint32_t>(2 * len) << 17);
          op.a_b = static_cast<uint32_t>(tiles[i0 + i].win * 4) | (static_cast<uint32_t>(i * (kBTileBytes / 16)) << 16);
          p->ops.push_back(op);
          ++runs;
          i += len;
        }
        for (size_t i = 0; i < n; ++i) {
          p->tile_src.push_back(tiles[i0 + i].src);
          p->tile_dbg.push_back(gi); p->tile_dbg.push_back(tiles[i0 + i].g); p->tile_dbg.push_back(kc);
          p->tile_dbg.push_back(tiles[i0 + i].win);
        }
        p->batches.push_back(runs | (static_cast<uint32_t>(n - 1) << 8) | (static_cast<uint32_t>(kc) << 16));
      }
      p->batches[first_batch] |= kBatchFirst;
      p->batches.back() |= kBatchLast;
    }
    G.batch_end = static_cast<uint32_t>(p->batches.size());
    if (G.batch_end - G.batch_begin > static_cast<uint32_t>(kMaxBatchesL) ||
        p->ops.size() - G.op_begin > static_cast<size_t>(kMaxOpsL))
      return false;
    p->groups.push_back(G);
  }
  return true;
}
}  // namespace

std::string build_plan(const int32_t* row_ptr, const int32_t* col_idx, int32_t nbr, int32_t nbc, Plan* p) {
  // ---- structure validation: the rules of validate_bsr (hw/sim/cpp/include/bsr_packer.hpp:364-436)
  if (nbr < 0 || nbc < 0) return "negative block grid";
  if (row_ptr[0] != 0) return "row_ptr[0] should be 0, got " + std::to_string(row_ptr[0]);
  for (int32_t i = 1; i <= nbr; ++i)
    if (row_ptr[i] < row_ptr[i - 1]) return "row_ptr not monotonically increasing at index " + std::to_string(i);
  const int64_t nnz = row_ptr[nbr];
  for (int32_t br = 0; br < nbr; ++br) {
    int32_t prev = -1;
    for (int32_t j = row_ptr[br]; j < row_ptr[br + 1]; ++j) {
      const int32_t c = col_idx[j];
      if (c < 0 || c >= nbc)
        return "col_idx[" + std::to_string(j) + "] = " + std::to_string(c) + " exceeds num_block_cols = " +
               std::to_string(nbc);
      if (c <= prev) return "col_idx not sorted in row " + std::to_string(br) + " at index " + std::to_string(j);
      prev = c;
    }
  }
  if (nbc > 65535 * kChunkTiles) return "too many K tiles for the tensor-core schedule";

  p->nbr = nbr;
  p->nbc = nbc;
  p->nnz = nnz;
  p->n_chunks = (nbc + kChunkTiles - 1) / kChunkTiles;
  int32_t want = p->group_rows > 0 ? std::min(p->group_rows, kMaxGroupRows) : kMaxGroupRows;
  for (;; --want) {
    if (want < 1) return "matrix too large for the tensor-core schedule (one block-row exceeds the launch tables)";
    const int32_t n_groups = nbr > 0 ? (nbr + want - 1) / want : 0;
    const int32_t rows_per = n_groups ? (nbr + n_groups - 1) / n_groups : 0;  // balanced split
    p->group_rows = rows_per;
    if (n_groups == 0 || schedule(row_ptr, col_idx, nbr, nbc, rows_per, p)) break;
  }
  p->n_tiles = static_cast<int64_t>(p->tile_src.size());

  // ---- device workspace layout: the B-tile blob and the repack source list
  size_t off = 0;
  p->off_blob = off;     off = align_up(off + std::max<size_t>(p->tile_src.size() * kBTileBytes, 16), 256);
  p->off_tilesrc = off;  off = align_up(off + std::max<size_t>(p->tile_src.size(), 1) * sizeof(TileSrc), 256);
  p->ws_bytes = off;
  return std::string();
}

}  // namespace accel
