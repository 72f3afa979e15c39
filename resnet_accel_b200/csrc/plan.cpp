// Host-side schedule builder (see plan.h).  Pure C++; no CUDA calls here.
#include "plan.h"

#include <algorithm>
#include <string>

namespace accel {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int32_t tmem_cols_for(int32_t rows) {
  (void)rows;
  return 256;  // accumulators + two activation stages (bsr_tc.cuh)
}

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

  p->nbr = nbr;
  p->nbc = nbc;
  p->nnz = nnz;
  p->n_chunks = (nbc + kChunkTiles - 1) / kChunkTiles;
  int32_t want = p->group_rows > 0 ? std::min(p->group_rows, kMaxGroupRows) : kDefaultGroupRows;
  const int32_t n_groups = nbr > 0 ? (nbr + want - 1) / want : 0;
  const int32_t rows_per = n_groups ? (nbr + n_groups - 1) / n_groups : 0;  // balanced split
  p->group_rows = rows_per;

  p->groups.clear(); p->batches.clear(); p->op_meta.clear(); p->op_src.clear();
  p->op_blob_off.clear(); p->op_meta_off.clear();

  size_t blob_bytes = 0;
  std::vector<int32_t> cursor(kMaxGroupRows);
  for (int32_t gi = 0; gi < n_groups; ++gi) {
    GroupInfo G{};
    G.br0 = gi * rows_per;
    G.n_rows = std::min(rows_per, nbr - G.br0);
    G.tmem_cols = tmem_cols_for(G.n_rows);
    G.batch_begin = static_cast<int32_t>(p->batches.size());
    for (int g = 0; g < G.n_rows; ++g) {
      cursor[g] = row_ptr[G.br0 + g];
      if (row_ptr[G.br0 + g + 1] > row_ptr[G.br0 + g]) G.nonempty |= (1u << g);
    }
    for (int32_t kc = 0; kc < p->n_chunks; ++kc) {
      const int32_t t_lo = kc * kChunkTiles, t_hi = std::min(nbc, t_lo + kChunkTiles);
      const size_t first_batch_of_chunk = p->batches.size();
      int ops_in_batch = 0;
      auto open_batch = [&]() {
        BatchInfo b{};
        b.blob_off16 = static_cast<uint32_t>(blob_bytes / 16);
        b.chunk = static_cast<uint16_t>(kc);
        b.n_ops = 0;
        b.flags = 0;
        p->batches.push_back(b);
        ops_in_batch = 0;
      };
      auto close_batch = [&]() {
        BatchInfo& b = p->batches.back();
        b.n_ops = static_cast<uint8_t>(ops_in_batch);
        // meta u16s sit right behind the tiles of the batch
        const size_t tiles = static_cast<size_t>(ops_in_batch) * kBTileBytes;
        const size_t base = static_cast<size_t>(b.blob_off16) * 16;
        for (int i = 0; i < ops_in_batch; ++i)
          p->op_meta_off[p->op_meta_off.size() - ops_in_batch + i] = static_cast<uint32_t>(base + tiles + 2 * i);
        blob_bytes = base + tiles + kBatchMetaBytes;
      };
      // Per block-row op lists of this chunk, then emitted round-robin across block-rows: consecutive
      // tcgen05.mma instructions then target different TMEM accumulators, so the small N=16 MMAs do not
      // serialise on the accumulate dependency of a single block-row.
      struct PendingOp { OpSrc src; int32_t win; };
      std::vector<std::vector<PendingOp>> per_row(G.n_rows);
      for (int g = 0; g < G.n_rows; ++g) {
        const int32_t end = row_ptr[G.br0 + g + 1];
        int32_t j = cursor[g];
        while (j < end && col_idx[j] < t_hi) {
          const int32_t t = col_idx[j] - t_lo;  // tile inside the chunk
          PendingOp op{{-1, -1}, 0};
          if (j + 1 < end && col_idx[j + 1] == col_idx[j] + 1 && col_idx[j + 1] < t_hi) {
            op.src.blk_lo = j; op.src.blk_hi = j + 1; op.win = t; j += 2;   // adjacent pair: one MMA
          } else if (t + 1 < kChunkTiles) {
            op.src.blk_lo = j; op.win = t; j += 1;                          // lone block in K slot 0
          } else {
            op.src.blk_hi = j; op.win = t - 1; j += 1;                      // last tile of the chunk: K slot 1
          }
          per_row[g].push_back(op);
        }
        cursor[g] = j;
      }
      bool any = false;
      for (size_t depth = 0;; ++depth) {
        bool found = false;
        for (int g = 0; g < G.n_rows; ++g) {
          if (depth >= per_row[g].size()) continue;
          found = true;
          if (!any || ops_in_batch == kOpsPerBatch) {
            if (any) close_batch();
            open_batch();
            any = true;
          }
          const PendingOp& op = per_row[g][depth];
          const BatchInfo& b = p->batches.back();
          p->op_src.push_back(op.src);
          p->op_meta.push_back(static_cast<uint16_t>((g & 15) | (op.win << 4)));
          p->op_blob_off.push_back(static_cast<uint32_t>(static_cast<size_t>(b.blob_off16) * 16 +
                                                         static_cast<size_t>(ops_in_batch) * kBTileBytes));
          p->op_meta_off.push_back(0);
          ++ops_in_batch;
        }
        if (!found) break;
      }
      if (any) {
        close_batch();
        p->batches[first_batch_of_chunk].flags |= 1;
        p->batches.back().flags |= 2;
        ++G.n_steps;
      }
    }
    G.batch_end = static_cast<int32_t>(p->batches.size());
    p->groups.push_back(G);
  }
  p->n_ops = static_cast<int64_t>(p->op_src.size());

  // ---- device workspace layout
  size_t off = 0;
  p->off_blob = off;     off = align_up(off + std::max<size_t>(blob_bytes, 16), 256);
  p->off_batches = off;  off = align_up(off + std::max<size_t>(p->batches.size(), 1) * sizeof(BatchInfo), 256);
  p->off_groups = off;   off = align_up(off + std::max<size_t>(p->groups.size(), 1) * sizeof(GroupInfo), 256);
  p->off_opsrc = off;    off = align_up(off + std::max<size_t>(p->op_src.size(), 1) * sizeof(OpSrc), 256);
  p->off_opoff = off;    off = align_up(off + std::max<size_t>(p->op_blob_off.size(), 1) * sizeof(uint32_t), 256);
  p->off_opmoff = off;   off = align_up(off + std::max<size_t>(p->op_meta_off.size(), 1) * sizeof(uint32_t), 256);
  p->off_opmeta = off;   off = align_up(off + std::max<size_t>(p->op_meta.size(), 1) * sizeof(uint16_t), 256);
  p->ws_bytes = off;
  return std::string();
}

}  // namespace accel
