// Host-side MMA schedule for one BSR weight matrix (Convention B, 14x14 blocks).
//
// The reference stores a block-row as a list of 14x14 int8 blocks (196 B, row-major) at K tiles
// col_idx[] (sw/training/export_bsr_14x14.py:241-272).  The tensor core wants K = 32 bytes per
// instruction and 16-byte core matrices, so at load time every block is re-laid into a "B tile" of
// 16 rows x 32 K-bytes: two 16-byte K slots, each holding one 14-wide block (+2 zero bytes).  Two
// stored blocks at adjacent K tiles of the same block-row share one tile; a lone block takes half.
//
// Schedule: block-rows are split into groups (one CTA column each).  Inside a group the K tiles are
// walked in chunks of 9 (one activation stage); inside a chunk the tiles are ordered by
// (window, block-row) so that block-rows that are adjacent AND use the same K window sit next to
// each other in memory: such a run of `len` tiles is issued as ONE tcgen05.mma with N = 16*len.
// The per-MMA operands are precomputed here as two 32-bit words (OpRec) that the kernel reads from
// its parameter bank with uniform loads - measured on B200 (tools/probe/mma_probe4..6): operands that
// travel through R2UR cost 60-140 cycles per MMA, operands derived in the uniform datapath ~10.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace accel {

constexpr int kBlock = 14;           // reference block size (export_bsr_14x14.py:48)
constexpr int kTile = 16;            // padded K slot / padded block-row height
constexpr int kChunkTiles = 9;       // K tiles per activation stage: 126 k = 14 channels x 9 taps of a 3x3 conv
constexpr int kMaxGroupRows = 11;    // block-rows per CTA: 11 x 16 accumulator columns + 2 x 36 activation columns <= 256
constexpr int kTilesPerBatch = 32;   // B tiles per weight stage (16 KB)
constexpr int kBTileBytes = 512;     // 16 rows x 32 bytes: [row_group(2)][k_slot(2)][8 rows][16 B]

// kernel-parameter tables (one launch): sized so that the whole parameter block stays below 32 KB
constexpr int kMaxGroupsL = 64;
constexpr int kMaxBatchesL = 1024;
constexpr int kMaxOpsL = 2816;

struct OpRec {           // one tcgen05.mma
  uint32_t d_n;          // bits 0..8: accumulator column (g*16); bits 17..22: N>>3 (already in idesc position)
  uint32_t a_b;          // bits 0..8: activation column inside the stage (window*4); bits 16..31: B offset / 16 inside the weight stage
};

// batch word: [0:8) MMAs, [8:14) tiles-1, bit 14 first batch of its chunk, bit 15 last batch of its chunk, [16:32) chunk
constexpr uint32_t kBatchFirst = 1u << 14;
constexpr uint32_t kBatchLast = 1u << 15;
inline uint32_t batch_runs(uint32_t w) { return w & 0xffu; }
inline uint32_t batch_tiles(uint32_t w) { return ((w >> 8) & 0x3fu) + 1; }
inline uint32_t batch_chunk(uint32_t w) { return w >> 16; }

struct GroupRec {        // 20 bytes, one per CTA column
  uint32_t br0_rows;     // first block-row | n_rows << 16
  uint32_t batch_begin;  // [batch_begin, batch_end) in the batch table
  uint32_t batch_end;
  uint32_t op_begin;     // first OpRec of the group
  uint32_t blob_off16;   // byte offset / 16 of the group's first B tile
};

struct TileSrc {         // repack kernel input: which stored blocks feed B tile `t`
  int32_t blk_lo;        // block placed in K slot 0 (bytes 0..13), -1 = none
  int32_t blk_hi;        // block placed in K slot 1 (bytes 16..29), -1 = none
};

struct Plan {
  int32_t nbr = 0, nbc = 0;
  int64_t nnz = 0;
  int32_t n_chunks = 0;
  int32_t group_rows = 0;
  std::vector<GroupRec> groups;
  std::vector<uint32_t> batches;
  std::vector<OpRec> ops;
  std::vector<TileSrc> tile_src;      // per B tile, in blob order
  // per tile: (group, local row, chunk, window) for tests / tooling
  std::vector<int32_t> tile_dbg;
  // device workspace layout (byte offsets)
  size_t off_blob = 0, off_tilesrc = 0;
  size_t ws_bytes = 0;
  int64_t n_tiles = 0;
  // filled by upload
  const uint8_t* ws_dev = nullptr;
  bool uploaded = false;
};

// Returns empty string on success, else a validate_bsr-style message (bsr_packer.hpp:364-436).
std::string build_plan(const int32_t* row_ptr, const int32_t* col_idx, int32_t nbr, int32_t nbc, Plan* plan);

}  // namespace accel
