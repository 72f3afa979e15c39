// Host-side MMA schedule for one BSR weight matrix (Convention B, 14x14 blocks).
//
// The reference stores a block-row as a list of 14x14 int8 blocks (196 B, row-major) at K tiles
// col_idx[] (sw/training/export_bsr_14x14.py:241-272).  The tensor core wants K=32 bytes per
// instruction and 16-byte core matrices, so at load time every block is re-laid into a
// "B tile" of 16 rows x 32 K-bytes:  two 16-byte K slots, each holding one 14-wide block
// (+2 zero bytes).  Two stored blocks at adjacent K tiles of the same block-row share one tile
// (one tcgen05.mma); a lone block takes half a tile.  The kernel walks, per group of block-rows,
// a list of fixed-capacity batches of such tiles that is streamed by bulk async copies.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace accel {

constexpr int kBlock = 14;           // reference block size (export_bsr_14x14.py:48)
constexpr int kTile = 16;            // padded K slot / padded block-row height
constexpr int kChunkTiles = 9;       // K tiles per activation stage: 126 k = 14 channels x 9 taps of a 3x3 conv
constexpr int kMaxGroupRows = 11;    // block-rows per CTA: 11 x 16 accumulator columns + 2 x 36 activation columns <= 256
constexpr int kDefaultGroupRows = 11;
constexpr int kOpsPerBatch = 32;     // B tiles per weight stage (one per lane of the decoding warp)
constexpr int kBTileBytes = 512;     // 16 rows x 32 bytes
constexpr int kBatchMetaBytes = 64;  // kOpsPerBatch x u16
constexpr int kBatchBytes = kOpsPerBatch * kBTileBytes + kBatchMetaBytes;

struct BatchInfo {       // 8 bytes, read by the loader / issuer / producer warps
  uint32_t blob_off16;   // byte offset / 16 of the batch blob inside the workspace
  uint16_t chunk;        // K chunk (kChunkTiles tiles) all ops of this batch live in
  uint8_t n_ops;         // 1..kOpsPerBatch
  uint8_t flags;         // bit0: first batch of its chunk, bit1: last batch of its chunk
};

struct GroupInfo {       // 32 bytes, one per CTA column
  int32_t br0;           // first block-row of the group
  int32_t n_rows;        // block-rows in the group (<= kMaxGroupRows)
  uint32_t nonempty;     // bit g: block-row br0+g has at least one stored block
  int32_t batch_begin;   // [batch_begin, batch_end) in the BatchInfo array
  int32_t batch_end;
  int32_t n_steps;       // number of distinct K chunks touched (activation stages to produce)
  int32_t tmem_cols;     // power of two >= 32 covering n_rows*16
  int32_t pad_;
};

struct OpSrc {           // repack kernel input: which stored blocks feed B tile `op`
  int32_t blk_lo;        // block placed in K slot 0 (bytes 0..13), -1 = none
  int32_t blk_hi;        // block placed in K slot 1 (bytes 16..29), -1 = none
};

struct Plan {
  int32_t nbr = 0, nbc = 0;
  int64_t nnz = 0;
  int32_t n_chunks = 0;
  int32_t group_rows = 0;
  std::vector<GroupInfo> groups;
  std::vector<BatchInfo> batches;
  std::vector<uint16_t> op_meta;   // per op: (g & 15) | (window_tile << 4)
  std::vector<OpSrc> op_src;       // per op
  std::vector<uint32_t> op_blob_off;  // byte offset of each op's B tile in the workspace
  std::vector<uint32_t> op_meta_off;  // byte offset of each op's u16 meta in the workspace
  // device workspace layout (byte offsets)
  size_t off_blob = 0, off_batches = 0, off_groups = 0, off_opsrc = 0, off_opoff = 0, off_opmeta = 0, off_opmoff = 0;
  size_t ws_bytes = 0;
  int64_t n_ops = 0;
  // filled by upload
  const uint8_t* ws_dev = nullptr;
  bool uploaded = false;
};

// Returns empty string on success, else a validate_bsr-style message (bsr_packer.hpp:364-436).
std::string build_plan(const int32_t* row_ptr, const int32_t* col_idx, int32_t nbr, int32_t nbc, Plan* plan);

}  // namespace accel
