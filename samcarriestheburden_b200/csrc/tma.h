// Host helper: cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace b200sam {
// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row pitch ld (elements);
// box = [box_rows, box_cols]; swizzle span must equal box_cols * 2 bytes (32 / 64 / 128 B) or be NONE.
int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   uint32_t box_cols, CUtensorMapSwizzle swizzle);
// 4-D bf16 tensor map over a token grid: qkv viewed as [B][64 (y)][64 (x)][cols]; box = [1][box_y][box_x][16 cols],
// SWIZZLE_32B.  One TMA op fetches a (box_y x box_x) patch of tokens as a slab of 32-byte rows in (y, x) order;
// out-of-grid tokens are zero-filled.
int make_tmap_bf16_grid4d(CUtensorMap* map, const void* ptr, uint64_t batch, uint64_t cols, uint32_t box_x,
                          uint32_t box_y);
int num_sms();
}  // namespace b200sam
