// Deterministic gradient accumulation for rows that repeat inside a batch (items; users when a host batch repeats them).
//
// TF1 sums duplicate IndexedSlices rows before the optimizer sees them (MF/model_api.py:83 -> AdamOptimizer ->
// _deduplicate_indexed_slices: unique + unsorted_segment_sum).  The default path reduces the per-triple gradient rows
// with fp32 red.global.add in L2, whose order is not fixed: trajectories with duplicates match the oracle to 1e-5 per
// step, not bit for bit.  In deterministic mode the step kernel STORES every per-triple row into a slot buffer
// (slot = position in the concatenated IndexedSlices: pos 0..B-1, then neg B..2B-1) and this file sums the slots of
// each row in ascending slot order -- the occurrence order of the oracle's dedup_sum (oracle/pda_oracle.py:409,
// oracle/csrc/pda_oracle.c:orc_scatter_add_rows) -- so that the whole trajectory is bit-identical to the oracle:
//   1. stable radix sort of (row id, slot) by row id (cub::DeviceRadixSort: library plumbing, not a hot-path kernel),
//   2. one group of lanes per run of equal row ids: g = s_first; g = g + s_next ... -> G[row] (plain store).
#include <cub/device/device_radix_sort.cuh>

#include "pda_kernels.h"

namespace pda {

__global__ void segsum_keys_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t B, int32_t* __restrict__ keys,
                                   int32_t* __restrict__ slots) {
    const int64_t n = b ? 2 * B : B;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = i < B ? a[i] : b[i - B];
        slots[i] = (int32_t)i;
    }
}

// sorted keys / slots; one warp per run head, lane-strided float4 chunks of the row; rows[slot] = [d] floats
__global__ void __launch_bounds__(256) segsum_rows_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ slots, int64_t n,
                                                          const float* __restrict__ rows, int d, float* __restrict__ G) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int q = d >> 2;
    for (int64_t i = warp_global; i < n; i += n_warps) {
        const int32_t k = keys[i];
        if (i > 0 && keys[i - 1] == k) continue;             // not a run head
        int64_t e = i + 1;
        while (e < n && keys[e] == k) ++e;
        for (int ch = lane; ch < q; ch += 32) {
            float4 g = *reinterpret_cast<const float4*>(rows + (int64_t)slots[i] * d + 4 * ch);     // 0 + s == s
            for (int64_t z = i + 1; z < e; ++z) {
                const float4 s = *reinterpret_cast<const float4*>(rows + (int64_t)slots[z] * d + 4 * ch);
                g.x = fadd(g.x, s.x); g.y = fadd(g.y, s.y); g.z = fadd(g.z, s.z); g.w = fadd(g.w, s.w);
            }
            *reinterpret_cast<float4*>(G + (int64_t)k * d + 4 * ch) = g;
        }
    }
}

size_t segsum_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (int)n);
    return bytes;
}

// ids a[0..B) (and b[0..B) when given) -> G[row] = ordered sum of rows[slot]; work: 4 int32 arrays of 2B + temp
int launch_segment_sum(const int32_t* a, const int32_t* b, int64_t B, const float* rows, int d, float* G, int32_t* work, void* temp,
                       size_t temp_bytes, int key_bits, cudaStream_t st) {
    const int64_t n = b ? 2 * B : B;
    int32_t *keys = work, *slots = work + n, *keys_s = work + 2 * n, *slots_s = work + 3 * n;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    segsum_keys_kernel<<<(int)blocks, 256, 0, st>>>(a, b, B, keys, slots);
    if (cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_s, slots, slots_s, (int)n, 0, key_bits, st) != cudaSuccess) return 1;
    int64_t wb = (n + 7) / 8;
    if (wb > 148 * 16) wb = 148 * 16;
    segsum_rows_kernel<<<(int)wb, 256, 0, st>>>(keys_s, slots_s, n, rows, d, G);
    return 0;
}

}  // namespace pda
