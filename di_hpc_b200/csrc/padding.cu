// padding.cu -- ragged-tensor Pad / GroupPad / Unpad (1-D, 2-D, 3-D) for sm_100a, and the host-side
// group splitters.  This is the data format on the input side of the trajectory-return path
// (ragged trajectories -> padded batches), SURVEY.md 8(f) item 4.
//
// Semantics: hpc_rll/origin/padding.py:47-56 (_Padding1D), 107-121 (Padding2D), 134-148 (Padding3D),
// 88-96 / 124-131 / 151-158 (UnPadding*), as exposed by hpc_rll/rl_utils/padding.py:
//   new_x[i, :shape_i] = x_i, elsewhere `value`;  mask[i, :shape_i] = 1, elsewhere `value` (int32 mask,
//   src/rl_utils/padding.cu:126-127).  Group mode pads every group of a size-sorted list to its own max.
// Replaces Pad{1,2,3}DForward / GroupPad{1,2,3}DForward / Unpad{1,2,3}DForward and their 9 kernels
// (src/rl_utils/padding.cu:110-589, include/hpc/rll/cuda/rl_utils/padding_kernel.h:100-233), which per
// call do up to 7 cudaMalloc + 7 blocking cudaMemcpy + 7 cudaFree and one block per tensor.
//
// B200 design: the per-tensor descriptors (source, destination, mask pointers + shapes) travel in the
// kernel PARAMETER space (`__grid_constant__`, 512 items = 24 KB per launch), so a call allocates nothing
// and copies nothing; one launch covers every tensor of every group, pad and unpad share the kernel, and
// the grid is 2-D (item x element block) so large tensors are spread over many CTAs.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace hpcrll {

struct PadItem {
    const float* src;
    float* dst;
    int32_t* mask;  // pad only
    int s[3];       // own shape, right-aligned (leading dims 1 for lower rank)
    int m[3];       // padded shape of its group, right-aligned
};
constexpr int kPadChunk = 512;
struct PadBatch {
    PadItem it[kPadChunk];
};

// Shapes are right-aligned into 3 dims (1-D: (1,1,L); 2-D: (1,A,B)), so an item is a list of ROWS of
// d2 contiguous elements.  One warp per row at a time, lanes along the row: no per-element division,
// coalesced reads and writes; rows are spread over the 8 warps of a CTA and over gridDim.y CTAs.
template <bool UNPAD, bool VEC>
__global__ void __launch_bounds__(256) pad_kernel(const __grid_constant__ PadBatch batch, float value, int ivalue) {
    const PadItem& p = batch.it[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = p.s[0], s1 = p.s[1], s2 = p.s[2], m1 = p.m[1], m2 = p.m[2];
    const int d0 = UNPAD ? s0 : p.m[0], d1 = UNPAD ? s1 : m1;
    const int64_t rows = static_cast<int64_t>(d0) * d1;
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + warp; r < rows; r += static_cast<int64_t>(gridDim.y) * 8) {
        const int a = static_cast<int>(r / d1), b = static_cast<int>(r - static_cast<int64_t>(a) * d1);
        if (UNPAD) {
            const float* __restrict__ src = p.src + (static_cast<int64_t>(a) * m1 + b) * m2;
            float* __restrict__ dst = p.dst + r * s2;
            if (VEC) {  // padded rows are 16-byte aligned (m2 % 4 == 0): 128-bit loads, scalar stores
                for (int c = lane * 4; c < s2; c += 128) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
                    dst[c] = v.x;
                    if (c + 1 < s2) dst[c + 1] = v.y;
                    if (c + 2 < s2) dst[c + 2] = v.z;
                    if (c + 3 < s2) dst[c + 3] = v.w;
                }
            } else {
                for (int c = lane; c < s2; c += 32) dst[c] = __ldg(src + c);
            }
        } else {
            const bool row_in = a < s0 && b < s1;
            const float* __restrict__ src = p.src + (static_cast<int64_t>(a) * s1 + b) * s2;
            float* __restrict__ dst = p.dst + r * m2;
            int32_t* __restrict__ msk = p.mask + r * m2;
            if (VEC) {  // scalar loads from the ragged source row, 128-bit stores of data and mask
                const int lim = row_in ? s2 : 0;
                for (int c = lane * 4; c < m2; c += 128) {
                    float4 v;
                    int4 k;
                    v.x = c < lim ? __ldg(src + c) : value;
                    v.y = c + 1 < lim ? __ldg(src + c + 1) : value;
                    v.z = c + 2 < lim ? __ldg(src + c + 2) : value;
                    v.w = c + 3 < lim ? __ldg(src + c + 3) : value;
                    k.x = c < lim ? 1 : ivalue;
                    k.y = c + 1 < lim ? 1 : ivalue;
                    k.z = c + 2 < lim ? 1 : ivalue;
                    k.w = c + 3 < lim ? 1 : ivalue;
                    *reinterpret_cast<float4*>(dst + c) = v;
                    *reinterpret_cast<int4*>(msk + c) = k;
                }
            } else {
                for (int c = lane; c < m2; c += 32) {
                    const bool inside = row_in && c < s2;
                    dst[c] = inside ? __ldg(src + c) : value;
                    msk[c] = inside ? 1 : ivalue;
                }
            }
        }
    }
}

static int pad_batch(const float* const* src, float* const* dst, int32_t* const* mask, const int32_t* shapes,
                     const int32_t* padded, int64_t n, int value, bool unpad, cudaStream_t stream) {
    HPC_REQUIRE(n >= 0, "pad: negative item count");
    if (n == 0) return HPC_RLL_OK;
    HPC_REQUIRE(src && dst && shapes && padded && (unpad || mask), "pad: null table");
    for (int64_t base = 0; base < n; base += kPadChunk) {
        const int cnt = static_cast<int>(std::min<int64_t>(kPadChunk, n - base));
        PadBatch batch;
        int64_t max_total = 1;
        bool vec = true;  // every padded row 16-byte aligned?
        for (int i = 0; i < cnt; ++i) {
            PadItem& p = batch.it[i];
            const int64_t k = base + i;
            p.src = src[k];
            p.dst = dst[k];
            p.mask = unpad ? nullptr : mask[k];
            int64_t tot = 1, own = 1;
            for (int d = 0; d < 3; ++d) {
                p.s[d] = shapes[k * 3 + d];
                p.m[d] = padded[k * 3 + d];
                HPC_REQUIRE(p.s[d] >= 0 && p.m[d] >= p.s[d], "pad: item %lld dim %d: shape %d exceeds padded %d",
                            (long long)k, d, p.s[d], p.m[d]);
                tot *= unpad ? p.s[d] : p.m[d];
                own *= p.s[d];
            }
            // an empty source tensor has no storage (null pointer) and is never dereferenced
            HPC_REQUIRE((own == 0 || p.src) && (tot == 0 || (p.dst && (unpad || p.mask))),
                        "pad: null tensor pointer at item %lld", (long long)k);
            max_total = std::max(max_total, tot);
            const void* padded_base = unpad ? static_cast<const void*>(p.src) : static_cast<const void*>(p.dst);
            vec = vec && (p.m[2] % 4 == 0) && aligned16(padded_base) && (unpad || aligned16(p.mask));
        }
        // ~2048 elements per CTA pass; enough CTAs per item to spread big tensors over the machine
        int64_t by = (max_total + 2047) / 2048;
        by = std::min<int64_t>(std::max<int64_t>(by, 1), 2048);
        const dim3 grid(static_cast<unsigned>(cnt), static_cast<unsigned>(by));
        const float fv = static_cast<float>(value);
        if (unpad) {
            if (vec) pad_kernel<true, true><<<grid, 256, 0, stream>>>(batch, fv, value);
            else pad_kernel<true, false><<<grid, 256, 0, stream>>>(batch, fv, value);
        } else {
            if (vec) pad_kernel<false, true><<<grid, 256, 0, stream>>>(batch, fv, value);
            else pad_kernel<false, false><<<grid, 256, 0, stream>>>(batch, fv, value);
        }
        count_launch();
        HPC_LAUNCH_CHECK();
    }
    return HPC_RLL_OK;
}

}  // namespace hpcrll

extern "C" {

int hpc_rll_pad_batch(const float* const* src, float* const* dst, int32_t* const* mask, const int32_t* shapes,
                      const int32_t* padded, int64_t n, int value, void* stream) {
    HPC_NVTX("pad_batch");
    return hpcrll::pad_batch(src, dst, mask, shapes, padded, n, value, false, hpcrll::as_stream(stream));
}

int hpc_rll_unpad_batch(const float* const* src, float* const* dst, const int32_t* shapes, const int32_t* padded,
                        int64_t n, void* stream) {
    HPC_NVTX("unpad_batch");
    return hpcrll::pad_batch(src, dst, nullptr, shapes, padded, n, 0, true, hpcrll::as_stream(stream));
}

// Optimal split of a size-sorted list into exactly `group` consecutive groups minimising the padded volume
//   sum_g  count_g * prod_d max_{i in g} shape_i[d]
// (the cost model of the reference's C++ splitter, src/rl_utils/padding.cu:44-108; hpc_rll/origin/padding.py:12-45
// is the same DP on element counts).  O(group * n^2) time, O(group * n) memory.  positions[0..group] receives the
// group boundaries (positions[0] = 0, positions[group] = n).  Ties pick the earliest split point.
int hpc_rll_oracle_split_group(const int64_t* shapes, int64_t n, int ndim, int group, int64_t* positions) {
    using namespace hpcrll;
    HPC_REQUIRE(shapes && positions, "oracle_split_group: null pointer");
    HPC_REQUIRE(n >= 1 && ndim >= 1 && ndim <= 3, "oracle_split_group: need n >= 1 and 1 <= ndim <= 3");
    HPC_REQUIRE(group >= 1 && group <= n, "oracle_split_group: group must be in [1, n]");
    const int64_t INF = INT64_MAX / 4;
    std::vector<int64_t> cost(static_cast<size_t>((n + 1) * (group + 1)), INF);
    std::vector<int64_t> from(static_cast<size_t>((n + 1) * (group + 1)), -1);
    auto at = [&](int64_t i, int j) -> size_t { return static_cast<size_t>(i * (group + 1) + j); };
    cost[at(0, 0)] = 0;
    for (int64_t i = 1; i <= n; ++i) {
        for (int j = 1; j <= group && j <= i; ++j) {
            int64_t mx[3] = {0, 0, 0};
            int64_t best = INF, best_k = -1;
            for (int64_t k = i - 1; k >= j - 1; --k) {  // group = items k .. i-1
                int64_t vol = 1;
                for (int d = 0; d < ndim; ++d) {
                    mx[d] = std::max(mx[d], shapes[k * ndim + d]);
                    vol *= mx[d];
                }
                if (cost[at(k, j - 1)] >= INF) continue;
                const int64_t c = cost[at(k, j - 1)] + vol * (i - k);
                if (c <= best) {  // descending k with <= : the smallest k among the minima wins
                    best = c;
                    best_k = k;
                }
            }
            cost[at(i, j)] = best;
            from[at(i, j)] = best_k;
        }
    }
    HPC_REQUIRE(from[at(n, group)] >= 0, "oracle_split_group: no feasible split");
    int64_t pos = n;
    for (int j = group; j >= 1; --j) {
        positions[j] = pos;
        pos = from[at(pos, j)];
    }
    positions[0] = 0;
    return HPC_RLL_OK;
}

// Random split ("sample" mode, hpc_rll/origin/padding.py:64-78 / src/rl_utils/padding.cu:8-43): draw group-1
// random boundaries, close each group at a sampled item, then merge neighbours whose padded shape is equal.
// starts[0..*n_groups] receives the group start indices followed by n.  Deterministic for a given seed.
int hpc_rll_sample_split_group(const int64_t* shapes, int64_t n, int ndim, int group, uint64_t seed,
                               int64_t* starts, int* n_groups) {
    using namespace hpcrll;
    HPC_REQUIRE(shapes && starts && n_groups, "sample_split_group: null pointer");
    HPC_REQUIRE(n >= 1 && ndim >= 1 && ndim <= 3 && group >= 1, "sample_split_group: bad sizes");
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    auto next = [&]() {
        st ^= st << 13;
        st ^= st >> 7;
        st ^= st << 17;
        return st;
    };
    std::vector<int64_t> ends;  // index of the last item of each group
    if (n > 2)
        for (int g = 0; g < group - 1; ++g) ends.push_back(1 + static_cast<int64_t>(next() % static_cast<uint64_t>(n - 2)));
    ends.push_back(n - 1);
    std::sort(ends.begin(), ends.end());
    ends.erase(std::unique(ends.begin(), ends.end()), ends.end());
    int count = 0;
    int64_t start = 0;
    int64_t prev_shape[3] = {-1, -1, -1};
    for (int64_t e : ends) {
        int64_t mx[3] = {0, 0, 0};
        for (int64_t i = start; i <= e; ++i)
            for (int d = 0; d < ndim; ++d) mx[d] = std::max(mx[d], shapes[i * ndim + d]);
        const bool same = count > 0 && mx[0] == prev_shape[0] && mx[1] == prev_shape[1] && mx[2] == prev_shape[2];
        if (!same) {
            starts[count++] = start;
            for (int d = 0; d < 3; ++d) prev_shape[d] = mx[d];
        }
        start = e + 1;
    }
    starts[count] = n;
    *n_groups = count;
    return HPC_RLL_OK;
}

}  // extern "C"
