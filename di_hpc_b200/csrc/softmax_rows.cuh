// softmax_rows.cuh -- streaming row-softmax building blocks for the (T,B,N) / (B,N) logits tensors
// of V-trace, UPGO and PPO.
//
// The reference spends one 256-thread block per row and five block-wide reductions
// (include/hpc/rll/cuda/rl_utils/vtrace_kernel.h:11-112, upgo_kernel.h:40-81,
// ppo_kernel.h:12-112) and materialises three (T,B,N) gradient buffers in the forward pass.
// Here a row lives in the registers of a sub-warp group of G lanes (G = 1..32, a power of two chosen
// so that every lane owns up to four 128-bit chunks: N=16 -> one lane per row, N=128 -> 8 lanes),
// is loaded once with 128-bit loads, reduced with shuffles only when G > 1, and needs ONE exp per
// element:
//     m = max x,  e_i = exp(x_i - m),  s = sum e_i,  t = sum e_i (x_i - m)
//     log p_i = x_i - (m + log s),  p_i = e_i / s,  H = -sum p log p = log s - t / s
// The backward recomputes the softmax from the logits instead of reading saved (T,B,N) buffers:
//     d/dx_k = c1*(1[k=a] - p_k) + c2*(-p_k*(log p_k + H))            
// HBM bytes per row: forward 4N (+8 action, +4 per per-row scalar), backward 4N in + 4N out.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace hpcrll {

// geometry of the row -> lane mapping, decided on the host
struct RowGeom {
    int G;      // lanes per row (power of two <= 32)
    int kmax;   // register chunks per lane, exact: 1..8 (0 = N too large for registers: looping kernel)
    int width;  // floats per chunk: 4 (N % 4 == 0, 16B-aligned bases), 2 (N even, 8B-aligned: Atari's 6 / 18), 1
};

// p0 / p1: the tensors the kernel reads or writes row-wise (p1 may be null)
inline RowGeom row_geom(int64_t N, const void* p0, const void* p1 = nullptr) {
    RowGeom g;
    const bool a16 = aligned16(p0) && (!p1 || aligned16(p1)), a8 = aligned8(p0) && (!p1 || aligned8(p1));
    g.width = (a16 && N % 4 == 0) ? 4 : ((a8 && N % 2 == 0) ? 2 : 1);
    const int64_t chunks = N / g.width;  // chunk = one load per lane
    // lanes per row: the smallest power of two that leaves <= 4 float4 (or <= 8 narrower) chunks per lane; the
    // chunk count per lane is EXACT (kernels are instantiated for 1..8), so N=6 is one lane x 3 float2, N=18 two
    // lanes x 5 float2, N=24 two lanes x 3 float4 -- the row kernels are issue-bound, idle register slots cost
    const int64_t cap = g.width == 4 ? 4 : 8;
    int G = 1;
    while (G < 32 && chunks > static_cast<int64_t>(G) * cap) G <<= 1;
    g.G = G;
    const int64_t per_lane = (chunks + G - 1) / G;
    g.kmax = per_lane <= 8 ? static_cast<int>(per_lane) : 0;
    return g;
}

#ifdef __CUDACC__

// stands in for -inf in masked register slots and clamps (x - m): finite, so 0 * it stays 0 -- the
// same guard torch.distributions.Categorical.entropy applies (clamp to finfo.min) for -inf logits
constexpr float kNegBig = -3.0e38f;

// e^d for the N-wide softmax terms (d = x - max <= 0, or a log-probability): FMUL + MUFU.EX2 instead of
// expf's 9 instructions -- the row kernels are issue-bound at small N (profiles/r01_ncu_ops.md).
// Error budget: ex2.approx is good to 2 ulp; rounding d*log2(e) adds a relative |d| * 6e-8, so a term e^d
// is off by at most |d| e^d * 6e-8 <= 2.2e-8 of the row's largest term -- far inside the 1e-5 parity bar
// (SURVEY.md 8c).  Per-row scalars (importance ratios) keep expf.  exp_term(-3e38) = 0 like expf.
__device__ __forceinline__ float exp_term(float d) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d * 1.4426950408889634f));
    return r;
}

__device__ __forceinline__ float group_max(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float group_sum(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One row held across a group of G lanes.  Element index of x[j*W+q] is ((j*G + lig)*W + q).
template <int KMAX, int WIDTH>
struct RowRegs {
    static_assert(WIDTH == 1 || WIDTH == 2 || WIDTH == 4, "chunk width");
    static constexpr int W = WIDTH;
    static constexpr int NE = KMAX * W;
    float x[NE];

    __device__ __forceinline__ int index(int j, int q, int G, int lig) const { return (j * G + lig) * W + q; }

    // default-cached (L1-allocating) loads: with G < 8 a 128-byte line is shared by several chunk
    // indices j, so later j hit in L1 instead of re-fetching sectors from L2
    __device__ __forceinline__ void load(const float* __restrict__ row, int N, int G, int lig, bool active) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            const int e0 = (j * G + lig) * W;
            if (W == 4) {
                float4 v = make_float4(kNegBig, kNegBig, kNegBig, kNegBig);
                if (active && e0 < N) v = __ldg(reinterpret_cast<const float4*>(row + e0));
                x[j * W + 0] = v.x;
                x[j * W + (W > 1 ? 1 : 0)] = v.y;
                x[j * W + (W > 2 ? 2 : 0)] = v.z;
                x[j * W + (W > 3 ? 3 : 0)] = v.w;
            } else if (W == 2) {
                float2 v = make_float2(kNegBig, kNegBig);
                if (active && e0 < N) v = __ldg(reinterpret_cast<const float2*>(row + e0));
                x[j * W + 0] = v.x;
                x[j * W + (W > 1 ? 1 : 0)] = v.y;
            } else {
                x[j] = (active && e0 < N) ? __ldg(row + e0) : kNegBig;
            }
        }
    }

    __device__ __forceinline__ float row_max(int G) const {
        float mm = x[0];
#pragma unroll
        for (int i = 1; i < NE; ++i) mm = fmaxf(mm, x[i]);
        return group_max(mm, G);
    }

    // one pass: s = sum exp(x-m), t = sum exp(x-m)*(x-m)  (group-wide); optionally keeps e_i
    template <bool WANT_T, bool KEEP_E>
    __device__ __forceinline__ void stats(int G, float m, float& s, float& t, float (&e)[NE]) const {
        float ss = 0.f, tt = 0.f;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            const float d = fmaxf(x[i] - m, kNegBig);
            const float ei = exp_term(d);  // masked slots: exp(-3e38) = 0
            ss += ei;
            if (WANT_T) tt = fmaf(ei, d, tt);
            if (KEEP_E) e[i] = ei;
        }
        s = group_sum(ss, G);
        t = WANT_T ? group_sum(tt, G) : 0.f;
    }

    // x[index == a] (0 if this lane does not hold it); sum over the group afterwards
    __device__ __forceinline__ float select(int a, int G, int lig) const {
        float sel = 0.f;
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
#pragma unroll
            for (int q = 0; q < W; ++q)
                if (index(j, q, G, lig) == a) sel = x[j * W + q];
        return sel;
    }
};

// Per-row log-prob in the two normalisations origin uses:
//   CATEGORICAL: logp = x - (m + log s)          torch.distributions.Categorical (vtrace.py:74, ppo.py:54)
//   otherwise  : logp = (x - m) - log s          F.cross_entropy / log_softmax   (upgo.py:16)
template <bool CATEGORICAL>
__device__ __forceinline__ float row_logp(float x, float m, float logs) {
    return CATEGORICAL ? x - (m + logs) : (x - m) - logs;
}

// ---------------------------------------------------------------------------------------------------
// Staged rows: action counts that are not a multiple of 4 (Atari's 6 / 9 / 18 ...) cannot use 128-bit
// row chunks, and scalar per-lane chunks waste lanes and shuffles.  For N <= 32 such tensors are instead
// read as a FLAT contiguous stream: a CTA copies kStageRows consecutive rows (kStageRows*N floats, 128-bit
// coalesced, row boundaries ignored) into shared memory with an odd pitch, then every thread owns ONE row
// (bank-conflict-free column walk, no shuffles, all lanes busy).  Gradients go back the same way.
// ---------------------------------------------------------------------------------------------------
constexpr int kStageRows = 256;
inline int stage_pitch(int N) { return N | 1; }
// measured (profiles/r01_ops.md): for N <= 8 the per-lane scalar chunks are as fast or faster; from N = 9 the
// staged path wins (N=18: 3.3 -> 2.0 ms for V-trace at T=512, B=32768).  Rows that can be read in 64-bit or
// 128-bit chunks (`width` > 1) stay on the register path.
inline bool use_staged_rows(int64_t N, int width) { return N > 8 && N <= 32 && width == 1; }
inline size_t stage_bytes(int N, int tiles) { return static_cast<size_t>(tiles) * kStageRows * stage_pitch(N) * sizeof(float); }

// copy rows [row0, row0 + kStageRows) of a contiguous (R, N) tensor into tile[r * P + c]
__device__ __forceinline__ void stage_rows(const float* __restrict__ g, int64_t R, int N, int P, int64_t row0,
                                           float* __restrict__ tile, bool aligned) {
    const int64_t rows = min(static_cast<int64_t>(kStageRows), R - row0);
    const int cnt = static_cast<int>(rows) * N;  // <= 256 * 32
    const float* __restrict__ src = g + row0 * N;
    const float invN = 1.f / static_cast<float>(N);
    if (aligned) {  // (row0 * N) % 4 == 0 because row0 is a multiple of kStageRows
        for (int i = threadIdx.x * 4; i < cnt; i += kStageRows * 4) {
            const int r = __float2int_rd((static_cast<float>(i) + 0.5f) * invN);  // exact for i < 2^13, N <= 32
            int c = i - r * N;
            int idx = r * P + c;  // running shared-memory index: +1 per element, +(P-N) more at a row end
            if (i + 3 < cnt) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tile[idx++] = e[q];
                    if (++c == N) c = 0, idx += P - N;
                }
            } else {
                for (int q = 0; i + q < cnt; ++q) {
                    tile[idx++] = __ldg(src + i + q);
                    if (++c == N) c = 0, idx += P - N;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < cnt; i += kStageRows) {
            const int r = __float2int_rd((static_cast<float>(i) + 0.5f) * invN);
            tile[r * P + (i - r * N)] = __ldg(src + i);
        }
    }
}

// inverse: write tile rows back to a contiguous (R, N) tensor
__device__ __forceinline__ void unstage_rows(float* __restrict__ g, int64_t R, int N, int P, int64_t row0,
                                             const float* __restrict__ tile, bool aligned) {
    const int64_t rows = min(static_cast<int64_t>(kStageRows), R - row0);
    const int cnt = static_cast<int>(rows) * N;
    float* __restrict__ dst = g + row0 * N;
    const float invN = 1.f / static_cast<float>(N);
    if (aligned) {
        for (int i = threadIdx.x * 4; i < cnt; i += kStageRows * 4) {
            const int r = __float2int_rd((static_cast<float>(i) + 0.5f) * invN);
            int c = i - r * N;
            int idx = r * P + c;
            float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (i + q < cnt) e[q] = tile[idx];
                ++idx;
                if (++c == N) c = 0, idx += P - N;
            }
            if (i + 3 < cnt) {
                st_stream4(reinterpret_cast<float4*>(dst + i), make_float4(e[0], e[1], e[2], e[3]));
            } else {
                for (int q = 0; i + q < cnt; ++q) dst[i + q] = e[q];
            }
        }
    } else {
        for (int i = threadIdx.x; i < cnt; i += kStageRows) {
            const int r = __float2int_rd((static_cast<float>(i) + 0.5f) * invN);
            dst[i] = tile[r * P + (i - r * N)];
        }
    }
}

// statistics of one row sitting in shared memory (one thread per row)
template <bool WANT_T>
__device__ __forceinline__ void staged_stats(const float* __restrict__ x, int N, float& m, float& s, float& t) {
    float mm = x[0];
    for (int k = 1; k < N; ++k) mm = fmaxf(mm, x[k]);
    float ss = 0.f, tt = 0.f;
    for (int k = 0; k < N; ++k) {
        const float d = fmaxf(x[k] - mm, kNegBig);
        const float e = exp_term(d);
        ss += e;
        if (WANT_T) tt = fmaf(e, d, tt);
    }
    m = mm;
    s = ss;
    t = tt;
}

#endif  // __CUDACC__

// ---- shared backward: grad[r,k] = (g1*c1[r])*(1[k=a_r]-p_k) + (g2*w_r*inv_n)*(-p_k(logp_k+H)) ----------
// g1, g2: device scalars (upstream gradients); g2 == nullptr skips the entropy term; w == nullptr -> 1.
// categorical selects the log-prob normalisation (see row_logp).
int launch_softmax_grad_rows(const float* logits, const int64_t* action, const float* c1, const float* w,
                             const float* g1, const float* g2, double inv_n, float* grad, int64_t R, int64_t N,
                             bool categorical, cudaStream_t stream);

// grid size for a row-streaming kernel: enough CTAs to fill the machine, grid-stride over row blocks
inline unsigned rows_grid(int64_t R, int rows_per_block) {
    int64_t blocks = (R + rows_per_block - 1) / rows_per_block;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<unsigned>(blocks);
}

// launch-time descriptor for finalize_terms: term k sums partials[off[k] .. off[k]+cnt[k]) * scale[k]
struct FinSpec {
    int off[5];
    int cnt[5];
    double scale[5];
};
int launch_finalize_terms(const double* partials, const FinSpec& spec, int nterms, float* out, cudaStream_t stream);

// dispatch helper: expands to the KMAX/WIDTH instantiation selected by a RowGeom
#define HPC_ROW_DISPATCH_W(ge, LAUNCH, WD)      \
    switch ((ge).kmax) {                        \
        case 1: LAUNCH(1, WD); break;           \
        case 2: LAUNCH(2, WD); break;           \
        case 3: LAUNCH(3, WD); break;           \
        case 4: LAUNCH(4, WD); break;           \
        case 5: LAUNCH(5, WD); break;           \
        case 6: LAUNCH(6, WD); break;           \
        case 7: LAUNCH(7, WD); break;           \
        default: LAUNCH(8, WD); break;          \
    }
#define HPC_ROW_DISPATCH(ge, LAUNCH)                                 \
    do {                                                             \
        if ((ge).width == 4) { HPC_ROW_DISPATCH_W(ge, LAUNCH, 4) }   \
        else if ((ge).width == 2) { HPC_ROW_DISPATCH_W(ge, LAUNCH, 2) } \
        else { HPC_ROW_DISPATCH_W(ge, LAUNCH, 1) }                   \
    } while (0)

}  // namespace hpcrll
