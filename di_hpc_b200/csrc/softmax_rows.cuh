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
    int kmax;   // register chunks per lane: 1, 2, 4 or 8 (0 = N too large for registers: looping kernel)
    int vec;    // 1: float4 chunks (N % 4 == 0 and 16B-aligned base), 0: scalar chunks
};

inline RowGeom row_geom(int64_t N, bool aligned) {
    RowGeom g;
    g.vec = (aligned && (N % 4) == 0) ? 1 : 0;
    const int64_t chunks = g.vec ? N / 4 : N;  // chunk = one load per lane
    int G = 1;
    while (G < 32 && chunks > static_cast<int64_t>(G) * 4) G <<= 1;  // aim at <= 4 chunks per lane
    g.G = G;
    const int64_t per_lane = (chunks + G - 1) / G;
    g.kmax = per_lane <= 1 ? 1 : (per_lane <= 2 ? 2 : (per_lane <= 4 ? 4 : (per_lane <= 8 ? 8 : 0)));
    return g;
}

#ifdef __CUDACC__

// stands in for -inf in masked register slots and clamps (x - m): finite, so 0 * it stays 0 -- the
// same guard torch.distributions.Categorical.entropy applies (clamp to finfo.min) for -inf logits
constexpr float kNegBig = -3.0e38f;

__device__ __forceinline__ float group_max(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float group_sum(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One row held across a group of G lanes.  Element index of x[j*W+q] is ((j*G + lig)*W + q).
template <int KMAX, bool VEC>
struct RowRegs {
    static constexpr int W = VEC ? 4 : 1;
    static constexpr int NE = KMAX * W;
    float x[NE];

    __device__ __forceinline__ int index(int j, int q, int G, int lig) const { return (j * G + lig) * W + q; }

    // default-cached (L1-allocating) loads: with G < 8 a 128-byte line is shared by several chunk
    // indices j, so later j hit in L1 instead of re-fetching sectors from L2
    __device__ __forceinline__ void load(const float* __restrict__ row, int N, int G, int lig, bool active) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            const int e0 = (j * G + lig) * W;
            if (VEC) {
                float4 v = make_float4(kNegBig, kNegBig, kNegBig, kNegBig);
                if (active && e0 < N) v = __ldg(reinterpret_cast<const float4*>(row + e0));
                x[j * W + 0] = v.x;
                x[j * W + (W > 1 ? 1 : 0)] = v.y;
                x[j * W + (W > 2 ? 2 : 0)] = v.z;
                x[j * W + (W > 3 ? 3 : 0)] = v.w;
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
            const float ei = expf(d);  // masked slots: exp(-3e38) = 0
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

// dispatch helper: expands to the KMAX/VEC instantiation selected by a RowGeom
#define HPC_ROW_DISPATCH(ge, LAUNCH)                       \
    do {                                                   \
        if ((ge).vec) {                                    \
            if ((ge).kmax == 1) { LAUNCH(1, true); }       \
            else if ((ge).kmax == 2) { LAUNCH(2, true); }  \
            else if ((ge).kmax == 4) { LAUNCH(4, true); }  \
            else { LAUNCH(8, true); }                      \
        } else {                                           \
            if ((ge).kmax == 1) { LAUNCH(1, false); }      \
            else if ((ge).kmax == 2) { LAUNCH(2, false); } \
            else if ((ge).kmax == 4) { LAUNCH(4, false); } \
            else { LAUNCH(8, false); }                     \
        }                                                  \
    } while (0)

}  // namespace hpcrll
