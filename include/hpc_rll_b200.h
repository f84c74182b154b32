/*
 * hpc_rll_b200.h -- C ABI of the B200-native trajectory-return library (libhpc_rll_b200.so).
 *
 * Drop-in boundary for DI-hpc's `hpc_rl_utils` native extension
 * (/root/reference/src/rl_utils/entry.cpp:8-39, declarations in
 * /root/reference/include/hpc/rll/cuda/rl_utils/entry.h:62-165).  The reference boundary is C++
 * (`std::vector<torch::Tensor>` in/out); here every entry point is plain C: device pointers, sizes,
 * scalars and a CUDA stream handle -- no torch types -- so it can be bound from ctypes / pybind /
 * cgo alike.  INTEGRATION.md shows the binding used by `hpc_rll.rl_utils`.
 *
 * Conventions
 *   - all tensors are fp32, row-major, resident on the CURRENT CUDA device; `action` tensors int64
 *   - (T,B) means T rows of B contiguous columns (time-major, as in the reference)
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - every function returns HPC_RLL_OK (0) or an error code; hpc_rll_last_error() gives the text
 *   - `global_B`: the batch may be one shard of a data-parallel global batch; mean-type losses and
 *     their gradients are normalised by the GLOBAL element count (0 means "B is the whole batch").
 *     Per-rank losses then add up (one all-reduce(SUM)) to the global loss.
 *   - gamma/lambda/... are doubles: the reference's Python layer computes e.g. gamma*lambda and
 *     gamma**nstep in double before the value meets an fp32 tensor (hpc_rll/origin/gae.py:30)
 *   - `workspace`: caller-provided device scratch of at least hpc_rll_workspace_bytes(op, ...) bytes
 *   - nothing here synchronises the device or the stream; no CPU fallback exists.
 */
#ifndef HPC_RLL_B200_H_
#define HPC_RLL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPC_RLL_OK 0
#define HPC_RLL_EINVAL 1 /* bad argument (null pointer, negative size, misaligned ...) */
#define HPC_RLL_ECUDA 2  /* a CUDA runtime/driver call failed */
#define HPC_RLL_ENOSUP 3 /* unsupported configuration */

/* op ids for hpc_rll_workspace_bytes / hpc_rll_debug_set_config */
#define HPC_RLL_OP_GAE 0
#define HPC_RLL_OP_TD_LAMBDA 1
#define HPC_RLL_OP_VTRACE 2
#define HPC_RLL_OP_UPGO 3
#define HPC_RLL_OP_PPO 4
#define HPC_RLL_OP_Q_NSTEP_TD 5
#define HPC_RLL_OP_DIST_NSTEP_TD 6
#define HPC_RLL_OP_QRDQN_NSTEP_TD 7
#define HPC_RLL_OP_IQN_NSTEP_TD 8
#define HPC_RLL_OP_COUNT 9

/* ---- library ------------------------------------------------------------------------------ */
const char* hpc_rll_version(void);
const char* hpc_rll_last_error(void); /* thread-local, valid until the next failing call */
/* number of CUDA kernels this library has launched in this process (bench.py: "gpu_launches") */
uint64_t hpc_rll_launch_count(void);
/* scratch bytes an op needs for sizes (T,B,N); N = action dim / n_atom / tau as relevant */
size_t hpc_rll_workspace_bytes(int op, int64_t T, int64_t B, int64_t N);
/* tuning/debug: force kernel configuration `cfg` for `op` (-1 = automatic) */
int hpc_rll_debug_set_config(int op, int cfg);

/* ---- GAE ------------------------------------------------------------------------------------
 * replaces GaeForward (/root/reference/src/rl_utils/gae.cu:8-28, kernel
 * include/hpc/rll/cuda/rl_utils/gae_kernel.h:10-29); semantics of hpc_rll/origin/gae.py:28-37.
 *   value (T+1,B), reward (T,B) -> adv (T,B).
 * hpc_rll_gae_backward is the adjoint the reference lacks (GAEFunction.backward returns None,
 * hpc_rll/rl_utils/gae.py:16-18):  grad_adv (T,B) -> grad_value (T+1,B), grad_reward (T,B).
 * The *_ld variants take row pitches in elements (>= B) for strided views. */
int hpc_rll_gae_forward(const float* value, const float* reward, float* adv, int64_t T, int64_t B, double gamma,
                        double lambda, void* stream);
int hpc_rll_gae_backward(const float* grad_adv, float* grad_value, float* grad_reward, int64_t T, int64_t B,
                         double gamma, double lambda, void* stream);
int hpc_rll_gae_forward_ld(const float* value, int64_t ld_value, const float* reward, int64_t ld_reward,
                           float* adv, int64_t ld_adv, int64_t T, int64_t B, double gamma, double lambda,
                           void* stream);
int hpc_rll_gae_backward_ld(const float* grad_adv, int64_t ld_grad_adv, float* grad_value, int64_t ld_grad_value,
                            float* grad_reward, int64_t ld_grad_reward, int64_t T, int64_t B, double gamma,
                            double lambda, void* stream);
/* HOST-buffer entry (end-to-end path): pinned or pageable host arrays in, host arrays out.  Runs
 * forward and backward on the current device, pipelining column blocks over H2D copy / kernels /
 * D2H copy on internal streams; returns after all results have landed in the host buffers.
 * Any of h_grad_adv/h_grad_value/h_grad_reward may be NULL together (forward only). */
int hpc_rll_gae_fwd_bwd_host(const float* h_value, const float* h_reward, const float* h_grad_adv, float* h_adv,
                             float* h_grad_value, float* h_grad_reward, int64_t T, int64_t B, double gamma,
                             double lambda);

#ifdef __cplusplus
}
#endif
#endif /* HPC_RLL_B200_H_ */
