"""numpy front-end of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package ``di_hpc_b200`` never does.

Each wrapper takes numpy arrays (fp32 -> liboracle_f32.so, fp64 -> liboracle_f64.so; the dtype of
the first floating input selects the build) and returns a dict of numpy arrays.  The upstream
gradient coefficients ``coef`` play the role of ``grad_output`` of each scalar loss.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_i64 = ctypes.c_int64
c_dbl = ctypes.c_double
c_int = ctypes.c_int
c_ptr = ctypes.c_void_p


def build(force: bool = False) -> None:
    """Compile oracle.c (gcc, OpenMP) into liboracle_f32.so / liboracle_f64.so next to it."""
    src = os.path.join(_HERE, "oracle.c")
    outs = [os.path.join(_HERE, "liboracle_f32.so"), os.path.join(_HERE, "liboracle_f64.so")]
    if not force and all(os.path.exists(o) and os.path.getmtime(o) >= os.path.getmtime(src) for o in outs):
        return
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)


def _lib(dtype):
    key = "f64" if np.dtype(dtype) == np.float64 else "f32"
    if key not in _LIBS:
        path = os.path.join(_HERE, "liboracle_%s.so" % key)
        if not os.path.exists(path):
            build()
        _LIBS[key] = (ctypes.CDLL(path), key)
    return _LIBS[key]


def _call(dtype, name, *args):
    lib, key = _lib(dtype)
    fn = getattr(lib, "%s_%s" % (name, key))
    fn.restype = None
    conv = []
    for a in args:
        if a is None:
            conv.append(c_ptr(None))
        elif isinstance(a, np.ndarray):
            assert a.flags["C_CONTIGUOUS"], "oracle needs contiguous arrays"
            conv.append(c_ptr(a.ctypes.data))
        else:
            conv.append(a)
    fn(*conv)


def _f(x, dtype):
    return None if x is None else np.ascontiguousarray(x, dtype=dtype)


def _i(x):
    return np.ascontiguousarray(x, dtype=np.int64)


def _coefs(c, n):
    arr = (c_dbl * n)(*([1.0] * n if c is None else [float(v) for v in c]))
    return arr


def set_threads(n: int) -> None:
    os.environ["OMP_NUM_THREADS"] = str(n)
    for lib, _ in _LIBS.values():
        pass
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(c_int(int(n)))
    except OSError:
        pass


def place(arr):
    """Copy `arr` (2-D) into a fresh array whose pages are first-touched by the OpenMP threads that
    will later stream them (NUMA placement for the CPU-baseline timing only)."""
    arr = np.ascontiguousarray(arr)
    out = np.empty_like(arr)
    rows, B = arr.shape
    _call(arr.dtype, "orc_place_copy", out, arr, c_i64(rows), c_i64(B))
    return out


# ------------------------------------------------------------------------------------------- gae
def gae_forward(value, reward, gamma=0.99, lambda_=0.97):
    dt = value.dtype
    value, reward = _f(value, dt), _f(reward, dt)
    T, B = reward.shape
    adv = np.empty((T, B), dtype=dt)
    _call(dt, "orc_gae_forward", value, reward, adv, c_i64(T), c_i64(B), c_dbl(gamma), c_dbl(lambda_))
    return adv


def gae_backward(grad_adv, gamma=0.99, lambda_=0.97):
    dt = grad_adv.dtype
    grad_adv = _f(grad_adv, dt)
    T, B = grad_adv.shape
    gv = np.empty((T + 1, B), dtype=dt)
    gr = np.empty((T, B), dtype=dt)
    _call(dt, "orc_gae_backward", grad_adv, gv, gr, c_i64(T), c_i64(B), c_dbl(gamma), c_dbl(lambda_))
    return dict(value=gv, reward=gr)


# ------------------------------------------------------------------------------------------- td_lambda
def td_lambda(value, reward, weight=None, gamma=0.9, lambda_=0.8, coef_loss=1.0):
    dt = value.dtype
    value, reward, weight = _f(value, dt), _f(reward, dt), _f(weight, dt)
    T, B = reward.shape
    loss = np.empty((1,), dtype=dt)
    ret = np.empty((T, B), dtype=dt)
    gv = np.empty((T + 1, B), dtype=dt)
    _call(dt, "orc_td_lambda", value, reward, weight, c_i64(T), c_i64(B), c_dbl(gamma), c_dbl(lambda_),
          c_dbl(coef_loss), loss, ret, gv)
    return dict(loss=loss[0], ret=ret, grad_value=gv)


# ------------------------------------------------------------------------------------------- vtrace
def vtrace(target_output, behaviour_output, action, value, reward, weight=None, gamma=0.99, lambda_=0.95,
           rho_clip_ratio=1.0, c_clip_ratio=1.0, rho_pg_clip_ratio=1.0, coef=None, want_grad=True):
    dt = value.dtype
    target_output, behaviour_output = _f(target_output, dt), _f(behaviour_output, dt)
    value, reward, weight, action = _f(value, dt), _f(reward, dt), _f(weight, dt), _i(action)
    T, B, N = target_output.shape
    losses = np.empty((3,), dtype=dt)
    ret = np.empty((T, B), dtype=dt)
    adv = np.empty((T, B), dtype=dt)
    gt = np.empty((T, B, N), dtype=dt) if want_grad else None
    gv = np.empty((T + 1, B), dtype=dt) if want_grad else None
    _call(dt, "orc_vtrace", target_output, behaviour_output, action, value, reward, weight, c_i64(T), c_i64(B),
          c_i64(N), c_dbl(gamma), c_dbl(lambda_), c_dbl(rho_clip_ratio), c_dbl(c_clip_ratio),
          c_dbl(rho_pg_clip_ratio), _coefs(coef, 3), losses, ret, adv, gt, gv)
    return dict(policy_loss=losses[0], value_loss=losses[1], entropy_loss=losses[2], ret=ret, adv=adv,
                grad_target_output=gt, grad_value=gv)


# ------------------------------------------------------------------------------------------- upgo
def upgo(target_output, rhos, action, rewards, bootstrap_values, coef_loss=1.0, want_grad=True):
    dt = rewards.dtype
    target_output, rhos, rewards = _f(target_output, dt), _f(rhos, dt), _f(rewards, dt)
    bootstrap_values, action = _f(bootstrap_values, dt), _i(action)
    T, B, N = target_output.shape
    loss = np.empty((1,), dtype=dt)
    ret = np.empty((T, B), dtype=dt)
    gt = np.empty((T, B, N), dtype=dt) if want_grad else None
    _call(dt, "orc_upgo", target_output, rhos, action, rewards, bootstrap_values, c_i64(T), c_i64(B), c_i64(N),
          c_dbl(coef_loss), loss, ret, gt)
    return dict(loss=loss[0], ret=ret, grad_target_output=gt)


# ------------------------------------------------------------------------------------------- ppo
def ppo(logits_new, logits_old, action, value_new, value_old, adv, return_, weight=None, clip_ratio=0.2,
        use_value_clip=True, dual_clip=None, coef=None, want_grad=True):
    dt = logits_new.dtype
    logits_new, logits_old = _f(logits_new, dt), _f(logits_old, dt)
    value_new, value_old, adv, return_, weight = (_f(value_new, dt), _f(value_old, dt), _f(adv, dt),
                                                  _f(return_, dt), _f(weight, dt))
    action = _i(action)
    B, N = logits_new.shape
    out = np.empty((5,), dtype=dt)
    gl = np.empty((B, N), dtype=dt) if want_grad else None
    gv = np.empty((B,), dtype=dt) if want_grad else None
    _call(dt, "orc_ppo", logits_new, logits_old, action, value_new, value_old, adv, return_, weight, c_i64(B),
          c_i64(N), c_dbl(clip_ratio), c_int(1 if use_value_clip else 0),
          c_dbl(-1.0 if dual_clip is None else dual_clip), _coefs(coef, 3), out, gl, gv)
    return dict(policy_loss=out[0], value_loss=out[1], entropy_loss=out[2], approx_kl=out[3], clipfrac=out[4],
                grad_logits_new=gl, grad_value_new=gv)


# ------------------------------------------------------------------------------------------- gae -> normalise -> ppo
def adv_stats(adv):
    """[mean, std + 1e-8] of the advantage normalisation described at hpc_rll/origin/ppo.py:43-47
    (``(adv - adv.mean()) / (adv.std() + 1e-8)``; torch.std is the unbiased estimator).  Moments are
    taken in fp64 and rounded once to the input dtype; the epsilon is added in the input dtype."""
    dt = adv.dtype
    a = np.asarray(adv, dtype=np.float64).reshape(-1)
    n = a.size
    s1, s2 = a.sum(), np.square(a).sum()
    mean = s1 / n
    var = (s2 - s1 * mean) / (n - 1) if n > 1 else np.nan
    sd = np.sqrt(max(var, 0.0)) if var == var else np.nan
    return np.array([mean, dt.type(sd) + dt.type(1e-8)], dtype=dt)


def normalize_adv(adv, stats):
    """(adv - mean) / (std + 1e-8), one subtraction and one IEEE division per element in adv's dtype."""
    return ((adv - stats[0]) / stats[1]).astype(adv.dtype)


def gae_norm_ppo(value, reward, logits_new, logits_old, action, value_new, value_old, return_, weight=None,
                 gamma=0.99, lambda_=0.97, clip_ratio=0.2, use_value_clip=True, dual_clip=None, coef=None):
    """SURVEY.md 8(f)3 chain: origin.gae (gae.py:28-37) -> normalisation (ppo.py:43-47) -> origin.ppo_error
    (ppo.py:51-80) on the flattened (T*B,) batch."""
    adv = gae_forward(value, reward, gamma, lambda_)
    st = adv_stats(adv)
    r = ppo(logits_new, logits_old, action, value_new, value_old, normalize_adv(adv, st).reshape(-1), return_, weight,
            clip_ratio, use_value_clip, dual_clip, coef)
    r["adv"], r["adv_mean"], r["adv_denom"] = adv, st[0], st[1]
    return r


# ------------------------------------------------------------------------------------------- n-step family
def q_nstep_td(q, next_n_q, action, next_n_action, reward, done, weight=None, gamma=0.99, rescale=False,
               coef_loss=1.0, want_grad=True):
    dt = q.dtype
    q, next_n_q, reward, done, weight = _f(q, dt), _f(next_n_q, dt), _f(reward, dt), _f(done, dt), _f(weight, dt)
    action, next_n_action = _i(action), _i(next_n_action)
    T, B = reward.shape
    N = q.shape[1]
    loss = np.empty((1,), dtype=dt)
    td = np.empty((B,), dtype=dt)
    gq = np.empty((B, N), dtype=dt) if want_grad else None
    _call(dt, "orc_q_nstep_td", q, next_n_q, action, next_n_action, reward, done, weight, c_i64(T), c_i64(B),
          c_i64(N), c_dbl(gamma), c_int(1 if rescale else 0), c_dbl(coef_loss), loss, td, gq)
    return dict(loss=loss[0], td_error_per_sample=td, grad_q=gq)


def dist_nstep_td(dist, next_n_dist, action, next_n_action, reward, done, weight=None, gamma=0.99, v_min=-10.0,
                  v_max=10.0, coef_loss=1.0, want_grad=True):
    dt = dist.dtype
    dist, next_n_dist, reward, done, weight = (_f(dist, dt), _f(next_n_dist, dt), _f(reward, dt), _f(done, dt),
                                               _f(weight, dt))
    action, next_n_action = _i(action), _i(next_n_action)
    T, B = reward.shape
    _, N, n_atom = dist.shape
    loss = np.empty((1,), dtype=dt)
    td = np.empty((B,), dtype=dt)
    gd = np.empty((B, N, n_atom), dtype=dt) if want_grad else None
    _call(dt, "orc_dist_nstep_td", dist, next_n_dist, action, next_n_action, reward, done, weight, c_i64(T),
          c_i64(B), c_i64(N), c_i64(n_atom), c_dbl(gamma), c_dbl(v_min), c_dbl(v_max), c_dbl(coef_loss), loss, td,
          gd)
    return dict(loss=loss[0], td_error_per_sample=td, grad_dist=gd)


def qrdqn_nstep_td(q, next_n_q, action, next_n_action, reward, done, weight=None, value_gamma=None, gamma=0.99,
                   coef_loss=1.0, want_grad=True):
    dt = q.dtype
    q, next_n_q, reward, done = _f(q, dt), _f(next_n_q, dt), _f(reward, dt), _f(done, dt)
    weight, value_gamma = _f(weight, dt), _f(value_gamma, dt)
    action, next_n_action = _i(action), _i(next_n_action)
    T, B = reward.shape
    _, N, tau = q.shape
    loss = np.empty((1,), dtype=dt)
    td = np.empty((B,), dtype=dt)
    gq = np.empty((B, N, tau), dtype=dt) if want_grad else None
    _call(dt, "orc_qrdqn_nstep_td", q, next_n_q, action, next_n_action, reward, done, weight, value_gamma,
          c_i64(tau), c_i64(T), c_i64(B), c_i64(N), c_dbl(gamma), c_dbl(coef_loss), loss, td, gq)
    return dict(loss=loss[0], td_error_per_sample=td, grad_q=gq)


def iqn_nstep_td(q, next_n_q, action, next_n_action, reward, done, replay_quantiles, weight=None,
                 value_gamma=None, gamma=0.99, kappa=1.0, coef_loss=1.0, want_grad=True):
    dt = q.dtype
    q, next_n_q, reward, done = _f(q, dt), _f(next_n_q, dt), _f(reward, dt), _f(done, dt)
    replay_quantiles, weight, value_gamma = _f(replay_quantiles, dt), _f(weight, dt), _f(value_gamma, dt)
    action, next_n_action = _i(action), _i(next_n_action)
    T, B = reward.shape
    tau, _, N = q.shape
    tau_p = next_n_q.shape[0]
    loss = np.empty((1,), dtype=dt)
    td = np.empty((B,), dtype=dt)
    gq = np.empty((tau, B, N), dtype=dt) if want_grad else None
    _call(dt, "orc_iqn_nstep_td", q, next_n_q, action, next_n_action, reward, done, replay_quantiles, weight,
          value_gamma, c_i64(tau), c_i64(tau_p), c_i64(T), c_i64(B), c_i64(N), c_dbl(gamma), c_dbl(kappa),
          c_dbl(coef_loss), loss, td, gq)
    return dict(loss=loss[0], td_error_per_sample=td, grad_q=gq)
