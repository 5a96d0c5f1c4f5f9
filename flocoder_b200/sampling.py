"""Drop-in for the integrator half of ``flocoder/sampling.py`` (lines 23-146).

Same function names, argument order, defaults and return values as the reference:
``warp_time``, ``rk4_step``, ``v_func_cfg``, ``generate_latents_rk4``, ``generate_latents``
(+ the legacy ``euler_sampler``, ``legacy/train_sd_flowers.py:50-67``).

When ``model`` is a :class:`flocoder_b200.unet.Unet` the whole trajectory runs inside the
C-ABI call ``flo_integrate``: stage times, the four U-Net evaluations per interval, the
classifier-free-guidance combine and the RK4 axpy/state update are all device-side (the RK
combination is the epilogue of the last convolution), with no per-evaluation host sync.  The
reference's per-evaluation ``.item()`` / ``empty_cache()`` / ``synchronize()`` calls
(``sampling.py:60,64-67``) carry no arithmetic and are not reproduced.

For any other callable ``model`` the functions keep the reference's generic semantics
(plain tensor algebra on whatever device the tensors live on) so ``rk4_step`` and
``v_func_cfg`` stay usable as building blocks; that generic composition is reference
behaviour for foreign models, not a fallback for ours.
"""
from __future__ import annotations

import random
from functools import partial

import torch

from . import _lib
from .unet import Unet

__all__ = ["warp_time", "rk4_step", "v_func_cfg", "generate_latents_rk4", "generate_latents",
           "euler_sampler", "time_grid"]


def warp_time(t, dt=None, s=.5):
    """Parametric time warp, ``sampling.py:23-33`` (s=.5 -> 2t^3-3t^2+2t)."""
    if s < 0 or s > 1.5:
        raise ValueError(f"s={s} is out of bounds.")
    tw = 4 * (1 - s) * t ** 3 + 6 * (s - 1) * t ** 2 + (3 - 2 * s) * t
    if dt:   # same operator precedence as the reference's derivative branch (sampling.py:32)
        return tw, dt * 12 * (1 - s) * t ** 2 + 12 * (s - 1) * t + (3 - 2 * s)
    return tw


@torch.no_grad()
def rk4_step(f, y, t, dt, debug=False):
    """One classic RK4 step for an arbitrary velocity function ``f(y, t)`` (``sampling.py:37-48``)."""
    k1 = f(y, t)
    t_mid = t + dt / 2
    k2 = f(y + dt * k1 / 2, t_mid)
    k3 = f(y + dt * k2 / 2, t_mid)
    k4 = f(y + dt * k3, t + dt)
    return y + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


@torch.no_grad()
def v_func_cfg(model, cond, cfg_strength, t_vec_template, x, t, t_scale=999, debug=False):
    """Velocity with optional classifier-free guidance (``sampling.py:51-76``)."""
    t_vec_template.fill_(float(t))
    t_vec = t_vec_template
    v = model(x, t_vec * t_scale, cond=cond)
    if cond and cond.get("class_cond") is not None and cfg_strength:
        cond_no_class = dict(cond)
        cond_no_class["class_cond"] = None
        v_no_class = model(x, t_vec * t_scale, cond=cond_no_class)
        v = v_no_class + cfg_strength * (v - v_no_class)
    return v


def time_grid(n_steps, start=0.0, device=None, dtype=torch.float32):
    """The grid ``generate_latents_rk4`` integrates over: linspace then ALWAYS ``warp_time``
    (``sampling.py:102,109,111`` -- ``if warp_time:`` tests the function object)."""
    return warp_time(torch.linspace(start, 1.0, n_steps, device=device, dtype=dtype))


def _device_dtype(model):
    p = next(model.parameters())
    return p.device, p.dtype


def _class_ids(model, cond):
    cls = Unet.split_cond(cond)
    if cls is not None and not model.class_condition:
        cls = None
    return cls


@torch.no_grad()
def generate_latents_rk4(model, shape, n_steps=50, cond=None, cfg_strength=3.0, source=None,
                         init_latents=None, init_strength=0.0, jitter_strength=0, debug=False):
    """RK4 sampling over the warped grid (``sampling.py:79-122``).  Returns ``(latents, nfe)``
    with ``nfe = n_steps*4`` exactly as the reference reports it (it over-counts by 4)."""
    device, dtype = _device_dtype(model)
    y = source if source is not None else torch.randn(shape, device=device, dtype=dtype)
    if init_latents is None:
        ts = torch.linspace(0, 1, n_steps, device="cpu", dtype=torch.float32)
        jitter_strength = 0
    else:
        y = (1 - init_strength) * y + init_strength * init_latents
        n_steps = max(1, int(n_steps * (1.0 - init_strength)))
        ts = torch.linspace(init_strength, 1.0, n_steps, device="cpu", dtype=torch.float32)
    ts = warp_time(ts)
    nfe = n_steps * 4

    if not isinstance(model, Unet):
        # foreign model: the reference's generic composition (sampling.py:112-119)
        ts_d = ts.to(device=device, dtype=dtype)
        t_vec = torch.zeros(shape[0], device=device, dtype=dtype)
        v_func = partial(v_func_cfg, model, cond, cfg_strength, t_vec)
        for i in range(len(ts_d) - 1):
            y = rk4_step(v_func, y, ts_d[i], ts_d[i + 1] - ts_d[i])
            if random.random() < 0.1 and jitter_strength > 0:
                y += torch.randn_like(y) * jitter_strength * (1 - ts_d[i])
        return y, nfe

    # ---- B200 fast path: the whole trajectory is one C-ABI call ----
    b, c, h, w = y.shape
    eng = model.engine(h, w)
    out_dtype = y.dtype
    state = y.to(device=eng.device, dtype=torch.float32).contiguous()
    if state.data_ptr() == y.data_ptr():
        state = state.clone()                      # never integrate in place on the caller's tensor
    cls = _class_ids(model, cond)
    cfg = float(cfg_strength) if (cls is not None and cfg_strength) else 0.0
    grid = ts.tolist()
    if len(grid) >= 2:
        if jitter_strength > 0:
            # stochastic jitter needs a host decision per interval (sampling.py:118-119)
            for i in range(len(grid) - 1):
                eng.integrate(state, grid[i:i + 2], _lib.FLO_RK4, class_ids=cls, cfg_strength=cfg)
                if random.random() < 0.1:
                    state += torch.randn_like(state) * jitter_strength * (1 - grid[i])
        else:
            eng.integrate(state, grid, _lib.FLO_RK4, class_ids=cls, cfg_strength=cfg)
    return state.to(out_dtype), nfe


@torch.no_grad()
def generate_latents(model, shape, method="rk4", n_steps=50, cond=None, cfg_strength=3.0, device=None,
                     source=None, init_latents=None, init_strength=0.0, debug=False):
    """Dispatch (``sampling.py:128-146``).  ``method='rk45'`` is a NameError in the reference
    (``generate_latents_rk45`` was removed, ``sampling.py:142-143``); here it is an explicit error."""
    if device is None:
        device = next(model.parameters()).device
    if method == "rk45":
        raise NotImplementedError("method='rk45' was removed from the reference (sampling.py:142-143 calls an "
                                  "undefined function); use method='rk4'")
    return generate_latents_rk4(model, shape, n_steps, cond, cfg_strength, source=source,
                                init_latents=init_latents, init_strength=init_strength)


@torch.no_grad()
def euler_sampler(model, shape, sample_N, device=None, cond=None, source=None, eps=1e-3):
    """Legacy fixed-step Euler (``legacy/train_sd_flowers.py:43,50-67``): ``dt = 1/N``,
    ``t_i = i/N*(1-eps)+eps``, ``x <- x + model(x, t_i*999, cond)*dt``.  Returns ``(x, nfe=N)``.

    ``cond`` and ``source`` are explicit here (the legacy script drew both from globals); like
    the legacy function the result is returned on the CPU."""
    dev, dtype = _device_dtype(model)
    x = source if source is not None else torch.randn(shape, device=dev, dtype=dtype)
    dt = 1.0 / sample_N
    times = [i / sample_N * (1 - eps) + eps for i in range(sample_N)]
    if not isinstance(model, Unet):
        x = x.detach().clone()
        for num_t in times:
            t = torch.ones(shape[0], device=x.device, dtype=x.dtype) * num_t
            x = x.detach().clone() + model(x, t * 999, cond) * dt
        return x.cpu(), sample_N
    b, c, h, w = x.shape
    eng = model.engine(h, w)
    out_dtype = x.dtype
    state = x.to(device=eng.device, dtype=torch.float32).contiguous().clone()
    cls = _class_ids(model, cond)
    # the legacy loop feeds fl32(num_t) (ones*num_t in fp32) and multiplies the velocity by fl32(dt)
    ts32 = torch.tensor(times, dtype=torch.float64).to(torch.float32).tolist()
    eng.integrate(state, ts32, _lib.FLO_EULER_LEGACY, dt=dt, class_ids=cls, cfg_strength=0.0)
    return state.to(out_dtype).cpu(), sample_N
