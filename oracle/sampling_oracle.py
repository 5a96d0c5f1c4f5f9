"""CPU restatement of flocoder's fixed-step ODE integrators (test oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Follows ``flocoder/sampling.py:23-146`` (warp_time, rk4_step, v_func_cfg,
generate_latents_rk4, generate_latents) and the legacy Euler sampler
``legacy/train_sd_flowers.py:43,50-67``.  Host syncs / allocator flushes of the
reference (``sampling.py:64-67,92-94``) carry no arithmetic and are dropped.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch

Tensor = torch.Tensor


def warp_time(t, dt=None, s=.5):
    """sampling.py:23-33.  s=.5 gives 2t^3 - 3t^2 + 2t."""
    if s < 0 or s > 1.5:
        raise ValueError(f"s={s} is out of bounds.")
    tw = 4 * (1 - s) * t ** 3 + 6 * (s - 1) * t ** 2 + (3 - 2 * s) * t
    if dt:
        # operator precedence exactly as the reference writes it (sampling.py:32)
        return tw, dt * 12 * (1 - s) * t ** 2 + 12 * (s - 1) * t + (3 - 2 * s)
    return tw


def rk4_step(f: Callable, y: Tensor, t: Tensor, dt: Tensor, trace: Optional[list] = None) -> Tensor:
    """sampling.py:37-48.  ``trace`` (oracle-only) records (stage_input, stage_time, velocity)."""
    def ev(x, tt):
        k = f(x, tt)
        if trace is not None:
            trace.append((x.clone(), tt.clone(), k.clone()))
        return k
    k1 = ev(y, t)
    t_half = t + dt / 2
    k2 = ev(y + dt * k1 / 2, t_half)
    k3 = ev(y + dt * k2 / 2, t_half)
    k4 = ev(y + dt * k3, t + dt)
    return y + (dt / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


def v_func_cfg(model, cond, cfg_strength, t_vec_template, x, t, t_scale=999):
    """sampling.py:51-76 without the host syncs."""
    t_vec_template.fill_(t.item())
    t_vec = t_vec_template
    v = model(x, t_vec * t_scale, cond=cond)
    if cond and cond.get("class_cond") is not None and cfg_strength:
        cond_no_class = dict(cond)
        cond_no_class["class_cond"] = None
        v_no_class = model(x, t_vec * t_scale, cond=cond_no_class)
        v = v_no_class + cfg_strength * (v - v_no_class)
    return v


def time_grid(n_steps: int, dtype=torch.float32, init_strength: Optional[float] = None) -> Tensor:
    """sampling.py:102,109,111: linspace then ALWAYS warp_time (``if warp_time:`` is truthy)."""
    if init_strength is None:
        ts = torch.linspace(0, 1, n_steps, dtype=dtype)
    else:
        ts = torch.linspace(init_strength, 1.0, n_steps, dtype=dtype)
    return warp_time(ts)


def rk4_stage_times(ts: Tensor) -> List[Tuple[Tensor, Tensor]]:
    """(t_stage, dt) for every function evaluation, computed like sampling.py:44-47,117."""
    out = []
    for i in range(len(ts) - 1):
        t, dt = ts[i], ts[i + 1] - ts[i]
        th = t + dt / 2
        out += [(t, dt), (th, dt), (th, dt), (t + dt, dt)]
    return out


@torch.no_grad()
def generate_latents_rk4(model, shape, n_steps=50, cond=None, cfg_strength=3.0, source=None,
                         init_latents=None, init_strength=0.0, jitter_strength=0, trace=None):
    """sampling.py:79-122.  Returns (latents, nfe) with nfe = n_steps*4 (over-counts by 4, as the reference)."""
    p = next(model.parameters())
    device, dtype = p.device, p.dtype
    y = source if source is not None else torch.randn(shape, device=device, dtype=dtype)
    if init_latents is None:
        ts = torch.linspace(0, 1, n_steps, device=device, dtype=dtype)
        jitter_strength = 0
    else:
        y = (1 - init_strength) * y + init_strength * init_latents
        n_steps = max(1, int(n_steps * (1.0 - init_strength)))
        ts = torch.linspace(init_strength, 1.0, n_steps, device=device, dtype=dtype)
    ts = warp_time(ts)
    t_vec = torch.zeros(shape[0], device=device, dtype=dtype)

    def f(x, t):
        return v_func_cfg(model, cond, cfg_strength, t_vec, x, t)

    for i in range(len(ts) - 1):
        y = rk4_step(f, y, ts[i], ts[i + 1] - ts[i], trace=trace)
        # jitter (sampling.py:118-119) is only live on the init_latents branch with
        # jitter_strength>0; it draws from the global RNG and is not reproduced here.
        if jitter_strength:
            raise NotImplementedError("jitter is stochastic; not part of the parity oracle")
    return y, n_steps * 4


@torch.no_grad()
def generate_latents(model, shape, method="rk4", n_steps=50, cond=None, cfg_strength=3.0,
                     device=None, source=None, init_latents=None, init_strength=0.0):
    """sampling.py:128-146.  'rk45' is a NameError in the reference (sampling.py:142-143)."""
    if method == "rk45":
        raise NameError("name 'generate_latents_rk45' is not defined")
    return generate_latents_rk4(model, shape, n_steps, cond, cfg_strength, source=source,
                                init_latents=init_latents, init_strength=init_strength)


@torch.no_grad()
def euler_sampler(model, shape, sample_N, cond=None, source=None, eps=1e-3, trace=None):
    """legacy/train_sd_flowers.py:50-67 with the noise and cond made explicit arguments.

    dt = 1/N (python float); t_i = i/N*(1-eps)+eps; x <- x + model(x, t_i*999, cond)*dt.
    """
    p = next(model.parameters())
    x = source.clone() if source is not None else torch.randn(shape, device=p.device, dtype=p.dtype)
    dt = 1.0 / sample_N
    for i in range(sample_N):
        num_t = i / sample_N * (1 - eps) + eps
        t = torch.ones(shape[0], device=x.device, dtype=x.dtype) * num_t
        pred = model(x, t * 999, cond)
        if trace is not None:
            trace.append((x.clone(), t.clone(), pred.clone()))
        x = x + pred * dt
    return x, sample_N
