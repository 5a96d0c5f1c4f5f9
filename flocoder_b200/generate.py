"""The ``generate_samples.py`` flow around the sampling path: checkpoint -> U-Net -> ``sampler``.

Mirrors ``load_models_once`` / ``load_models`` (``generate_samples.py:61-118``), ``infer_latent_shape`` (``:120-138``) and
``generate_batch`` (``:141-159``) with the reference's argument names.  Differences, all forced by the scope of this package:

* the codec (SD-VAE / VQGAN, ``flocoder/codecs.py``) is outside the path: pass ``codec=`` (any module with ``decode`` /
  ``encode`` / ``parameters``) or ``setup_codec=`` (a callable ``(config, device) -> codec``, the reference's own
  ``setup_codec`` fits) instead of having one built from the config;
* ``config`` may be a plain nested ``dict`` (Hydra / OmegaConf are not needed): keys are read with the reference's
  ``ldcfg`` precedence (``flow`` section, then ``preencoding``, then ``codec``, then top level; ``general.py:50-75``);
* the U-Net is :class:`flocoder_b200.unet.Unet`; ``dim`` / ``channels`` come from ``init_conv.weight`` as in
  ``generate_samples.py:91-95``; ``dim_mults`` / ``n_classes`` come from the config when it names them (the reference's
  source) and otherwise from the checkpoint tensors (:func:`flocoder_b200.unet.infer_unet_config`); the reference's invalid
  ``condition=`` keyword (``:100``, a ``TypeError`` there) is not passed.
"""
from __future__ import annotations

import time
from typing import Any, Callable, Optional

import torch

from .sampling import sampler
from .unet import Unet, infer_unet_config

__all__ = ["ldcfg", "load_models_once", "load_models", "infer_latent_shape", "generate_batch"]

_codec = None
_vmodel = None
_vmodel_path = None
_config = None


def _as_dict(config) -> dict:
    if hasattr(config, "to_container"):
        return config.to_container(resolve=True)
    if isinstance(config, dict):
        return config
    try:                                               # OmegaConf without importing it
        from omegaconf import OmegaConf
        return OmegaConf.to_container(config, resolve=True)
    except Exception:
        return dict(config)


def ldcfg(config, key, default=None):
    """``general.py:50-75``: section precedence flow > preencoding > codec > top level (``default`` when absent)."""
    assert config is not None, "ldcfg: config is None, and needs to be not-None"
    cfg = _as_dict(config)
    for section in ("flow", "preencoding", "codec"):
        sec = cfg.get(section)
        if isinstance(sec, dict) and key in sec:
            return sec[key]
    return cfg.get(key, default)


def infer_latent_shape(config, debug=False):
    """``generate_samples.py:120-138``."""
    cfg = _as_dict(config)
    choice = (cfg.get("codec") or {}).get("choice")
    image_size = ldcfg(cfg, "image_size", 128)
    if choice == "sd":
        return (4, image_size // 8, image_size // 8)
    if choice == "noop":
        return (3, image_size, image_size)
    if choice == "resize":
        return (3, cfg.get("image_size", 32), cfg.get("image_size", 32))
    if choice == "vqgan":
        ds = 2 ** ldcfg(cfg, "num_downsamples", 3)
        return (ldcfg(cfg, "vq_embedding_dim", 4), image_size // ds, image_size // ds)
    raise ValueError(f"Invalid codec_choice = {choice}")


@torch.no_grad()
def load_models_once(vmodel_path, config, device=None, use_half=False, codec=None,
                     setup_codec: Optional[Callable[[Any, torch.device], torch.nn.Module]] = None,
                     compute_dtype: Optional[str] = None):
    """Load (once per ``(vmodel_path, config)``) the codec and the velocity model; returns ``(codec, vmodel)``."""
    global _codec, _vmodel, _config, _vmodel_path
    if _codec is None or _vmodel is None or _config != config or vmodel_path != _vmodel_path:
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        device = torch.device(device)
        if codec is None:
            if setup_codec is None:
                raise ValueError("the codec is outside flocoder_b200: pass codec=<module> or setup_codec=<callable(config, device)>")
            codec = setup_codec(config, device)
        new_codec = codec.to(device).eval()
        path = vmodel_path if str(vmodel_path).endswith(".pt") else f"checkpoints/{vmodel_path}.pt"
        checkpoint = torch.load(path, map_location="cpu", weights_only=False)
        state_dict = checkpoint["model_state_dict"]
        kw = infer_unet_config(state_dict)                      # dim, channels from init_conv.weight (generate_samples.py:91-95)
        cfg = _as_dict(config) if config is not None else {}
        flow = cfg.get("flow") if isinstance(cfg.get("flow"), dict) else None
        if flow is not None:
            if isinstance(flow.get("unet"), dict) and "n_classes" in flow["unet"]:
                kw["n_classes"] = int(flow["unet"]["n_classes"])
            if "dim_mults" in flow:
                kw["dim_mults"] = [int(m) for m in flow["dim_mults"]]
        vmodel = Unet(compute_dtype=compute_dtype, **kw).to(device)
        vmodel.load_state_dict(state_dict, strict=False)        # generate_samples.py:104
        _codec, _vmodel, _config, _vmodel_path = new_codec, vmodel.eval(), config, vmodel_path
    if use_half:
        return _codec.half(), _vmodel
    return _codec, _vmodel


def load_models(vmodel_path, config, device, **kwargs):
    """``generate_samples.py:116-118``."""
    return load_models_once(vmodel_path, config, device, **kwargs)


@torch.no_grad()
def generate_batch(vmodel, codec, latent_shape, method, n_steps, cfg_strength, device, curr_batch_size, is_midi=False,
                   keep_gray=False):
    """One batch through :func:`flocoder_b200.sampling.sampler` (``generate_samples.py:141-159``); returns
    ``(decoded, latents, nfe)``."""
    codec, vmodel = codec.to(device), vmodel.to(device)
    start = time.time()
    pred_latents, decoded_pred, nfe = sampler(
        model=vmodel, codec=codec, method=method, batch_size=curr_batch_size, n_steps=n_steps, cond=None, n_classes=0,
        latent_shape=tuple(latent_shape), cfg_strength=cfg_strength, is_midi=is_midi, keep_gray=keep_gray)
    generate_batch.last_seconds = time.time() - start
    return decoded_pred, pred_latents, nfe
