"""Checkpoint / config I/O in the on-disk layout the reference uses, without depending on diffusers.

The reference class inherits diffusers ModelMixin + ConfigMixin (compressive_vq_model.py:9-11,33-36) and is
loaded with `CompressiveVQModel.from_pretrained(path, subfolder='tokenizer', low_cpu_mem_usage=False)`
(inference/predict.py:94-95, train_gpt.py:136-139), `from_config(path)` (mbrl/video_predictor.py:46) and saved
with `save_pretrained(dir)` (train_tokenizer.py:96-101).  Layout: <dir>/config.json +
<dir>/diffusion_pytorch_model.safetensors (fallback diffusion_pytorch_model.bin).
"""
from __future__ import annotations

import inspect
import json
import os
from typing import Any, Dict, Optional

import torch

WEIGHTS_SAFE = "diffusion_pytorch_model.safetensors"
WEIGHTS_BIN = "diffusion_pytorch_model.bin"


class ConfigDict(dict):
    """dict with attribute access (mirrors how callers read diffusers' FrozenDict: cfg.x and cfg['x'])."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class HubMixin:
    config_name = "config.json"

    def register_config(self, init_locals: Dict[str, Any]):
        sig = inspect.signature(type(self).__init__)
        cfg = ConfigDict()
        for name in sig.parameters:
            if name == "self":
                continue
            v = init_locals[name]
            cfg[name] = list(v) if isinstance(v, tuple) else v
        cfg["_class_name"] = type(self).__name__
        object.__setattr__(self, "_config", cfg)

    @property
    def config(self) -> ConfigDict:
        return self._config

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def _resolve(cls, path: str, subfolder: Optional[str]) -> str:
        d = os.path.join(path, subfolder) if subfolder else path
        if os.path.isfile(d):  # a config file was given directly
            return d
        if not os.path.isdir(d):
            raise FileNotFoundError(f"{d} is not a directory (offline build: hub downloads are not supported)")
        return d

    @classmethod
    def load_config(cls, path: str, subfolder: Optional[str] = None) -> Dict[str, Any]:
        d = cls._resolve(path, subfolder)
        cfg_file = d if os.path.isfile(d) else os.path.join(d, cls.config_name)
        with open(cfg_file) as fh:
            raw = json.load(fh)
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
        return {k: (tuple(v) if isinstance(v, list) else v) for k, v in raw.items() if k in accepted}

    @classmethod
    def from_config(cls, config, **overrides):
        if isinstance(config, (str, os.PathLike)):
            config = cls.load_config(str(config))
        else:
            accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
            config = {k: (tuple(v) if isinstance(v, list) else v) for k, v in dict(config).items() if k in accepted}
        config.update(overrides)
        return cls(**config)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, revision=None,
                        variant=None, use_safetensor=True, use_safetensors=None, low_cpu_mem_usage=False,
                        device_map=None, ignore_mismatched_sizes=False, torch_dtype=None, **unused):
        d = cls._resolve(str(pretrained_model_name_or_path), subfolder)
        model = cls(**cls.load_config(d))
        safe = os.path.join(d, WEIGHTS_SAFE)
        if os.path.exists(safe):
            from safetensors.torch import load_file
            state = load_file(safe)
        elif os.path.exists(os.path.join(d, WEIGHTS_BIN)):
            state = torch.load(os.path.join(d, WEIGHTS_BIN), map_location="cpu")
        else:
            raise FileNotFoundError(f"no {WEIGHTS_SAFE} / {WEIGHTS_BIN} under {d}")
        state = _convert_legacy_attention_keys(state)
        if ignore_mismatched_sizes:
            own = model.state_dict()
            state = {k: v for k, v in state.items() if k in own and own[k].shape == v.shape}
            model.load_state_dict(state, strict=False)
        else:
            model.load_state_dict(state, strict=True)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        model.eval()
        return model

    def save_pretrained(self, save_directory, is_main_process: bool = True, save_function=None,
                        safe_serialization: bool = True, variant=None, state_dict=None, **unused):
        if not is_main_process:
            return
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w") as fh:
            json.dump(dict(self.config), fh, indent=2, sort_keys=True)
        state = state_dict if state_dict is not None else self.state_dict()
        state = {k: v.detach().cpu().contiguous() for k, v in state.items()}
        if save_function is not None:
            save_function(state, os.path.join(save_directory, WEIGHTS_BIN))
        elif safe_serialization:
            from safetensors.torch import save_file
            save_file(state, os.path.join(save_directory, WEIGHTS_SAFE), metadata={"format": "pt"})
        else:
            torch.save(state, os.path.join(save_directory, WEIGHTS_BIN))


def _convert_legacy_attention_keys(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Old diffusers checkpoints name mid-block attention weights query/key/value/proj_attn; newer ones
    to_q/to_k/to_v/to_out.0 (diffusers converts on load; so do we)."""
    ren = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}
    out = {}
    for k, v in state.items():
        if ".attentions." in k:
            for a, b in ren.items():
                k = k.replace(a, b)
        out[k] = v
    return out
