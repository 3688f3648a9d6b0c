"""Drop-in for ``vocoders.hifigan`` (+ the registry of ``vocoders.base_vocoder``).

Reference: /root/reference/NeuralSeq/vocoders/hifigan.py:17-69 (``load_model``, ``HifiGAN.spec2wav``),
NeuralSeq/vocoders/base_vocoder.py:2-19 (``register_vocoder`` / ``get_vocoder_cls``),
NeuralSeq/vocoders/vocoder_utils.py:7-15 (spectral-subtraction ``denoise`` post-filter, out of scope:
it needs librosa STFT; requesting it raises).

``spec2wav`` takes numpy ``[T, 80]`` (optionally ``f0`` ``[T]``) and returns numpy ``[T*hop]``; it goes
through the host-buffer C-ABI entry ``agpt_hifigan_vocode_host`` (pinned H2D, generator, D2H).
"""
from __future__ import annotations

import glob
import importlib
import json
import os
import re

import numpy as np
import torch

from ..modules.hifigan.hifigan import HifiGanGenerator
from ..utils import hparams as _hp

VOCODERS = {}


def register_vocoder(cls):
    VOCODERS[cls.__name__.lower()] = cls
    VOCODERS[cls.__name__] = cls
    return cls


def get_vocoder_cls(hparams):
    name = hparams["vocoder"]
    if name in VOCODERS:
        return VOCODERS[name]
    pkg, cls_name = ".".join(name.split(".")[:-1]), name.split(".")[-1]
    return getattr(importlib.import_module(pkg), cls_name)


def _load_yaml_config(path):
    """YAML with recursive ``base_config`` inheritance (NeuralSeq/utils/hparams.py:43-63)."""
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    merged = {}
    bases = cfg.get("base_config", [])
    if isinstance(bases, str):
        bases = [bases]
    for b in bases:
        bp = os.path.normpath(os.path.join(os.path.dirname(path), b)) if b.startswith(".") else b
        if os.path.exists(bp):
            merged.update(_load_yaml_config(bp))
    merged.update({k: v for k, v in cfg.items() if k != "base_config"})
    return merged


def load_model(config_path, checkpoint_path):
    if not torch.cuda.is_available():
        raise RuntimeError("audiogpt_b200 HifiGAN needs a CUDA device (no CPU fallback)")
    device = torch.device("cuda")
    ckpt = torch.load(checkpoint_path, map_location="cpu")
    if ".yaml" in config_path:
        config = _load_yaml_config(config_path)
        state = ckpt["state_dict"]["model_gen"]
    elif ".json" in config_path:
        config = json.load(open(config_path, "r"))
        state = ckpt["generator"]
    else:
        raise ValueError(config_path)
    config.setdefault("use_pitch_embed", False)
    config.setdefault("audio_sample_rate", 22050)
    model = HifiGanGenerator(config)
    model.load_state_dict(state, strict=True)
    model.remove_weight_norm()
    model = model.eval().to(device)
    print(f"| Loaded model parameters from {checkpoint_path}.")
    print(f"| HifiGAN device: {device}.")
    return model, config, device


@register_vocoder
class HifiGAN:
    def __init__(self, model=None, config=None):
        if model is not None:                       # direct construction (tests, benchmark)
            self.model, self.config = model, config or model.h
            self.device = next(model.parameters()).device
            return
        hp = _hp.resolve()
        base_dir = hp["vocoder_ckpt"]
        config_path = f"{base_dir}/config.yaml"
        if os.path.exists(config_path):
            ckpts = glob.glob(f"{base_dir}/model_ckpt_steps_*.ckpt")
            ckpt = sorted(ckpts, key=lambda x: int(re.findall(r"model_ckpt_steps_(\d+).ckpt", x)[0]))[-1]
            print("| load HifiGAN: ", ckpt)
            self.model, self.config, self.device = load_model(config_path=config_path, checkpoint_path=ckpt)
        else:
            config_path = f"{base_dir}/config.json"
            self.model, self.config, self.device = load_model(config_path=config_path,
                                                              checkpoint_path=f"{base_dir}/generator_v1")

    def spec2wav(self, mel, **kwargs):
        hp = _hp.resolve()
        mel = np.asarray(mel, dtype=np.float32)
        f0 = kwargs.get("f0")
        if f0 is not None and hp.get("use_nsf"):
            c = torch.from_numpy(mel).unsqueeze(0).transpose(2, 1).contiguous().to(self.device)
            y = self.model(c, torch.as_tensor(np.asarray(f0, dtype=np.float32))[None, :].to(self.device)).view(-1)
            wav = y.cpu().numpy()
        else:
            wav = self.model.vocode_host(np.ascontiguousarray(mel.T)[None], device=self.device).reshape(-1)
        if hp.get("vocoder_denoise_c", 0.0) > 0:
            raise NotImplementedError("vocoder_denoise_c > 0 (librosa spectral subtraction) is outside the "
                                      "accelerated path")
        return wav
