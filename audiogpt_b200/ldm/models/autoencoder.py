"""Drop-in for ``ldm.models.autoencoder.AutoencoderKL`` -- the DECODE side.

Reference: /root/reference/text_to_audio/Make_An_Audio/ldm/models/autoencoder.py:304-354
(``decode(z) = decoder(post_quant_conv(z))``) with ``Decoder`` from
ldm/modules/diffusionmodules/model.py:462-568.  Same constructor keywords as the reference's
``first_stage_config`` (``ddconfig``, ``lossconfig``, ``embed_dim``, ``ckpt_path``, ``ignore_keys``,
``image_key``, ``colorize_nlabels``, ``monitor``), same ``decode(z)`` signature, same state-dict keys for
what it owns (``post_quant_conv.*``, ``decoder.*``); checkpoints load with ``strict=False`` exactly like
``init_from_ckpt`` does (``encoder.*``, ``quant_conv.*``, ``loss.*`` entries are ignored: the encoder only
runs for inpainting / training and is outside the accelerated path).

Arithmetic: libagpt_b200.so (csrc/vae.cu).  CUDA only, inference only.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from ... import _lib, paramtree, specs


class AutoencoderKL(nn.Module, _lib.HandleOwner):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None):
        nn.Module.__init__(self)
        _lib.HandleOwner.__init__(self)
        dd = dict(ddconfig)
        assert dd.get("double_z", True), "AutoencoderKL needs double_z (autoencoder.py:320)"
        if dd.get("attn_type", "vanilla") != "vanilla" or dd.get("use_linear_attn", False):
            raise NotImplementedError("audiogpt_b200.AutoencoderKL supports attn_type='vanilla' only")
        if not dd.get("resamp_with_conv", True) or dd.get("tanh_out", False) or dd.get("give_pre_end", False):
            raise NotImplementedError("audiogpt_b200.AutoencoderKL: resamp_with_conv / tanh_out / give_pre_end variants")
        self.image_key = image_key
        self.embed_dim = int(embed_dim)
        self.cfg = dict(embed_dim=int(embed_dim), z_channels=int(dd["z_channels"]), resolution=int(dd["resolution"]),
                        in_channels=int(dd.get("in_channels", 1)), out_ch=int(dd["out_ch"]), ch=int(dd["ch"]),
                        ch_mult=[int(v) for v in dd["ch_mult"]], num_res_blocks=int(dd["num_res_blocks"]),
                        attn_resolutions=[int(v) for v in dd["attn_resolutions"]], dropout=0.0, double_z=True)
        self._shapes = specs.vae_decoder_param_shapes(self.cfg)
        paramtree.build(self, self._shapes)
        self._engine_sig = None
        if monitor is not None:
            self.monitor = monitor
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    # ------------------------------------------------------------------ engine
    def _cfg_struct(self):
        c = _lib.VaeCfg()
        cfg = self.cfg
        c.embed_dim, c.z_channels, c.ch, c.out_ch = cfg["embed_dim"], cfg["z_channels"], cfg["ch"], cfg["out_ch"]
        c.num_levels, c.num_res_blocks = len(cfg["ch_mult"]), cfg["num_res_blocks"]
        for i, m in enumerate(cfg["ch_mult"]):
            c.ch_mult[i] = m
            # the decoder walks the levels top-down starting at resolution / 2^(levels-1) (model.py:481,517-518)
            c.attn_at_level[i] = 1 if (cfg["resolution"] // 2 ** i) in cfg["attn_resolutions"] else 0
        return c

    def _ensure_engine(self, device):
        sig = (paramtree.params_signature(self), device.index)
        if self._h.value and sig == self._engine_sig:
            return
        self._destroy()
        _lib.require_cuda()
        arr, keep = _lib.host_weight_array([paramtree.get_param(self, k).data for k in self._shapes])
        cfg = self._cfg_struct()
        h = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().agpt_vae_create(C.byref(cfg), arr, len(keep), idx, C.byref(h)))
        self._h = h
        self._engine_sig = sig

    @torch.no_grad()
    def decode(self, z):
        """z [B, embed_dim, H, W] -> [B, out_ch, H * 2^(levels-1), W * 2^(levels-1)]  (autoencoder.py:351-354)"""
        if not z.is_cuda:
            raise RuntimeError("audiogpt_b200.AutoencoderKL runs on CUDA only (no CPU fallback)")
        self._ensure_engine(z.device)
        z = z.contiguous().float()
        B, _, H, W = z.shape
        f = 2 ** (len(self.cfg["ch_mult"]) - 1)
        out = torch.empty((B, self.cfg["out_ch"], H * f, W * f), device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().agpt_vae_decode(self._h, _lib.fptr(z), B, H, W, _lib.fptr(out), _lib.cur_stream(z.device)))
        return out

    def encode(self, x):
        raise NotImplementedError("audiogpt_b200.AutoencoderKL accelerates decode() only (the encoder is used by "
                                  "inpainting / training, outside the SURVEY.md 8 hot path)")

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("training forward is out of scope; call decode(z)")

    def get_last_layer(self):
        return paramtree.get_param(self, "decoder.conv_out.weight")
