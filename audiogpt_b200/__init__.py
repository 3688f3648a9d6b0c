"""audiogpt_b200 -- B200-native (sm_100a) back-end for AudioGPT's generative hot path.

Drop-in classes (same names / signatures / state-dict layouts as the reference):

    audiogpt_b200.modules.hifigan.hifigan.HifiGanGenerator
    audiogpt_b200.vocoders.hifigan.HifiGAN
    audiogpt_b200.modules.diff.net.DiffNet
    audiogpt_b200.modules.diff.shallow_diffusion_tts.GaussianDiffusion
    audiogpt_b200.ldm.modules.diffusionmodules.openaimodel.UNetModel
    audiogpt_b200.ldm.models.diffusion.ddim.DDIMSampler
    audiogpt_b200.ldm.models.autoencoder.AutoencoderKL          (decode side; not installed over the reference class)
    audiogpt_b200.modules.fastspeech.pe.PitchExtractor
    audiogpt_b200.vocoder.bigvgan.models.BigVGAN / VocoderBigVGAN

All arithmetic lives in libagpt_b200.so (audiogpt_b200/csrc, C ABI in include/agpt_b200.h).
There is no CPU fallback.
"""
__version__ = "0.1.0"

# reference module name -> (our module, attributes to graft onto the reference module)
_INSTALL_MAP = {
    "modules.hifigan.hifigan": ("audiogpt_b200.modules.hifigan.hifigan", ["HifiGanGenerator"]),
    "vocoders.hifigan": ("audiogpt_b200.vocoders.hifigan", ["HifiGAN", "load_model"]),
    "modules.diff.net": ("audiogpt_b200.modules.diff.net", ["DiffNet"]),
    "modules.diff.shallow_diffusion_tts": ("audiogpt_b200.modules.diff.shallow_diffusion_tts",
                                           ["GaussianDiffusion", "noise_like"]),
    "ldm.modules.diffusionmodules.openaimodel": ("audiogpt_b200.ldm.modules.diffusionmodules.openaimodel",
                                                 ["UNetModel"]),
    "ldm.models.diffusion.ddim": ("audiogpt_b200.ldm.models.diffusion.ddim", ["DDIMSampler"]),
    "vocoder.bigvgan.models": ("audiogpt_b200.vocoder.bigvgan.models", ["BigVGAN", "VocoderBigVGAN"]),
    "modules.fastspeech.pe": ("audiogpt_b200.modules.fastspeech.pe", ["PitchExtractor"]),
}


def install(strict: bool = False):
    """Make AudioGPT's tool classes pick up the B200 back-end.

    Call once, after the reference's packages are importable (``sys.path`` contains
    ``NeuralSeq/`` and ``text_to_audio/Make_An_Audio/``) and before ``audio-chatgpt.py`` builds its
    tools.  For every reference module that is importable, the hot-path classes are replaced in
    place (``setattr`` on the reference module), so ``from modules.hifigan.hifigan import
    HifiGanGenerator``, ``instantiate_from_config({'target':
    'ldm.modules.diffusionmodules.openaimodel.UNetModel', ...})`` and ``DDIMSampler(model)`` all
    resolve to the drop-ins.  Modules that are not importable are registered in ``sys.modules``
    as aliases of ours when ``strict`` is False.  Returns the list of patched names."""
    import importlib
    import sys
    patched = []
    for ref_name, (our_name, attrs) in _INSTALL_MAP.items():
        ours = importlib.import_module(our_name)
        try:
            ref = importlib.import_module(ref_name)
        except Exception:
            if strict:
                raise
            sys.modules.setdefault(ref_name, ours)
            patched.append(ref_name + " (aliased)")
            continue
        for a in attrs:
            theirs = getattr(ref, a, None)
            mine = getattr(ours, a)
            # configs the drop-in does not cover (e.g. the inpainting AttentionBlock UNet) keep the reference class
            if hasattr(mine, "_reference_cls") and isinstance(theirs, type) and theirs is not mine:
                mine._reference_cls = theirs
            setattr(ref, a, mine)
        patched.append(ref_name)
    # the vocoder registry of the reference keeps its own dict: register ours there too
    try:
        bv = importlib.import_module("vocoders.base_vocoder")
        from .vocoders.hifigan import HifiGAN
        bv.VOCODERS["hifigan"] = HifiGAN
        bv.VOCODERS["HifiGAN"] = HifiGAN
    except Exception:
        pass
    return patched
