"""Global ``hparams`` dict, as NeuralSeq/utils/hparams.py exposes it.

The reference's hot-path modules read this dict at construction (and one default
argument at import time: shallow_diffusion_tts.py:44).  When audiogpt_b200 is
installed over the reference tree (audiogpt_b200.install()), the reference's own
``utils.hparams.hparams`` object is used instead so that both sides see the same
configuration; stand-alone, this module-level dict is the configuration.
"""
hparams = {}


def set_hparams_from_dict(d, clear=True):
    if clear:
        hparams.clear()
    hparams.update(d)
    return hparams


def resolve():
    """The dict the reference code would see, if its package is importable; else ours."""
    import sys
    mod = sys.modules.get("utils.hparams")
    if mod is not None and hasattr(mod, "hparams") and mod.__name__ != __name__:
        return mod.hparams
    return hparams
