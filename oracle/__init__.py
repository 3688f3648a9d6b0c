"""CPU oracle for the AudioGPT generative hot path -- TEST INFRASTRUCTURE ONLY.

This package is a from-scratch, functional (state-dict in, tensor out) fp32
restatement of the reference's arithmetic for the path named in
BASELINE.json:north_star:

    HifiGanGenerator.forward        NeuralSeq/modules/hifigan/hifigan.py:144-169
    DiffNet.forward                 NeuralSeq/modules/diff/net.py:107-130
    GaussianDiffusion.p_sample/...  NeuralSeq/modules/diff/shallow_diffusion_tts.py:134-283
    UNetModel.forward               text_to_audio/Make_An_Audio/ldm/modules/diffusionmodules/openaimodel.py:711-744
    DDIMSampler.sample              text_to_audio/Make_An_Audio/ldm/models/diffusion/ddim.py:58-225

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
/ ``--impl reference`` legs may import it.  The product package
(``audiogpt_b200``) never does: it has no CPU fallback and fails loudly when
its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures for
this path, so the oracle is pinned against outputs of the
reference's own modules executed in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``); see
``tests/test_oracle_golden.py``.
"""
