"""PitchExtractor parity on the GPU (through the C ABI) vs golden vectors made by the reference module
(tests/golden/make_golden.py:golden_pe) and vs the CPU oracle.  Stated tolerance: rel-RMSE <= 1e-4 on pitch_pred and on
the denormalised F0 (15 stacked convs with 8 Batch/Group/LayerNorms; the same gate as a UNet / VAE forward; measured
5e-6 small, 3e-5 base); voiced / unvoiced and padding decisions must agree exactly."""
import numpy as np
import pytest
import torch

from audiogpt_b200 import specs
from audiogpt_b200.modules.fastspeech.pe import PitchExtractor
from audiogpt_b200.utils.hparams import set_hparams_from_dict
from conftest import load_golden, rel_rmse, rmse

pytestmark = pytest.mark.gpu


def build(cfg):
    set_hparams_from_dict(dict(hidden_size=cfg["hidden_size"], predictor_hidden=cfg["predictor_hidden"], ffn_padding="SAME",
                               predictor_kernel=cfg["predictor_kernel"], pitch_type="frame", use_uv=True, pitch_norm="log"))
    pe = PitchExtractor(cfg["n_mel_bins"], conv_layers=cfg["conv_layers"])
    sd = specs.synth_pe(cfg, 606)
    assert set(pe.state_dict().keys()) == set(specs.pe_param_shapes(cfg).keys())
    pe.load_state_dict(sd, strict=True)
    return pe.eval().to("cuda"), sd


@pytest.mark.parametrize("name,cfg,B,T", [("pe_small", specs.PE_SMALL, 2, 37), ("pe_base", specs.PE_BASE, 2, 150)])
def test_pe_vs_reference(name, cfg, B, T):
    g = load_golden(name)
    pe, _ = build(cfg)
    mel = specs.synth_tensor((B, T, 80), seed=71, scale=1.0, shift=-2.5)
    mel[1, -T // 5:] = 0
    r = pe(mel.cuda())
    e1, e2 = rel_rmse(r["pitch_pred"].cpu(), g["pitch_pred"]), rel_rmse(r["f0_denorm_pred"].cpu(), g["f0_denorm_pred"])
    print(f"{name}: pitch_pred rel-RMSE {e1:.2e}  f0_denorm rel-RMSE {e2:.2e}")
    assert e1 < 1e-4 and e2 < 1e-4
    assert np.array_equal(r["f0_denorm_pred"].cpu().numpy() == 0, g["f0_denorm_pred"] == 0)     # uv + padding decisions
    assert float(r["f0_denorm_pred"][1, -T // 5:].abs().max()) == 0.0


@pytest.mark.parametrize("B,T", [(1, 1), (3, 7), (1, 1000)])
def test_pe_ragged_vs_oracle(B, T):
    from oracle import pe_ref
    cfg = specs.PE_SMALL
    pe, sd = build(cfg)
    mel = specs.synth_tensor((B, T, 80), seed=300 + T, scale=1.0, shift=-2.5)
    ref = pe_ref.pe_forward(sd, cfg, mel)
    r = pe(mel.cuda())
    assert rel_rmse(r["pitch_pred"].cpu(), ref["pitch_pred"]) < 1e-4
    uv_margin = ref["pitch_pred"][:, :, 1].abs() > 1e-4          # frames whose voicing decision is not a coin flip
    got0 = (r["f0_denorm_pred"].cpu() == 0)
    assert torch.equal(got0[uv_margin], (ref["f0_denorm_pred"] == 0)[uv_margin])


def test_cpu_tensor_raises():
    pe, _ = build(specs.PE_SMALL)
    with pytest.raises(RuntimeError, match="CUDA only"):
        pe(torch.zeros(1, 4, 80))
