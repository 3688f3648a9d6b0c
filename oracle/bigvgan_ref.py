"""CPU restatement of the BigVGAN generator (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows text_to_audio/Make_An_Audio/vocoder/bigvgan/models.py:177-203 (BigVGAN.forward),
:72-81 (AMPBlock1.forward), :120-126 (AMPBlock2.forward), activations.py:46-57 / :104-117
(Snake / SnakeBeta), alias_free_torch/act.py:22-27, resample.py:22-31 (UpSample1d.forward),
filter.py:80-90 (LowPassFilter1d.forward).  Functional: state dict in, waveform out.
Pinned against the reference's own module in tests/golden/bigvgan_small.npz.
"""
import torch
import torch.nn.functional as F


def aa_activation(x, alpha, beta, filt, logscale):
    """Activation1d(Snake|SnakeBeta): x2 up-FIR (replicate pad 5, conv_transpose stride 2, x2 gain, crop 15/15)
    -> x + 1/(b + 1e-9) sin^2(a x) -> replicate pad (5, 6) -> 12-tap low-pass, stride 2.   x: [B, C, T]"""
    C = x.shape[1]
    f = filt.reshape(1, 1, -1).expand(C, -1, -1)
    u = F.pad(x, (5, 5), mode="replicate")
    u = 2 * F.conv_transpose1d(u, f, stride=2, groups=C)
    u = u[..., 15:-15]
    a = alpha[None, :, None]
    b = (beta if beta is not None else alpha)[None, :, None]
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    u = u + (1.0 / (b + 1e-9)) * torch.pow(torch.sin(u * a), 2)
    u = F.pad(u, (5, 6), mode="replicate")
    return F.conv1d(u, f, stride=2, groups=C)


def bigvgan_forward(sd, h, mel):
    """mel [B, num_mels, T] -> wav [B, 1, T * prod(upsample_rates)]"""
    logscale = bool(h.get("snake_logscale", False))
    has_beta = str(h["activation"]) == "snakebeta"

    def act(prefix, x):
        return aa_activation(x, sd[prefix + ".act.alpha"], sd[prefix + ".act.beta"] if has_beta else None,
                             sd[prefix + ".upsample.filter"], logscale)

    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (ks, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}"
            y = x
            if str(h["resblock"]) == "1":
                for n, d in enumerate(dil):
                    xt = act(f"{p}.activations.{2 * n}", y)
                    xt = F.conv1d(xt, sd[f"{p}.convs1.{n}.weight"], sd[f"{p}.convs1.{n}.bias"], dilation=d,
                                  padding=(ks * d - d) // 2)
                    xt = act(f"{p}.activations.{2 * n + 1}", xt)
                    xt = F.conv1d(xt, sd[f"{p}.convs2.{n}.weight"], sd[f"{p}.convs2.{n}.bias"], padding=(ks - 1) // 2)
                    y = xt + y
            else:
                for n, d in enumerate(dil):
                    xt = act(f"{p}.activations.{n}", y)
                    xt = F.conv1d(xt, sd[f"{p}.convs.{n}.weight"], sd[f"{p}.convs.{n}.bias"], dilation=d,
                                  padding=(ks * d - d) // 2)
                    y = xt + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = act("activation_post", x)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)
