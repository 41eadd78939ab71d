"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's StyleGAN2 hot path.

This file is the *oracle* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or
the CPU baseline -- never as (part of) the product path.

It restates, in plain functional torch (fp32 by default, fp64 when the tensors passed in are fp64), the
arithmetic of

  * model/stylegan2_common_layers.py  (upfirdn2d :42-72, make_kernel :83-91, Upsample :94-112,
    Blur :136-152, EqualConv2d :155-184, EqualLinear :193-230, ModulatedConv2d :250-349,
    NoiseInjection :388-431, StyledConv :447-486, ToRGB :489-511, get_w_frm_z :514-533,
    ConvLayer :752-799, ResBlock :802-820, FusedLeakyReLU :22-39, PixelNorm :75-80)
  * model/stg2_generator.py           (StyledGenerator.forward :249-328, Generator.forward :159-209)
  * model/stg2_discriminator.py       (Discriminator.forward :48-76)
  * loss_functions/losses.py          (grad_penalty_loss :87-99, PathLengthRegularizor :102-124)

The functions take the reference's ``state_dict`` (same key names / shapes) so the same weights can be fed
to the reference modules, to this oracle and to the CUDA path.  The formulation is deliberately *not* the
reference's (no per-sample weight materialisation / grouped convolutions): it follows the closed forms of
SURVEY.md Appendix A, and is pinned against the unmodified reference modules by ``oracle/make_golden.py``
(run in the build container; results committed under ``tests/golden/``) and ``tests/test_oracle_golden.py``.

Parity status: pinned by live execution of the reference modules (the reference ships no golden vectors
for these ops -- SURVEY.md §8c).  PPL (path-length regulariser) is *parity unpinned*: the reference's code
is unrunnable (losses.py:102-124); ``path_length_penalty`` implements the adopted rule.
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------- A1: upfirdn2d
def make_kernel(k):
    """cl.py:83-91 -- outer product of a 1-D tap list, normalised to unit sum."""
    k = torch.as_tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def upfirdn2d(x, k, up=1, down=1, pad=(0, 0)):
    """cl.py:42-72.  out[y,x] = sum_{a,b} k[a,b] * U[y*down + (kh-1-a) - p0, x*down + (kw-1-b) - p0]
    where U is x with (up-1) zeros inserted after every sample.  Same pad on both axes (cl.py:53);
    negative pads crop (cl.py:59-60)."""
    n, c, h, w = x.shape
    kh, kw = k.shape
    p0, p1 = pad
    u = x.new_zeros(n, c, h * up, w * up)
    u[:, :, ::up, ::up] = x
    u = F.pad(u, (max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)))
    u = u[:, :, max(-p0, 0): u.shape[2] - max(-p1, 0), max(-p0, 0): u.shape[3] - max(-p1, 0)]
    ho = u.shape[2] - kh + 1
    wo = u.shape[3] - kw + 1
    out = x.new_zeros(n, c, ho, wo)
    for a in range(kh):
        for b in range(kw):
            out = out + k[a, b].to(x.dtype) * u[:, :, kh - 1 - a: kh - 1 - a + ho, kw - 1 - b: kw - 1 - b + wo]
    return out[:, :, ::down, ::down]


def upsample2(x, k4):
    """Upsample(kernel, factor=2) cl.py:94-112: kernel*4, pad (2,1)."""
    return upfirdn2d(x, k4, up=2, down=1, pad=(2, 1))


# ----------------------------------------------------------------------------- elementwise
def fused_leaky_relu(x, bias, slope=0.2, scale=SQRT2):
    """FusedLeakyReLU cl.py:22-39; bias is (1,C,1,1)."""
    return scale * F.leaky_relu(x + bias, slope)


def pixel_norm(x):
    """PixelNorm cl.py:75-80."""
    return x * torch.rsqrt((x * x).mean(dim=1, keepdim=True) + 1e-8)


def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """EqualLinear cl.py:212-230 (apply_sqrt2_fac_in_eq_lin is False on every activated layer)."""
    y = x @ (weight * (lr_mul / math.sqrt(weight.shape[1]))).t()
    if bias is not None:
        y = y + bias * lr_mul
    if activation:
        y = F.leaky_relu(y, 0.2)
    return y


def equal_conv2d(x, weight, bias=None, stride=1, padding=0):
    """EqualConv2d cl.py:155-184."""
    co, ci, k, _ = weight.shape
    return F.conv2d(x, weight * (1.0 / math.sqrt(ci * k * k)), bias, stride=stride, padding=padding)


# ----------------------------------------------------------------------------- A3: modulated conv
def modulated_conv2d(x, style, weight, mod_weight, mod_bias, demodulate=True, upsample=False, blur_kernel=None):
    """ModulatedConv2d.forward cl.py:307-349 in the modulate-input / shared-weight / demodulate-output form
    (SURVEY Appendix A3):  y = d[b,o] * conv(W~, s[b,i] * x[b,i]).
    ``weight`` is (1,Co,Ci,k,k); ``blur_kernel`` is the (already x4) 4x4 buffer of the upsample branch."""
    _, co, ci, k, _ = weight.shape
    s = equal_linear(style, mod_weight, mod_bias)                       # (B,Ci)   cl.py:311
    wt = weight[0] * (1.0 / math.sqrt(ci * k * k))                       # W~       cl.py:288-289,312
    xs = x * s[:, :, None, None]
    if upsample:
        acc = F.conv_transpose2d(xs, wt.transpose(0, 1), stride=2, padding=0)   # (2H+1)^2  cl.py:322-331
        acc = upfirdn2d(acc, blur_kernel, pad=(1, 1))                            # cl.py:272-278,333
    else:
        acc = F.conv2d(xs, wt, padding=k // 2)                                   # cl.py:343-347
    if demodulate:
        q = (wt * wt).sum(dim=(2, 3))                                            # (Co,Ci)
        d = torch.rsqrt((s * s) @ q.t() + 1e-8)                                  # cl.py:315-316
        acc = acc * d[:, :, None, None]
    return acc


def noise_injection(image, cond, sd, prefix):
    """NoiseInjection.forward cl.py:421-431: three plain 3x3 convs (+ReLU) on the condition map."""
    h = F.relu(F.conv2d(cond, sd[prefix + "noise_conv.0.weight"], sd[prefix + "noise_conv.0.bias"], padding=1))
    h = F.relu(F.conv2d(h, sd[prefix + "noise_conv.2.weight"], sd[prefix + "noise_conv.2.bias"], padding=1))
    h = F.conv2d(h, sd[prefix + "noise_conv.4.weight"], sd[prefix + "noise_conv.4.bias"], padding=1)
    return image + h


def styled_conv(x, style, cond, sd, prefix, upsample):
    """StyledConv.forward cl.py:479-486 = modconv -> noise(cond) -> FusedLeakyReLU."""
    y = modulated_conv2d(x, style, sd[prefix + "conv.weight"], sd[prefix + "conv.modulation.weight"],
                         sd[prefix + "conv.modulation.bias"], True, upsample,
                         sd.get(prefix + "conv.blur.kernel"))
    y = noise_injection(y, cond, sd, prefix + "noise.")
    return fused_leaky_relu(y, sd[prefix + "activate.bias"])


def to_rgb(x, style, skip, sd, prefix):
    """ToRGB.forward cl.py:502-511: 1x1 modconv without demod + bias + Upsample(skip)."""
    y = modulated_conv2d(x, style, sd[prefix + "conv.weight"], sd[prefix + "conv.modulation.weight"],
                         sd[prefix + "conv.modulation.bias"], demodulate=False)
    y = y + sd[prefix + "bias"]
    if skip is not None:
        y = y + upsample2(skip, sd[prefix + "upsample.kernel"])
    return y


# ----------------------------------------------------------------------------- M1: generator
def cond_pyramid_level(cond, size):
    """gen.py:309-314: F.interpolate(bilinear, align_corners=False) by an integer power-of-two factor equals
    the mean of the central 2x2 of every s x s block."""
    s = cond.shape[-1] // size
    if s == 1:
        return cond
    a = s // 2 - 1
    return 0.25 * (cond[:, :, a::s, a::s] + cond[:, :, a::s, a + 1::s] + cond[:, :, a + 1::s, a::s]
                   + cond[:, :, a + 1::s, a + 1::s])


def mapping_network(z, sd, n_mlp=8, prefix="z_to_w."):
    """get_w_frm_z cl.py:514-524 with lr_mlp=0.01 (gen.py:237): PixelNorm then n_mlp activated EqualLinear."""
    h = pixel_norm(z)
    for i in range(1, n_mlp + 1):
        h = equal_linear(h, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"], lr_mul=0.01, activation=True)
    return h


def synthesis(w, cond, sd, step=6, prefix="generator."):
    """Generator.forward gen.py:159-209 for a single style (inject_index beyond range, gen.py:166-167),
    core_tensor_res=4: blocks 0..step, condition injected as 'noise' at every resolution."""
    b = w.shape[0]
    x = sd[prefix + "const_input.input"].repeat(b, 1, 1, 1)              # gen.py:196
    rgb = None
    for i in range(step + 1):
        c = cond_pyramid_level(cond, 4 * 2 ** i)
        p = f"{prefix}progression.{i}."
        if i == 0:
            x = styled_conv(x, w, c, sd, p + "st_cv1.", upsample=False)  # one_conv_block gen.py:62-63
        else:
            x = styled_conv(x, w, c, sd, p + "st_cv1.", upsample=True)
            x = styled_conv(x, w, c, sd, p + "st_cv2.", upsample=False)
        rgb = to_rgb(x, w, rgb, sd, f"{prefix}to_rgb.{i}.")              # gen.py:201-207
    return rgb


def generator_forward(cond, input_indices, sd, step=6, n_mlp=8):
    """StyledGenerator.forward gen.py:249-328 in rendered-condition mode: integer indices select the frozen
    identity embedding (gen.py:34-46,275); float (B,512) tensors are fed to z_to_w directly (gen.py:272-273)."""
    if input_indices.dtype in (torch.float32, torch.float64):
        z = input_indices
    else:
        z = sd["image_embedding.embd_weight"][input_indices]
    w = mapping_network(z, sd, n_mlp)
    return synthesis(w, cond, sd, step)


# ----------------------------------------------------------------------------- M2: discriminator
def conv_layer(x, sd, prefix, k, downsample, activate=True, bias=True):
    """ConvLayer cl.py:752-799: [Blur] -> EqualConv2d -> FusedLeakyReLU."""
    idx = 0
    if downsample:
        p = 2 + (k - 1)                                                   # (len(blur)-factor) + (k-1)
        x = upfirdn2d(x, sd[f"{prefix}0.kernel"], pad=((p + 1) // 2, p // 2))
        idx = 1
    x = equal_conv2d(x, sd[f"{prefix}{idx}.weight"], None, stride=2 if downsample else 1,
                     padding=0 if downsample else k // 2)
    if activate and bias:
        x = fused_leaky_relu(x, sd[f"{prefix}{idx + 1}.bias"])
    return x


def res_block(x, sd, prefix):
    """ResBlock.forward cl.py:813-820."""
    y = conv_layer(x, sd, prefix + "conv1.", 3, False)
    y = conv_layer(y, sd, prefix + "conv2.", 3, True)
    s = conv_layer(x, sd, prefix + "skip.", 1, True, activate=False, bias=False)
    return (y + s) / SQRT2


def minibatch_stddev(x, group=4):
    """disc.py:59-65: group = min(B,4); std over the group axis of view(group, B/group, 1, C, H, W)."""
    b, c, h, w = x.shape
    g = min(b, group)
    v = x.view(g, b // g, 1, c, h, w)
    sdv = torch.sqrt(v.var(0, unbiased=False) + 1e-8)
    m = sdv.mean(dim=(2, 3, 4), keepdim=True).squeeze(2)                  # (B/g,1,1,1)
    return torch.cat([x, m.repeat(g, 1, h, w)], dim=1)


def discriminator_forward(img, cond, sd, size=256):
    """Discriminator.forward disc.py:48-76."""
    x = torch.cat([img, cond], dim=1) if cond is not None else img
    x = conv_layer(x, sd, "convs.0.", 1, False)
    n_blocks = int(math.log2(size)) - 2
    for j in range(1, n_blocks + 1):
        x = res_block(x, sd, f"convs.{j}.")
    x = minibatch_stddev(x)
    x = conv_layer(x, sd, "final_conv.", 3, False)
    x = x.reshape(x.shape[0], -1)
    x = equal_linear(x, sd["final_linear.0.weight"], sd["final_linear.0.bias"], activation=True)
    return equal_linear(x, sd["final_linear.1.weight"], sd["final_linear.1.bias"])


# ----------------------------------------------------------------------------- L1/L2: regularisers
def r1_penalty(scores, real_img):
    """grad_penalty_loss(inputs=[real_img], outs=scores, step=None) losses.py:87-99 -> (B,) = 5*||dD/dx||^2."""
    (g,) = torch.autograd.grad(scores.sum(), real_img, create_graph=True)
    return 5.0 * g.reshape(g.shape[0], -1).pow(2).sum(dim=1)


def path_length_penalty(cond, input_indices, sd, pl_noise, pl_mean=0.0, step=6, pl_decay=0.01):
    """Adopted rule for PathLengthRegularizor (reference code losses.py:102-124 is unrunnable):
    differentiate w.r.t. w = z_to_w(embd[idx]); y = pl_noise / sqrt(numel(img)) (batch included, :114);
    length = mean_b ||J^T y||_2 (:116); ema <- ema + decay*len - ema (:119); penalty (len-ema)^2 (:122);
    create_graph=True so the penalty has a gradient.  Returns (penalty, new_ema, length)."""
    z = sd["image_embedding.embd_weight"][input_indices]
    w = mapping_network(z, sd)
    if not w.requires_grad:            # mapping weights frozen in the caller's state dict: w is then the leaf
        w = w.detach().requires_grad_(True)
    img = synthesis(w, cond, sd, step)
    y = pl_noise / math.sqrt(img.numel())
    (g,) = torch.autograd.grad((img * y).sum(), w, create_graph=True)
    length = g.pow(2).sum(dim=1).sqrt().mean()
    ema = pl_mean + pl_decay * length - pl_mean
    return (length - ema).pow(2), ema.detach(), length


def d_logistic_loss(real_scores, fake_scores):
    """train.py:144,171-172."""
    return F.softplus(-real_scores).mean() + F.softplus(fake_scores).mean()


def g_nonsat_loss(fake_scores):
    """train.py:203."""
    return F.softplus(-fake_scores).mean()
