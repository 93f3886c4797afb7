"""CPU oracle for the VQ-IMG hot path — TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, functional restatement (torch CPU fp32 ops + numpy for the
integer-valued codebook argmin) of the reference algorithm in
  /root/reference/models/vqvae.py:8-39      (VQBASE)
  /root/reference/models/modules.py:35-41   (nonlinearity, Normalize)
  /root/reference/models/modules.py:44-81   (Upsample, Downsample)
  /root/reference/models/modules.py:84-136  (ResnetBlock)
  /root/reference/models/modules.py:139-191 (AttnBlock)
  /root/reference/models/modules.py:199-240 (Encoder)
  /root/reference/models/modules.py:337-369 (Decoder)
  /root/reference/models/modules.py:451-528 (Codebook)
  /root/reference/losses/loss_seg.py:6-22   (BCELossWithQuant)
It is driven by a plain ``state_dict`` (dict of tensors with the reference's key names) and the
``ddconfig`` dict, so it shares no code with the product package.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and only as the checker or the reported CPU baseline. The product path
(``make-a-scene_b200/``) never imports this module and fails loudly without its CUDA library.

Parity pin: ``oracle/make_golden.py`` runs the REAL reference (imported from /root/reference in the
authoring container) and stores inputs/outputs/grads under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against those fixtures. The arithmetic itself is PyTorch/ATen
(torch 2.11.0+cu128), which the reference does not pin.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

GN_GROUPS = 32      # modules.py:41
GN_EPS = 1e-6       # modules.py:41
BN_EPS = 1e-5       # nn.SyncBatchNorm default, vqvae.py:16
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------- primitives
def swish(x):
    """modules.py:35-37 / :194-196 — x * sigmoid(x)."""
    return x * torch.sigmoid(x)


def normalize(x, sd, p):
    """modules.py:40-41 — GroupNorm(32, C, eps=1e-6, affine)."""
    return F.group_norm(x, GN_GROUPS, sd[p + ".weight"], sd[p + ".bias"], GN_EPS)


def conv(x, sd, p, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def upsample(x, sd, p):
    """modules.py:55-59 — nearest x2 then conv3x3 pad 1."""
    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    return conv(x, sd, p + ".conv", 1, 1)


def downsample(x, sd, p):
    """modules.py:74-78 — zero pad right/bottom by one, conv3x3 stride 2 pad 0."""
    x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
    return conv(x, sd, p + ".conv", 2, 0)


def resnet_block(x, sd, p):
    """modules.py:119-136 (dropout p=0.0 is the identity, modules.py:224)."""
    h = swish(normalize(x, sd, p + ".norm1"))
    h = conv(h, sd, p + ".conv1", 1, 1)
    h = swish(normalize(h, sd, p + ".norm2"))
    h = conv(h, sd, p + ".conv2", 1, 1)
    if (p + ".nin_shortcut.weight") in sd:
        x = conv(x, sd, p + ".nin_shortcut", 1, 0)
    return x + h


def attn_block(x, sd, p):
    """modules.py:167-191 — single-head spatial attention, softmax over keys."""
    h_ = normalize(x, sd, p + ".norm")
    q = conv(h_, sd, p + ".q")
    k = conv(h_, sd, p + ".k")
    v = conv(h_, sd, p + ".v")
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    h_ = conv(h_, sd, p + ".proj_out")
    return x + h_


# ----------------------------------------------------------------------------- layer plans
def encoder_plan(in_channels=3, channels=(128, 128, 128, 256, 512, 512), attn_resolutions=(32,),
                 resolution=512, dropout=0.0, num_res_blocks=2, z_channels=256, **kwargs):
    """Layer list of Encoder.__init__, modules.py:217-237. Returns [(kind, cin, cout)]."""
    plan = [("conv3", in_channels, channels[0])]
    for i in range(len(channels) - 1):
        cin, cout = channels[i], channels[i + 1]
        for _ in range(num_res_blocks):
            plan.append(("res", cin, cout))
            cin = cout
            if resolution in attn_resolutions:
                plan.append(("attn", cin, cin))
        if i < len(channels) - 2:
            plan.append(("down", channels[i + 1], channels[i + 1]))
            resolution //= 2
    c = channels[-1]
    plan += [("res", c, c), ("attn", c, c), ("res", c, c), ("norm", c, c), ("swish", c, c),
             ("conv3", c, z_channels)]
    return plan


def decoder_plan(out_channels=3, channels=(128, 128, 128, 256, 512, 512), attn_resolutions=(32,),
                 resolution=512, dropout=0.0, num_res_blocks=2, z_channels=256, **kwargs):
    """Layer list of Decoder.__init__, modules.py:338-366."""
    ch_mult = list(channels[1:])
    nres = len(ch_mult)
    block_in = ch_mult[nres - 1]
    curr_res = resolution // 2 ** (nres - 1)
    plan = [("conv3", z_channels, block_in), ("res", block_in, block_in), ("attn", block_in, block_in),
            ("res", block_in, block_in)]
    for i in reversed(range(nres)):
        block_out = ch_mult[i]
        for _ in range(num_res_blocks + 1):
            plan.append(("res", block_in, block_out))
            block_in = block_out
            if curr_res in attn_resolutions:
                plan.append(("attn", block_in, block_in))
        if i > 0:
            plan.append(("up", block_in, block_in))
        curr_res *= 2
    plan += [("norm", block_in, block_in), ("swish", block_in, block_in), ("conv3", block_in, out_channels)]
    return plan


def run_plan(plan, sd, prefix, x, taps=None):
    """nn.Sequential forward over the plan (modules.py:239-240, 368-369)."""
    for i, (kind, _cin, _cout) in enumerate(plan):
        p = f"{prefix}.{i}"
        if kind == "conv3":
            x = conv(x, sd, p, 1, 1)
        elif kind == "res":
            x = resnet_block(x, sd, p)
        elif kind == "attn":
            x = attn_block(x, sd, p)
        elif kind == "down":
            x = downsample(x, sd, p)
        elif kind == "up":
            x = upsample(x, sd, p)
        elif kind == "norm":
            x = normalize(x, sd, p)
        elif kind == "swish":
            x = swish(x)
        else:
            raise ValueError(kind)
        if taps is not None:
            taps[p] = x
    return x


# ----------------------------------------------------------------------------- codebook
def codebook_distances(z_flat, E):
    """modules.py:501-503 — d = sum(z^2) + sum(e^2) - 2 z.e^T, same association."""
    return (torch.sum(z_flat ** 2, dim=1, keepdim=True) + torch.sum(E ** 2, dim=1)
            - 2 * torch.einsum("bd,dn->bn", z_flat, E.t()))


def codebook_argmin_numpy(z_flat: np.ndarray, E: np.ndarray) -> np.ndarray:
    """numpy fp32 restatement of modules.py:501-505 (integer-valued result, first-index ties)."""
    z = np.asarray(z_flat, dtype=np.float32)
    e = np.asarray(E, dtype=np.float32)
    zz = np.sum(z * z, axis=1, keepdims=True, dtype=np.float32)
    ee = np.sum(e * e, axis=1, dtype=np.float32)
    d = (zz + ee) - np.float32(2.0) * (z @ e.T)
    return np.argmin(d, axis=1).astype(np.int64)


def codebook_gap_fp64(z_flat, E, idx_a, idx_b):
    """fp64 distance gap |d[idx_a]-d[idx_b]| per row and ulp(d) in fp32 — classifies mismatches as ties."""
    z = z_flat.double()
    e = E.double()
    da = ((z - e[idx_a]) ** 2).sum(1)
    db = ((z - e[idx_b]) ** 2).sum(1)
    ulp = torch.abs(da).float().clamp_min(1e-30)
    ulp = torch.nextafter(ulp, torch.full_like(ulp, float("inf"))) - ulp
    return (da - db).abs(), ulp.double()


def codebook_forward(z, E, beta=0.25):
    """Steady-state Codebook.forward (modules.py:470-473, 501-517): returns z_q (NCHW), loss, idx."""
    zp = z.permute(0, 2, 3, 1).contiguous()
    zf = zp.view(-1, E.shape[1])
    d = codebook_distances(zf, E)
    idx = torch.argmin(d, dim=1)
    z_q = F.embedding(idx, E).view(zp.shape)
    loss = torch.mean((z_q.detach() - zp) ** 2) + beta * torch.mean((z_q - zp.detach()) ** 2)
    z_q = zp + (z_q - zp).detach()
    return z_q.permute(0, 3, 1, 2).contiguous(), loss, idx


def codebook_entry(E, indices, shape=None):
    """modules.py:519-528."""
    z_q = F.embedding(indices, E)
    if shape is not None:
        z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
    return z_q


# ----------------------------------------------------------------------------- VQBASE
def quant_conv(h, sd, training=True):
    """vqvae.py:14-17,22 — Conv1x1 + (Sync)BatchNorm; batch statistics in training mode."""
    h = conv(h, sd, "quant_conv.0")
    return F.batch_norm(h, sd["quant_conv.1.running_mean"].detach().clone(), sd["quant_conv.1.running_var"].detach().clone(),
                        sd["quant_conv.1.weight"], sd["quant_conv.1.bias"], training, BN_MOMENTUM, BN_EPS)


def vqbase_forward(sd, ddconfig, x, quantize=True, training=True, beta=0.25, taps=None):
    """VQBASE.forward (vqvae.py:36-39). ``quantize=False`` is the warm-up bypass of modules.py:482-484."""
    h = run_plan(encoder_plan(**ddconfig), sd, "encoder.model", x, taps)
    h = quant_conv(h, sd, training)
    if taps is not None:
        taps["quant_conv"] = h
    if quantize:
        quant, diff, idx = codebook_forward(h, sd["quantize.embedding.weight"], beta)
    else:
        quant, diff, idx = h, h.new_tensor(0), None
    q = conv(quant, sd, "post_quant_conv")
    dec = run_plan(decoder_plan(**ddconfig), sd, "decoder.model", q, taps)
    return dec, diff, idx


def proxy_loss(img, dec, diff):
    """Benchmark proxy loss: L1 reconstruction + codebook term (loss_img.py:79,124)."""
    return (img - dec).abs().mean() + diff


# ----------------------------------------------------------------------------- seg loss
def bce_loss_with_quant(qloss, target, prediction, image_channels=159, codebook_weight=1.0):
    """losses/loss_seg.py:6-22 — pos_weight 20 on channels 153..157."""
    w = torch.ones(image_channels)
    if image_channels >= 158:
        w[153:158] = 20
    bce = F.binary_cross_entropy_with_logits(prediction.permute(0, 2, 3, 1), target.permute(0, 2, 3, 1),
                                             pos_weight=w)
    return bce + codebook_weight * qloss
