"""CPU oracle for the tier-2 token transformer — TEST INFRASTRUCTURE ONLY (same rules as vqgan_oracle.py).

Functional torch-CPU restatement of /root/reference/models/transformer.py (default flags: cogview_pb_relax=True,
sandwich LayerNorm, no prescale, no rudalle_relax, no KV cache):
  gelu                      transformer.py:11-14
  SelfAttention             transformer.py:17-115   (PB-relax is a per-(batch,head) constant shift: softmax-invariant)
  MLP                       transformer.py:118-139
  TransformerLayer          transformer.py:142-210
  Transformer (+ mask)      transformer.py:213-272
  MakeAScene.forward        transformer.py:349-378
Pinned by tests/golden/transformer_tiny.pt (generated from the real reference by oracle/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F


def gelu(x):
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


def layer_norm(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def linear(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def self_attention(x, sd, p, heads, mask):
    b, s, h = x.shape
    hd = h // heads
    qkv = linear(x, sd, p + ".qkv")
    q, k, v = torch.split(qkv, h, dim=-1)
    q, k, v = [t.view(b, s, heads, hd).permute(0, 2, 1, 3) for t in (q, k, v)]
    scores = torch.matmul(q / math.sqrt(hd), k.transpose(-1, -2))
    scores = mask * scores - (1.0 - mask) * 10000.0                   # transformer.py:62
    alpha = 32.0                                                       # PB-relax, transformer.py:63-70
    sc = scores / alpha
    mx = sc.detach().view(b, heads, -1).max(dim=-1)[0][..., None, None]
    scores = (sc - mx) * alpha
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous().view(b, s, h)
    return linear(ctx, sd, p + ".out_proj")


def transformer_layer(x, sd, p, heads, mask):
    a = self_attention(layer_norm(x, sd, p + ".ln_in"), sd, p + ".attn", heads, mask)
    x = x + layer_norm(a, sd, p + ".first_ln_sandwich")
    m = linear(gelu(linear(layer_norm(x, sd, p + ".ln_out"), sd, p + ".mlp.lin1")), sd, p + ".mlp.lin2")
    return x + layer_norm(m, sd, p + ".second_ln_sandwich")


def make_a_scene_forward(sd, cfg, text_tokens, seg_tokens, img_tokens):
    """cfg: dict(num_layers, hidden_dim, num_attn_heads, image_vocab_size, seg_vocab_size, text_vocab_size,
    image_tokens_per_dim, seg_tokens_per_dim, text_length)."""
    tl, sp, ip = cfg["text_length"], cfg["seg_tokens_per_dim"], cfg["image_tokens_per_dim"]
    il, sl = ip * ip, sp * sp
    total = tl + sl + il
    text_range = torch.arange(tl) + (cfg["text_vocab_size"] - tl)
    text_tokens = torch.where(text_tokens == 0, text_range, text_tokens)
    emb = [F.embedding(text_tokens, sd["text_token_embedding.weight"]) + sd["text_pos_embeddings.weight"][:text_tokens.shape[1]]]
    ids = torch.arange(seg_tokens.shape[-1])
    emb.append(F.embedding(seg_tokens, sd["seg_token_embedding.weight"]) + sd["seg_row_embeddings.weight"][ids // sp]
               + sd["seg_col_embeddings.weight"][ids % sp])
    if img_tokens is not None:
        ids = torch.arange(img_tokens.shape[-1])
        emb.append(F.embedding(img_tokens, sd["image_token_embedding.weight"]) + sd["image_row_embeddings.weight"][ids // ip]
                   + sd["image_col_embeddings.weight"][ids % ip])
    x = torch.cat(emb, dim=1)
    am = torch.tril(torch.ones(x.shape[0], 1, total, total))
    am[:, :, :-il, :-il] = 1
    am = am[:, :, :x.shape[1], :x.shape[1]]
    mask = am * sd["transformer.mask"][:am.size(2), :am.size(3)]       # transformer.py:262-263 -> plain causal
    for i in range(cfg["num_layers"]):
        x = transformer_layer(x, sd, f"transformer.layers.{i}", cfg["num_attn_heads"], mask)
    x = layer_norm(x, sd, "transformer.final_ln")
    logits = linear(layer_norm(x, sd, "to_logits.0"), sd, "to_logits.1")
    return logits[:, -il - 1:-1, :]


# ------------------------------------------------------------------------------------------------ sampling
# The reference has no working sampler / KV cache (transformer.py:73-115 vs :176-210); its only specification of the
# autoregressive path is the non-cached forward above.  The functions below restate the cached algorithm the CUDA path
# implements (MakeAScene.generate) on the CPU so it can be pinned against the reference fixture without a GPU:
# a decode step over the cache must give the logits make_a_scene_forward gives at that position.
def _attend_cached(q, kc, vc, heads):
    """q [R,H] (one new position per row) against caches kc / vc [R,heads,T,hd] holding positions 0..T-1 (incl. this one)."""
    r, h = q.shape
    hd = h // heads
    qh = q.view(r, heads, 1, hd)
    p = torch.softmax(torch.matmul(qh / math.sqrt(hd), kc.transpose(-1, -2)), dim=-1)   # every cached key is visible
    return torch.matmul(p, vc).reshape(r, h)


def generate_logits(sd, cfg, text_tokens, seg_tokens, img_tokens=None, guidance_scale=None, uncond_text_tokens=None):
    """Per-position logits [B, image_length, V] of KV-cached decoding: greedy when img_tokens is None, teacher-forced
    otherwise; with guidance_scale the cond / uncond streams are mixed as uncond + s * (cond - uncond) before the arg-max
    (Make-A-Scene paper 3.4; the unconditional stream sees all-padding text).  Returns (tokens, logits)."""
    heads, layers = cfg["num_attn_heads"], cfg["num_layers"]
    tl, sp, ip = cfg["text_length"], cfg["seg_tokens_per_dim"], cfg["image_tokens_per_dim"]
    il = ip * ip
    B = text_tokens.shape[0]
    cfg_on = guidance_scale is not None and float(guidance_scale) != 1.0
    if cfg_on:
        if uncond_text_tokens is None:
            uncond_text_tokens = torch.zeros_like(text_tokens)
        text_all, seg_all = torch.cat([text_tokens, uncond_text_tokens], 0), torch.cat([seg_tokens, seg_tokens], 0)
    else:
        text_all, seg_all = text_tokens, seg_tokens
    R = text_all.shape[0]
    # ---- prefix: full causal pass, recording k / v of every layer
    text_range = torch.arange(tl) + (cfg["text_vocab_size"] - tl)
    tt = torch.where(text_all == 0, text_range, text_all)
    ids = torch.arange(seg_all.shape[-1])
    x = torch.cat([F.embedding(tt, sd["text_token_embedding.weight"]) + sd["text_pos_embeddings.weight"][:tt.shape[1]],
                   F.embedding(seg_all, sd["seg_token_embedding.weight"]) + sd["seg_row_embeddings.weight"][ids // sp]
                   + sd["seg_col_embeddings.weight"][ids % sp]], dim=1)
    P = x.shape[1]
    h = x.shape[-1]
    hd = h // heads
    mask = torch.tril(torch.ones(P, P))
    kcs, vcs = [], []
    for i in range(layers):
        p = f"transformer.layers.{i}"
        qkv = linear(layer_norm(x, sd, p + ".ln_in"), sd, p + ".attn.qkv")
        _, k, v = torch.split(qkv, h, dim=-1)
        kcs.append(k.view(R, P, heads, hd).permute(0, 2, 1, 3))
        vcs.append(v.view(R, P, heads, hd).permute(0, 2, 1, 3))
        x = transformer_layer(x, sd, p, heads, mask)

    def logits_of(hid):
        return linear(layer_norm(layer_norm(hid, sd, "transformer.final_ln"), sd, "to_logits.0"), sd, "to_logits.1")
    logits = logits_of(x[:, -1])
    toks, kept = [], []
    for t in range(il):
        mixed = logits[B:] + float(guidance_scale) * (logits[:B] - logits[B:]) if cfg_on else logits
        kept.append(mixed)
        tok = img_tokens[:, t] if img_tokens is not None else mixed.argmax(-1)
        toks.append(tok)
        if t == il - 1:
            break
        tok_all = torch.cat([tok, tok], 0) if cfg_on else tok
        xt = (F.embedding(tok_all, sd["image_token_embedding.weight"]) + sd["image_row_embeddings.weight"][t // ip]
              + sd["image_col_embeddings.weight"][t % ip])
        for i in range(layers):
            p = f"transformer.layers.{i}"
            qkv = linear(layer_norm(xt, sd, p + ".ln_in"), sd, p + ".attn.qkv")
            q, k, v = torch.split(qkv, h, dim=-1)
            kcs[i] = torch.cat([kcs[i], k.view(R, heads, 1, hd)], dim=2)
            vcs[i] = torch.cat([vcs[i], v.view(R, heads, 1, hd)], dim=2)
            a = linear(_attend_cached(q, kcs[i], vcs[i], heads), sd, p + ".attn.out_proj")
            xt = xt + layer_norm(a, sd, p + ".first_ln_sandwich")
            m = linear(gelu(linear(layer_norm(xt, sd, p + ".ln_out"), sd, p + ".mlp.lin1")), sd, p + ".mlp.lin2")
            xt = xt + layer_norm(m, sd, p + ".second_ln_sandwich")
        logits = logits_of(xt)
    return torch.stack(toks, 1), torch.stack(kept, 1)
