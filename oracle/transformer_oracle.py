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
