"""Tokens/s of classifier-free-guided sampling with the KV cache (BASELINE.json configs[4]: 24 layers, 1024 wide, 16
heads, 128 text + 256 segmentation tokens -> 256 image tokens of an 8192-code vocabulary; random-init weights).
One JSON line; not part of bench.py's headline (the reference has no sampler to compare with: SURVEY.md 8f-3).
Usage: python tools/bench_sampling.py [--batch B] [--reps N]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "make-a-scene_b200")]
import torch  # noqa: E402
from mas_b200 import _lib  # noqa: E402
from models.transformer import MakeAScene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--layers", type=int, default=24)
ap.add_argument("--graphs", action="store_true", help="replay each position's decode step from a CUDA graph")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = dict(num_layers=args.layers, hidden_dim=1024, num_attn_heads=16, image_vocab_size=8192, seg_vocab_size=1024, text_vocab_size=49408,
           image_tokens_per_dim=16, seg_tokens_per_dim=16, text_length=128)
m = MakeAScene(**cfg).to(dev).eval()
m.device = dev
B = args.batch
text = torch.randint(1, 40000, (B, 128), device=dev)
seg = torch.randint(0, 1024, (B, 256), device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
m.generate(text, seg, guidance_scale=3.0, temperature=1.0, top_k=64, generator=gen, use_graphs=args.graphs)     # warm-up (captures)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = _lib.launch_count()
e0.record()
for _ in range(args.reps):
    toks = m.generate(text, seg, guidance_scale=3.0, temperature=1.0, top_k=64, generator=gen, use_graphs=args.graphs)
e1.record()
torch.cuda.synchronize()
sec = e0.elapsed_time(e1) * 1e-3 / args.reps
params = sum(p.numel() for n, p in m.named_parameters() if "embedding" not in n)
print(json.dumps({"metric": "image tokens/s, classifier-free guided sampling with KV cache", "value": B * 256 / sec, "unit": "tokens/s",
                  "batch": B, "rows": 2 * B, "cuda_graphs": bool(args.graphs), "seconds_per_image": sec, "launches_per_image": (_lib.launch_count() - l0) // args.reps,
                  "weight_stream_gb_per_token": params * 4 / 1e9,
                  "weight_stream_gbs": params * 4 * 255 / sec / 1e9, "config": cfg}))
