cd /root/repo
python - <<'PY'
import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/make-a-scene_b200", "/root/repo/tools"]
import torch
from mas_b200 import _lib as L, ops
from micro_conv import bench
dev = torch.device("cuda:0")
ops.set_operand_format("f16")
B, C, H = 32, 128, 256
x = torch.randn(B, C, H, H, device=dev).contiguous(memory_format=torch.channels_last)
w = torch.randn(C, C, 3, 3, device=dev) * 0.03
b = torch.zeros(C, device=dev)
x16 = ops.to_half(x)
for pf in ("1", "0", "1", "0"):
    os.environ["MAS_TMA_RES_PREFETCH"] = pf
    ms = bench(lambda: ops.conv3x3_h_raw(x16, w, b, x, want_stats=True), it=10, warm=3)
    print("residual+stats prefetch=%s  %.3f ms" % (pf, ms), flush=True)
PY
python bench.py --graph --no-cpu-baseline --steps 6 --warmup 3 --step-only 2>/dev/null | tail -1
python bench.py --no-cpu-baseline --steps 6 --warmup 3 --step-only 2>/dev/null | tail -1
