"""CPU: the C-ABI library loads and exports every symbol include/mas_b200.h declares (no compute calls);
host-side module structure mirrors the reference (state_dict keys, init order)."""
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mas_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from mas_b200 import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    # the ctypes table and the header agree one to one
    assert sorted(_lib.exported_symbols()) == names
    assert lib.mas_version() == 100
    assert isinstance(lib.mas_last_error(), bytes)
    assert ctypes.sizeof(_lib.Tensor4) == 64


def test_product_fails_loudly_without_cuda():
    import pytest
    from mas_b200 import ops
    with pytest.raises(RuntimeError):
        ops.nhwc(torch.zeros(1, 4, 2, 2))


def test_state_dict_matches_reference_keys_and_init():
    g = torch.load(os.path.join(GOLDEN, "vqbase_tiny.pt"), weights_only=False)
    from models import VQBASE
    torch.manual_seed(0)
    m = VQBASE(g["ddconfig"], g["n_embed"], g["embed_dim"], 10, 100)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["state_dict"].keys())
    for k, v in g["state_dict"].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    m.load_state_dict(g["state_dict"])
    assert m.decoder.model[-1].weight.shape == (3, 32, 3, 3)            # last_layer contract, train.py:96
    assert m.quantize.q_counter == 0 and m.quantize.reservoir is None    # plain attributes, modules.py:466-468
    assert (m.quantize.q_start_collect, m.quantize.q_init, m.quantize.q_re_end, m.quantize.q_re_step) == (10, 30, 300, 5)


def test_img_config_init_is_bit_identical_to_reference_init():
    """Same construction order => torch.manual_seed(0) reproduces the reference's initial weights exactly."""
    g = torch.load(os.path.join(GOLDEN, "vqbase_img_64.pt"), weights_only=False)
    from models import VQBASE
    torch.manual_seed(0)
    m = VQBASE(g["ddconfig"], 8192, 256, 3000, 12500)
    with torch.no_grad():
        m.quantize.embedding.weight.normal_()
    sd = m.state_dict()
    assert len(sd) == 348 and sum(p.numel() for p in m.parameters()) == g["n_params"] == 95219075
    for k, (s, a) in g["param_checks"].items():
        v = sd[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(s)) and abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, a), k


def test_yaml_model_blocks_instantiate():
    import yaml
    from models import VQBASE
    for name in ("img_config.yaml", "seg_config.yaml"):
        cfg = yaml.safe_load(open(os.path.join(ROOT, "make-a-scene_b200", "conf", name)))["model"]
        assert cfg.pop("_target_") == "models.VQBASE"
        if name == "img_config.yaml":
            continue  # 95M params: covered by the test above
        cfg["ddconfig"]["channels"] = [32, 32, 64]  # keep the CPU test light; stale keys must be tolerated
        m = VQBASE(**cfg)
        assert m.encoder.model[0].in_channels == 159


def test_weight_pack_cache_is_keyed_by_identity_and_version(monkeypatch):
    """Host logic of ops._packed_conv_weight (no kernels run: the C-ABI call is stubbed): a weight is packed once per
    version (forward + data-gradient images in one pass), re-used by every later forward / backward until the version
    counter moves, and never confused with another tensor that re-uses its address or id."""
    import gc
    import torch
    from mas_b200 import ops
    calls = []
    monkeypatch.setattr(ops.L, "call", lambda name, *a: calls.append(name))
    ops._packs.clear()
    for f16, pair, single in ((False, "mas_pack_conv3x3_tc_pair", "mas_pack_conv3x3_tc"),
                              (True, "mas_pack_conv3x3_tc16", "mas_pack_conv3x3_tc16")):
        calls.clear()
        w = torch.zeros(128, 128, 3, 3)
        dev = w.device
        f1 = ops._packed_conv_weight(w, w, 128, 128, False, dev, prepack=True, f16=f16)
        assert calls == [pair]
        d1 = ops._packed_conv_weight(w, w, 128, 128, True, dev, f16=f16)      # backward of the same step: no packing launch
        d2 = ops._packed_conv_weight(w, w, 128, 128, True, dev, f16=f16)      # a second backward (retained graph): still none
        f2 = ops._packed_conv_weight(w, w, 128, 128, False, dev, prepack=True, f16=f16)   # next forward, same weights: none
        assert calls == [pair] and d1 is d2 and f1 is f2 and d1 is not f1
        assert f1.dtype == (torch.float16 if f16 else torch.float32)
        w.add_(1.0)                                                         # optimiser step: version changed -> repack
        ops._packed_conv_weight(w, w, 128, 128, True, dev, f16=f16)
        assert calls == [pair, single]
        ops._packed_conv_weight(w, w, 128, 128, False, dev, prepack=True, f16=f16)
        assert len(calls) == 3
        # not pair-eligible (Cout % 128 != 0 for the data-gradient image): plain packing of the requested image only
        w2 = torch.zeros(128, 64, 3, 3)
        ops._packed_conv_weight(w2, w2, 128, 64, False, dev, prepack=True, f16=f16)
        assert calls[-1] == single
        n_live = len(ops._packs)
        del w, w2, f1, f2, d1, d2
        gc.collect()
        assert len(ops._packs) == n_live - 2                                 # entries die with their weights


def test_shadow_and_amax_attachments_follow_the_tensor_version():
    """ops.shadow_of / ops.amax_of only trust what a producing kernel attached while the tensor is unchanged since (same
    version counter, same storage address, same device and shape): host-side bookkeeping of the fp16 operand shadows."""
    import torch
    from mas_b200 import ops
    t = torch.zeros(2, 8, 4, 4).contiguous(memory_format=torch.channels_last)
    sh16 = torch.zeros_like(t, dtype=torch.float16)
    bound = torch.ones(1)
    assert ops.shadow_of(t) is None
    t._mas_shadow = (sh16, bound, t._version, t.data_ptr())
    got = ops.shadow_of(t)
    assert got is not None and got[0] is sh16 and got[1] is bound
    t.add_(1.0)                                   # in-place update: the shadow no longer describes the tensor
    assert ops.shadow_of(t) is None
    u = torch.zeros(2, 8, 4, 4).contiguous(memory_format=torch.channels_last)
    u._mas_shadow = (torch.zeros(2, 8, 2, 2, dtype=torch.float16), bound, u._version, u.data_ptr())
    assert ops.shadow_of(u) is None               # shape mismatch
    am = torch.ones(1)
    ops.attach_amax(u, am)
    assert getattr(u, "_mas_amax")[0] is am
    u.mul_(2.0)
    assert getattr(u, "_mas_amax")[1] != u._version


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """The drop-in boundary is a C ABI: include/mas_b200.h compiles as C99 (no C++ types in the signatures) and a C program
    links against the shared library and calls an entry that needs no GPU."""
    import shutil
    import subprocess
    from mas_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    src = tmp_path / "use_abi.c"
    src.write_text('#include "mas_b200.h"\n#include <stdio.h>\nint main(void) { printf("%d\\n", mas_version()); return mas_last_error() ? 0 : 1; }\n')
    exe = tmp_path / "use_abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir, "-lmas_b200",
                        "-Wl,-rpath," + libdir, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and int(out.stdout.strip()) == _lib.load().mas_version()
