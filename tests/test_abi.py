"""CPU: the C-ABI library loads and exports every symbol include/mas_b200.h declares (no compute calls);
host-side module structure mirrors the reference (state_dict keys, init order)."""
import os
import re

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


def test_dgrad_pack_handover_is_keyed_by_version_and_shape(monkeypatch):
    """Host logic of ops._packed_conv_weight (no kernels run: the C-ABI call is stubbed): the data-gradient packing made
    in the forward is handed to exactly one backward of the same, unmodified weight."""
    import torch
    from mas_b200 import ops
    calls = []
    monkeypatch.setattr(ops.L, "call", lambda name, *a: calls.append(name))
    ops._dgrad_packs.clear()
    w = torch.zeros(128, 128, 3, 3)
    dev = w.device
    ops._packed_conv_weight(w, w, 128, 128, False, dev, prepack=True)
    assert calls == ["mas_pack_conv3x3_tc_pair"] and len(ops._dgrad_packs) == 1
    ops._packed_conv_weight(w, w, 128, 128, True, dev)                 # backward of the same step: no packing launch
    assert calls == ["mas_pack_conv3x3_tc_pair"] and not ops._dgrad_packs
    ops._packed_conv_weight(w, w, 128, 128, True, dev)                 # a second backward (retained graph): repacks
    assert calls[-1] == "mas_pack_conv3x3_tc" and len(calls) == 2
    ops._packed_conv_weight(w, w, 128, 128, False, dev, prepack=True)
    w.add_(1.0)                                                         # optimiser step in between: version changed
    ops._packed_conv_weight(w, w, 128, 128, True, dev)
    assert calls[-1] == "mas_pack_conv3x3_tc"
    ops._packed_conv_weight(w, w, 128, 128, False, dev, prepack=True)
    key = next(iter(ops._dgrad_packs))
    ver, wd, _shape = ops._dgrad_packs[key]
    ops._dgrad_packs[key] = (ver, wd, (256, 128, 3, 3))                # a different weight that reused the address
    n = len(calls)
    ops._packed_conv_weight(w, w, 128, 128, True, dev)
    assert len(calls) == n + 1 and calls[-1] == "mas_pack_conv3x3_tc"
    ops._packed_conv_weight(w, w, 64, 128, False, dev, prepack=True)    # not pair-eligible: plain packing, nothing stored
    assert calls[-1] == "mas_pack_conv3x3_tc" and not ops._dgrad_packs
