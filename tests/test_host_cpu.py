"""CPU-side checks: the C-ABI library loads and exports every symbol include/pmv_b200.h declares (no compute
calls without a GPU), the reference-shaped modules keep the reference state_dict, host index logic."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from pmv_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pmv_b200.h")).read()
    declared = set(re.findall(r"\b(pmv_[a-z0-9_]+)\s*\(", hdr))
    assert "pmv_gemm" in declared and "pmv_pool_ln_fwd" in declared and "pmv_attention_fwd" in declared
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, missing
    assert declared - {"pmv_last_error"} == set(_lib.exported_symbols())  # ctypes table covers the whole header
    lib = _lib.lib()
    assert lib.pmv_version() >= 100
    assert lib.pmv_has_tcgen05() in (0, 1)


def test_product_path_fails_loudly_without_gpu_tensors():
    from pmv_b200 import ops
    x = torch.randn(4, 96)
    with pytest.raises((AssertionError, RuntimeError)):
        ops.layernorm_fwd(x, torch.ones(96), torch.zeros(96), torch.float32)


def test_state_dict_matches_reference_keys():
    from oracle import mvit_oracle as orc
    from pmv_b200 import mvit
    m = mvit.MViT(mvit.MVITV2_S)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == orc.param_shapes(orc.MVITV2_S)
    assert sum(p.numel() for p in m.parameters()) == 34537744
    assert mvit.block_schedule(mvit.MVITV2_B) == orc.block_schedule(orc.MVITV2_B)


def test_unsupported_configurations_raise():
    from pmv_b200.attention import MultiScaleAttention
    kw = dict(dim=96, dim_out=96, input_size=[2, 8, 8], num_heads=1, kernel_q=[3, 3, 3], kernel_kv=[3, 3, 3],
              stride_q=[1, 1, 1], stride_kv=[1, 2, 2])
    MultiScaleAttention(**kw)
    for bad in (dict(pool_first=True), dict(mode="avg"), dict(separate_qkv=True), dict(num_heads=2), dict(mode="bogus")):
        with pytest.raises(NotImplementedError):
            MultiScaleAttention(**{**kw, **bad})


@pytest.mark.parametrize("q,k", [(56, 7), (7, 14), (4, 7), (3, 6), (14, 14), (5, 3)])
def test_rel_index_table_matches_oracle(q, k):
    from oracle import mvit_oracle as orc
    from pmv_b200 import ops
    got = ops.rel_index_table(q, k, "cpu").view(q, k)
    assert torch.equal(got.long(), orc.rel_index(q, k))
    assert int(got.min()) >= 0 and int(got.max()) <= 2 * max(q, k) - 2


def test_aug_ld():
    from pmv_b200 import ops
    assert ops.aug_ld((8, 7, 7)) == 128 and ops.aug_ld((8, 14, 14)) == 160 and ops.aug_ld((16, 7, 7)) == 128


def test_checkpoint_loader_matches_reference_surgery(tmp_path):
    """a 224-crop MViTv2-S checkpoint loaded into a 160-crop model.  The rel-pos tables change
    length and are interpolated; the result must equal what the reference's own load_checkpoint produced
    (tests/golden/checkpoint_surgery.npz, oracle/make_golden.py::run_checkpoint_case)."""
    import numpy as np
    import torch
    from oracle import detgen, mvit_oracle as orc
    from pmv_b200 import mvit
    from pmv_b200.checkpoint import load_checkpoint
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "checkpoint_surgery.npz"))
    seed = int(z["seed"])
    params = detgen.det_params(orc.param_shapes(orc.MVITV2_S), seed)
    path = str(tmp_path / "ck.pyth")
    torch.save({"model_state": {"module." + k: v for k, v in params.items()}, "epoch": 3}, path)  # DDP-style names
    model = mvit.MViT(dict(mvit.MVITV2_S, crop=(160, 160)))
    before = {k: v.clone() for k, v in model.state_dict().items()}
    epoch = load_checkpoint(path, model)
    assert epoch == int(z["epoch"]) == 3
    sd = model.state_dict()
    for k in z.files:
        if k.startswith("blocks."):
            assert tuple(sd[k].shape) == z[k].shape and sd[k].shape != params[k].shape or "rel_pos_t" in k
            assert float((sd[k] - torch.from_numpy(z[k])).abs().max()) < 1e-6, k
    names = list(z["names"])
    assert names == list(sd.keys())
    got = np.array([float(v.double().sum()) for v in sd.values()])
    assert np.allclose(got, z["checksum"], rtol=1e-6, atol=1e-6)
    rep = load_checkpoint.last_report
    assert rep["not_used"] == [] and rep["not_loaded"] == [] and rep["missing"] == []
    assert any(float((before[k] - sd[k]).abs().max()) > 0 for k in sd)
    # same-shape load == strict load; epoch_reset
    model224 = mvit.MViT(mvit.MVITV2_S)
    assert load_checkpoint({"model_state": params, "epoch": 7}, model224, epoch_reset=True) == -1
    for k, v in model224.state_dict().items():
        assert torch.equal(v, params[k]), k


def test_param_groups_match_reference_construct_optimizer():
    """Row f3 host logic: the zero-weight-decay grouping equals what the reference's construct_optimizer built for its own
    MViTv2-S (fixture written by oracle/make_golden.py from models/optimizer.py:29-83)."""
    from pmv_b200 import mvit
    from pmv_b200.optim import param_groups
    z = np.load(os.path.join(ROOT, "tests", "golden", "optimizer_reference.npz"))
    model = mvit.MViT(mvit.MVITV2_S)
    names = {id(p): n for n, p in model.named_parameters()}
    groups = param_groups(model, 0.05, zero_wd_1d=True)
    assert len(groups) == int(z["ngroups"])
    for gi, g in enumerate(groups):
        assert g["weight_decay"] == float(z[f"group{gi}_weight_decay"])
        assert [names[id(p)] for p in g["params"]] == [str(n) for n in z[f"group{gi}_names"]]
    # ZERO_DECAY_POS_CLS moves the relative-position tables and the cls token into the zero group (video_model_builder.py:2027-2049)
    model.cfg = dict(model.cfg, zero_decay_pos_cls=True)
    g2 = param_groups(model, 0.05, zero_wd_1d=True)
    zero = {names[id(p)] for p in g2[-1]["params"]}
    assert "cls_token" in zero and "blocks.0.attn.rel_pos_h" in zero and "blocks.3.attn.rel_pos_t" in zero
    assert "blocks.0.attn.qkv.weight" not in zero


def test_zero_arena_never_hands_out_uncleared_memory():
    """Host logic of the split-K weight-gradient arena (ops.ZeroArena): sizes itself from the previous step's demand, hands out
    zeroed, 256-byte aligned slices after reset(), and returns None (fallback to torch.zeros) when it was not reset."""
    from pmv_b200.ops import ZeroArena
    dev = torch.device("cpu")
    a = ZeroArena()
    a.reset(dev)
    assert a.take(1000, dev) is None and a.take(70, dev) is None        # first step: nothing allocated yet, demand recorded
    a.reset(dev)                                                         # second step: sized for 1024 + 128 floats
    x, y = a.take(1000, dev), a.take(70, dev)
    assert x is not None and y is not None and x.numel() == 1000 and y.numel() == 70
    assert float(x.abs().sum()) == 0.0 and float(y.abs().sum()) == 0.0
    assert (y.data_ptr() - x.data_ptr()) % 256 == 0 and y.data_ptr() - x.data_ptr() >= 4000
    x.fill_(3.0); y.fill_(5.0)                                            # gradients of this step
    # a second backward without reset(): the arena must not hand the dirty memory out again
    z = a.take(1000, dev)
    assert z is None or float(z.abs().sum()) == 0.0
    if z is not None:
        assert z.data_ptr() >= y.data_ptr() + 70 * 4
    a.reset(dev)                                                         # next step: everything cleared again
    x2 = a.take(1000, dev)
    assert x2 is not None and float(x2.abs().sum()) == 0.0
