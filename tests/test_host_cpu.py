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


def test_gelu_polynomial_of_the_gemm_epilogue():
    """The degree-8 erf-GELU polynomial compiled into csrc/gemm_tc_kernel.cuh (phi2): coefficients in the source equal
    oracle/fit_gelu.py's table, and the float32 Horner form stays within 1.1e-5 (Phi) / 5e-5 (gelu, on [-8, 8]) of the exact erf form
    the reference uses (common.py:13,21 nn.GELU())."""
    from oracle import fit_gelu
    src = open(os.path.join(ROOT, "portrait-mode-video_b200", "csrc", "gemm_tc_kernel.cuh")).read()
    body = src[src.index("float2 phi2(float2 x)"):src.index("float2 gelu2(float2 x)")]
    found = [float(v) for v in re.findall(r"make_float2\((-?\d\.\d+e[+-]\d+)f,", body)]
    assert found == list(fit_gelu.COEFFS), found
    phi_err, gelu_err = fit_gelu.sweep()
    assert phi_err < 1.1e-5 and gelu_err < 5e-5, (phi_err, gelu_err)
    # the committed fit script reproduces a table of the same quality, and the two polynomials agree as functions
    refit = tuple(fit_gelu.fit())
    phi_err2, _ = fit_gelu.sweep(coeffs=refit)
    assert phi_err2 < 1.5e-5, phi_err2
    xs = np.linspace(-6, 6, 20001)
    assert np.max(np.abs(fit_gelu.phi_poly_f32(xs, refit) - fit_gelu.phi_poly_f32(xs))) < 2.5e-5


def test_zero_arena_is_per_owner_and_safe_with_live_gradients():
    """the split-K weight-gradient arena must not be shared between models, must not hand out memory while
    gradients of zero_grad(set_to_none=False) live in it, and must not be reallocated once a graph captured it."""
    from pmv_b200 import ops
    dev = torch.device("cpu")
    pa = [torch.nn.Parameter(torch.zeros(4, 4)) for _ in range(2)]
    pb = [torch.nn.Parameter(torch.zeros(4, 4))]
    a, b = ops.attach_arena(pa), ops.attach_arena(pb)
    assert a is not b and pa[0]._pmv_arena is a and pa[1]._pmv_arena is a and pb[0]._pmv_arena is b
    assert ops.attach_arena(pa) is a  # an optimizer over the same model shares the reducer's arena
    for arena in (a, b):
        arena.reset(dev); arena.take(100, dev); arena.reset(dev)
    ga = a.take(100, dev); ga.fill_(1.0)
    b.reset(dev)                                   # the other model's zero_grad
    gb = b.take(100, dev)
    assert float(ga.sum()) == 100.0 and float(gb.sum()) == 0.0 and ga.data_ptr() != gb.data_ptr()
    # zero_grad(set_to_none=False): gradients stay alive in the arena -> nothing more is handed out until the next reset
    a.exhaust()
    assert a.take(10, dev) is None and float(ga.sum()) == 100.0
    a.reset(dev)
    assert a.take(10, dev) is not None
    # frozen after capture: a larger demand must not replace the buffer a captured graph writes to
    a.frozen = True
    ptr = a.buf.data_ptr()
    a.take(10 ** 6, dev)
    a.reset(dev)
    assert a.buf.data_ptr() == ptr


def test_drop_path_consumes_rng_like_the_reference():
    """common.py:46-59: one torch.rand of B values per DropPath call.  pmv_b200.common.drop_path_scale (what the blocks
    call branch by branch in the default mode) must leave the generator in the same state and give the same masks."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference tree not present")
    ref_common = __import__("sys").modules.get("slowfast.models.common") or (ref_loader.load_attention() and __import__("sys").modules["slowfast.models.common"])
    from pmv_b200.common import drop_path_scale
    x = torch.ones(8, 5, 3)
    rates = [0.0, 0.1, 0.1, 0.2]
    torch.manual_seed(7)
    want = [ref_common.drop_path(x, r, True) for r in rates]
    tail_ref = torch.rand(3)
    torch.manual_seed(7)
    got = []
    for r in rates:
        s = drop_path_scale(8, r, True, x.device)
        got.append(x if s is None else x * s.view(8, 1, 1))
    tail = torch.rand(3)
    for a_, b_ in zip(want, got):
        assert torch.equal(a_, b_)
    assert torch.equal(tail, tail_ref)


def test_reference_mvit_builds_with_the_block_swapped():
    """The drop-in seam: the reference's own video_model_builder.MViT constructed with
    ``MultiScaleBlock`` replaced by pmv_b200's has the reference's state_dict (397 tensors, same shapes) and loads the
    reference's weights strictly.  Runs in a subprocess: the full-model loader stubs third-party imports process-wide."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference tree not present")
    import subprocess
    import sys
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
from oracle import ref_loader
ref, cfg = ref_loader.load_full_model()
vmb = sys.modules["slowfast.models.video_model_builder"]
import pmv_b200.attention as ours
vmb.MultiScaleBlock = ours.MultiScaleBlock            # the three-line patch of INTEGRATION.md
swapped = vmb.MViT(cfg)
a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
b = {k: tuple(v.shape) for k, v in swapped.state_dict().items()}
assert a == b, sorted(set(a) ^ set(b))[:5]
assert len(a) == 397 and sum(p.numel() for p in swapped.parameters()) == 34537744
swapped.load_state_dict(ref.state_dict(), strict=True)
assert all(type(blk) is ours.MultiScaleBlock for blk in swapped.blocks)
assert [blk.dim_out for blk in swapped.blocks] == [blk.dim_out for blk in ref.blocks]
print("SWAP_OK", len(a))
""" % (ROOT, os.path.join(ROOT, "portrait-mode-video_b200"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SWAP_OK 397" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
