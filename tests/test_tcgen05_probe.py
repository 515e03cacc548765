"""Pins the shared-memory / tensor-memory operand layouts the tcgen05 kernels rely on, with a single-tile
UMMA whose shared-memory image and descriptors are built on the host.  Integer-valued bf16 data, so the
fp32 result must match exactly."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SW = {128: 2, 64: 4, 32: 6, 0: 0}


def bf16_bits(a: np.ndarray) -> np.ndarray:
    return (a.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def swz(addr: np.ndarray, sw_bytes: int) -> np.ndarray:
    bits = {128: 3, 64: 2, 32: 1, 0: 0}[sw_bytes]
    return addr ^ (((addr >> 7) & ((1 << bits) - 1)) << 4)


def image_k_major(mat: np.ndarray, sw_bytes: int, base: int, img: np.ndarray):
    """mat [rows, cols] (cols = reduction axis, cols*2 == sw_bytes): rows at pitch sw_bytes, 8-row atoms."""
    rows, cols = mat.shape
    assert cols * 2 == sw_bytes
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    addr = swz(r * sw_bytes + c * 2, sw_bytes) + base
    img.view(np.uint16)[addr.ravel() // 2] = bf16_bits(mat).ravel()
    return rows * sw_bytes


def image_mn_major(mat: np.ndarray, sw_bytes: int, base: int, img: np.ndarray):
    """mat [mn, k] stored transposed: groups of (sw_bytes/2) mn-elements; inside a group row = k (pitch sw_bytes)."""
    mn, k = mat.shape
    g = sw_bytes // 2
    ngroups = (mn + g - 1) // g
    group_bytes = k * sw_bytes
    i, j = np.meshgrid(np.arange(mn), np.arange(k), indexing="ij")
    addr = (i // g) * group_bytes + swz(j * sw_bytes + (i % g) * 2, sw_bytes) + base
    img.view(np.uint16)[addr.ravel() // 2] = bf16_bits(mat).ravel()
    return ngroups * group_bytes, group_bytes


def desc(lbo, sbo, sw_bytes):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | SW[sw_bytes] << 61


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(img, da, db, a_off, b_off, idsc, nk, a_step, b_step, N, tmem_a=None):
    from pmv_b200 import _lib as L
    dimg = torch.from_numpy(img.copy()).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    ta = torch.from_numpy(np.ascontiguousarray(tmem_a).view(np.int32).copy()).cuda() if tmem_a is not None else None
    L.check(L.lib().pmv_probe_umma(dimg.data_ptr(), img.nbytes, da, db, a_off, b_off, idsc, nk, a_step, b_step,
                                   1 if tmem_a is not None else 0, ta.data_ptr() if ta is not None else None,
                                   tmem_a.shape[1] if tmem_a is not None else 0, out.data_ptr(), N,
                                   torch.cuda.current_stream().cuda_stream), "probe")
    torch.cuda.synchronize()
    return out.cpu().numpy()


RNG = np.random.default_rng(0)
RESULTS = {}


def rnd(*shape):
    return RNG.integers(-3, 4, size=shape).astype(np.float32)


def record(name, got, want):
    ok = bool(np.array_equal(got, want))
    RESULTS[name] = {"ok": ok, "max_abs_err": float(np.nanmax(np.abs(got - want))) if not np.isnan(got).all() else None}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_results.json", "w") as f:
        json.dump(RESULTS, f, indent=1)
    return ok


def test_k_major_sw128():
    K, N = 64, 128
    A, B = rnd(128, K), rnd(N, K)
    img = np.zeros(64 * 1024, np.uint8)
    a_bytes = image_k_major(A, 128, 0, img)
    image_k_major(B, 128, a_bytes, img)
    got = run(img, desc(16, 1024, 128), desc(16, 1024, 128), 0, a_bytes, idesc(128, N, 0, 0), K // 16, 32, 32, N)
    assert record("k_major_sw128", got, A @ B.T)


def test_mn_major_sw128_b():
    K, N = 64, 128
    A, B = rnd(128, K), rnd(N, K)
    img = np.zeros(64 * 1024, np.uint8)
    a_bytes = image_k_major(A, 128, 0, img)
    _, gb = image_mn_major(B, 128, a_bytes, img)
    want = A @ B.T
    got = run(img, desc(16, 1024, 128), desc(gb, 1024, 128), 0, a_bytes, idesc(128, N, 0, 1), K // 16, 32, 2048, N)
    ok = record("mn_major_sw128_b(lbo=group,sbo=1024)", got, want)
    # alternative reading of the LBO / SBO roles, recorded for diagnosis only
    alt = run(img, desc(16, 1024, 128), desc(1024, gb, 128), 0, a_bytes, idesc(128, N, 0, 1), K // 16, 32, 2048, N)
    record("mn_major_sw128_b(lbo=1024,sbo=group)", alt, want)
    assert ok


def test_mn_major_sw128_both():
    K, N = 64, 96
    A, B = rnd(128, K), rnd(N, K)
    img = np.zeros(64 * 1024, np.uint8)
    a_bytes, ga = image_mn_major(A, 128, 0, img)
    _, gb = image_mn_major(B, 128, a_bytes, img)
    got = run(img, desc(ga, 1024, 128), desc(gb, 1024, 128), 0, a_bytes, idesc(128, N, 1, 1), K // 16, 2048, 2048, N)
    assert record("mn_major_sw128_both_n96", got, A @ B.T)


def test_k_major_sw64():
    K, N = 32, 128
    A, B = rnd(128, K), rnd(N, K)
    img = np.zeros(64 * 1024, np.uint8)
    a_bytes = image_k_major(A, 64, 0, img)
    image_k_major(B, 64, a_bytes, img)
    got = run(img, desc(16, 512, 64), desc(16, 512, 64), 0, a_bytes, idesc(128, N, 0, 0), K // 16, 32, 32, N)
    assert record("k_major_sw64", got, A @ B.T)


def test_mn_major_sw64_b_n96():
    K, N = 64, 96  # V tile of the attention kernel: 64 keys x 96 channels, channel axis contiguous
    A, B = rnd(128, K), rnd(N, K)
    img = np.zeros(64 * 1024, np.uint8)
    a_bytes = image_k_major(A, 128, 0, img)
    _, gb = image_mn_major(B, 64, a_bytes, img)
    got = run(img, desc(16, 1024, 128), desc(gb, 512, 64), 0, a_bytes, idesc(128, N, 0, 1), K // 16, 32, 1024, N)
    assert record("mn_major_sw64_b_n96", got, A @ B.T)


def test_a_from_tmem():
    K, N = 64, 96  # P (128 x 64 keys) from tensor memory, V MN-major from shared memory
    A, B = rnd(128, K), rnd(N, K)
    img = np.zeros(64 * 1024, np.uint8)
    _, gb = image_mn_major(B, 64, 0, img)
    bits = bf16_bits(A).astype(np.uint32)
    packed = bits[:, 0::2] | (bits[:, 1::2] << 16)  # column c holds elements (2c, 2c+1)
    got = run(img, 0, desc(gb, 512, 64), 0, 0, idesc(128, N, 0, 1), K // 16, 8, 1024, N, tmem_a=packed)
    assert record("a_from_tmem_packed_pairs", got, A @ B.T)
