"""tcgen05 GEMM at the shapes bench.py actually runs (BASELINE.json configs[1]: B=256, T=32, V=128, d=128, F=512 =>
33 024 event tokens x 4 224 features, 8 448 time tokens x 16 512 features).

The small-shape kernel tests give every CTA one tile; these give each persistent CTA (or CTA pair) several, so the TMEM
accumulator slot reuse / mbarrier phase flips, the cross-tile side-tensor and bias prefetch, the one-tile-ahead row
constants, M-fastest rasterisation, CTA pairs at M = 33 024 and the split-K heuristic of the dW GEMMs are all reached.

Checker (test infrastructure only): the torch contract of tests/ops_emulator.py evaluated in fp32 on the GPU
(torch.matmul with TF32 off) and, for a subset, the independent FFMA kernel (force_simt).  Besides the whole-matrix
relative error every 128 x 256 output tile is checked on its own, so a fault that corrupts "every third tile" cannot
hide in the global norm.  `test_injected_accumulator_fault_is_caught` proves it: DX_GEMM_FAULT=1 makes the MMA issuer
skip the accumulator reset from the third tile of a CTA on (a slot-reuse bug) and the same checks must fail.
"""
import os

import pytest
import torch

import ops_emulator as E

pytestmark = pytest.mark.gpu

BF = torch.bfloat16
EV, EVD = 33024, 4224        # event axis: tokens, features
TM, TMD = 8448, 16512        # time axis
F, D = 512, 128
TOL_BF16_OUT = 4e-3          # one bf16 rounding of an fp32-accumulated value: 2^-9/sqrt(3) = 1.1e-3 relative L2 (north star: 2e-2)
TOL_F32_OUT = 2e-5           # fp32 accumulation order only (north star: 1e-3)


@pytest.fixture(scope="module")
def ops():
    from multimodal_edema_prediction_b200 import ops as o
    assert o.L.lib().dx_device_ok() == 1, "not an sm_100 device"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return o


def rnd(*shape, dtype=BF, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed * 7919 + sum(shape))
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def urnd(n, seed, lo=0.5):
    g = torch.Generator(device="cuda").manual_seed(seed * 104729 + n)
    return torch.rand(n, device="cuda", generator=g) + lo


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def tile_rel(got, ref, bm=128, bn=256):
    """max over the bm x bn output tiles of ||got-ref|| / rms-scaled ||ref|| of that tile."""
    M, N = ref.shape
    d = (got.float() - ref.float()) ** 2
    r = ref.float() ** 2
    pm, pn = (-M) % bm, (-N) % bn
    d = torch.nn.functional.pad(d, (0, pn, 0, pm)).reshape((M + pm) // bm, bm, (N + pn) // bn, bn).sum((1, 3))
    r = torch.nn.functional.pad(r, (0, pn, 0, pm)).reshape((M + pm) // bm, bm, (N + pn) // bn, bn).sum((1, 3))
    floor = r.mean() * 1e-3           # nearly-empty edge tiles
    return float((d / (r + floor)).sqrt().max())


def run_case(ops, name, *, force_simt=False):
    """Builds the operands of one production GEMM, runs the CUDA kernel and the fp32 torch contract -> dict of
    (got, ref, tolerance) per output."""
    s = 0.02
    kw, outs = {}, {}
    if name == "qkv_event":                      # qkv = (x @ Wqkv^T) * s_a            33024 x 384 x 4224 (128x192 pair tiles)
        a, b = rnd(EV, EVD, seed=1), rnd(3 * D, EVD, seed=2, scale=s)
        kw = dict(row_scale=urnd(EV, 1))
        M, N = EV, 3 * D
    elif name == "ffn_in_event":                 # h = gelu((x1 @ W1^T) s_f + b1)       33024 x 512 x 4224 (CTA pair, 258 tiles)
        a, b = rnd(EV, EVD, seed=3), rnd(F, EVD, seed=4, scale=s)
        kw = dict(row_scale=urnd(EV, 2), bias=rnd(F, dtype=torch.float32, seed=5), act=ops.ACT_GELU)
        M, N = EV, F
        outs["out2"] = True
    elif name == "ffn_in_time":                  #                                       8448 x 512 x 16512 (66 pair tiles, K=16512)
        a, b = rnd(TM, TMD, seed=6), rnd(F, TMD, seed=7, scale=s / 2)
        kw = dict(row_scale=urnd(TM, 3), bias=rnd(F, dtype=torch.float32, seed=8), act=ops.ACT_GELU)
        M, N = TM, F
        outs["out2"] = True
    elif name == "ffn_out_event":                # x2 = x1 + h @ W2^T + b2 (+ row norms) 33024 x 4224 x 512 (HBM-bound, 17 N tiles)
        a, b = rnd(EV, F, seed=9), rnd(EVD, F, seed=10, scale=0.05)
        kw = dict(bias=rnd(EVD, dtype=torch.float32, seed=11), res=rnd(EV, EVD, seed=12))
        M, N = EV, EVD
        outs["rowsq"] = True
    elif name == "outproj_time":                 # x1 = x + o @ Wo^T (+ row norms)       8448 x 16512 x 128
        a, b = rnd(TM, D, seed=13), rnd(TMD, D, seed=14, scale=0.1)
        kw = dict(res=rnd(TM, TMD, seed=15))
        M, N = TM, TMD
        outs["rowsq"] = True
    elif name == "dx_event":                     # dx1 = dx2 + dfs @ W1 - x1 * coef      33024 x 4224 x 512, B MN-major
        a, b = rnd(EV, F, seed=16), rnd(F, EVD, seed=17, scale=0.05)
        kw = dict(b_mn=True, res=rnd(EV, EVD, seed=18), cx=rnd(EV, EVD, seed=19), coef_num=rnd(EV, dtype=torch.float32, seed=20),
                  coef_den=urnd(EV, 4, 1.0))
        M, N = EV, EVD
    elif name == "dx_time_qkv":                  # dx = dx1 + dqkv @ Wqkv - x * coef      8448 x 16512 x 384, B MN-major
        a, b = rnd(TM, 3 * D, seed=21), rnd(3 * D, TMD, seed=22, scale=0.05)
        kw = dict(b_mn=True, res=rnd(TM, TMD, seed=23), cx=rnd(TM, TMD, seed=24), coef_num=rnd(TM, dtype=torch.float32, seed=25),
                  coef_den=urnd(TM, 5, 1.0))
        M, N = TM, TMD
    elif name == "gelu_bwd_event":               # dfs, df = GELU'(pre) (dx2 @ W2)        33024 x 512 x 4224, B MN-major
        a, b = rnd(EV, EVD, seed=26), rnd(EVD, F, seed=27, scale=s)
        kw = dict(b_mn=True, act=ops.ACT_GELU_BWD, aux=rnd(EV, F, seed=28), aux_bias=rnd(F, dtype=torch.float32, seed=29),
                  row_scale2=urnd(EV, 6))
        M, N = EV, F
        outs["out2"] = outs["rowdot"] = True
    elif name == "do_time":                      # do = dx1 @ Wo                           8448 x 128 x 16512, B MN-major
        a, b = rnd(TM, TMD, seed=30), rnd(TMD, D, seed=31, scale=s / 2)
        kw = dict(b_mn=True)
        M, N = TM, D
    elif name == "dw1_time":                     # dW1 += dfs^T @ x1                       512 x 16512 x 8448, both MN-major, f32 acc
        a, b = rnd(TM, F, seed=32), rnd(TM, TMD, seed=33, scale=s)
        kw = dict(a_mn=True, b_mn=True, accumulate=True)
        M, N = F, TMD
        outs["f32"] = True
    elif name == "dw2_time":                     # dW2 += dx2^T @ h                        16512 x 512 x 8448 (CTA pair / split-K)
        a, b = rnd(TM, TMD, seed=34), rnd(TM, F, seed=35, scale=s)
        kw = dict(a_mn=True, b_mn=True, accumulate=True)
        M, N = TMD, F
        outs["f32"] = True
    elif name == "dwqkv_event":                  # dWqkv += dqkv^T @ x                     384 x 4224 x 33024
        a, b = rnd(EV, 3 * D, seed=36), rnd(EV, EVD, seed=37, scale=s)
        kw = dict(a_mn=True, b_mn=True, accumulate=True)
        M, N = 3 * D, EVD
        outs["f32"] = True
    elif name == "dwqkv_time":                   # dWqkv += dqkv^T @ x                     384 x 16512 x 8448 (split-K: 195 tiles)
        a, b = rnd(TM, 3 * D, seed=38), rnd(TM, TMD, seed=39, scale=s)
        kw = dict(a_mn=True, b_mn=True, accumulate=True)
        M, N = 3 * D, TMD
        outs["f32"] = True
    else:
        raise KeyError(name)
    f32 = outs.get("f32", False)
    odt = torch.float32 if f32 else BF

    def alloc():
        o = {"out": torch.full((M, N), 0.5, device="cuda", dtype=odt) if f32 else torch.empty(M, N, device="cuda", dtype=odt)}
        if outs.get("out2"):
            o["out2"] = torch.empty(M, N, device="cuda", dtype=BF)
        if outs.get("rowsq"):
            o["row_sumsq"] = torch.zeros(M, device="cuda")
        if outs.get("rowdot"):
            o["row_dot"] = torch.zeros(M, device="cuda")
        return o

    got = alloc()
    ops.gemm_(a, b, act_dtype=BF, force_simt=force_simt, **got, **kw)
    ref = alloc()
    ref = {k: (v.float() if k in ("out", "out2") else v) for k, v in ref.items()}
    E.gemm_(a, b, **ref, **kw)
    torch.cuda.synchronize()
    tol = TOL_F32_OUT if f32 else TOL_BF16_OUT
    res = {}
    for k in got:
        res[k] = (got[k], ref[k], tol if k in ("out", "out2") else 1e-3)
    return res


CASES = ["qkv_event", "ffn_in_event", "ffn_in_time", "ffn_out_event", "outproj_time", "dx_event", "dx_time_qkv", "gelu_bwd_event",
         "do_time", "dw1_time", "dw2_time", "dwqkv_event", "dwqkv_time"]


def check(res, name):
    for k, (g, r, tol) in res.items():
        assert torch.isfinite(g.float()).all(), (name, k)
        e = rel(g, r)
        assert e < tol, (name, k, e)
        if g.dim() == 2:
            te = tile_rel(g, r)
            assert te < 3 * tol, (name, k, "worst tile", te)


@pytest.mark.parametrize("name", CASES)
def test_production_shape_vs_fp32_contract(ops, name):
    check(run_case(ops, name), name)


@pytest.mark.parametrize("name", ["ffn_in_event", "dx_event", "dw2_time", "do_time"])
def test_production_shape_tc_vs_ffma_kernel(ops, name):
    """Two independent kernels (tcgen05 / FFMA), same bf16 operands, same epilogue contract."""
    r_tc, r_ff = run_case(ops, name), run_case(ops, name, force_simt=True)
    for k in r_tc:
        g, f, tol = r_tc[k][0], r_ff[k][0], r_tc[k][2]
        assert rel(g, f) < tol, (name, k, rel(g, f))
        if g.dim() == 2:
            assert tile_rel(g, f) < 3 * tol, (name, k)


@pytest.mark.parametrize("name", ["ffn_out_event", "ffn_in_event", "dwqkv_time"])
def test_injected_accumulator_fault_is_caught(ops, name):
    """Sensitivity of the checks above: with DX_GEMM_FAULT=1 the MMA issuer keeps accumulating onto the stale TMEM slot from
    the third tile of each persistent CTA on.  The production-shape checks must reject that result (and pass again once
    the fault is switched off), i.e. the multi-tile path really is exercised and verified."""
    os.environ["DX_GEMM_FAULT"] = "1"
    try:
        with pytest.raises(AssertionError):
            check(run_case(ops, name), name)
    finally:
        del os.environ["DX_GEMM_FAULT"]
    check(run_case(ops, name), name)
