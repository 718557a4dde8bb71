"""Per-kernel parity on a real B200: every CUDA kernel wrapper (through the C ABI of libduett_b200.so) against the
torch contract in tests/ops_emulator.py, fp32 and bf16 storage.  Tolerances: fp32 1e-4 (re-association only; the
north-star bound is 1e-3), bf16 2e-2 (north-star bound)."""
import math

import numpy as np

import pytest
import torch

import ops_emulator as E

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cpu(*ts):
    return [None if t is None else t.detach().cpu().clone() for t in ts]


@pytest.fixture(scope="module")
def ops():
    from multimodal_edema_prediction_b200 import ops as o
    assert o.L.lib().dx_device_ok() == 1, "not an sm_100 device"
    return o


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_plain(ops, dt, a_mn, b_mn):
    M, N, K = 520, 392, 264
    a = rnd(*((K, M) if a_mn else (M, K)), dtype=dt, seed=1)
    b = rnd(*((K, N) if b_mn else (N, K)), dtype=dt, seed=2)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm_(a, b, a_mn=a_mn, b_mn=b_mn, out=out)
    ref = torch.empty(M, N)
    E.gemm_(*cpu(a, b), a_mn=a_mn, b_mn=b_mn, out=ref)
    assert rel(out, ref) < 1e-5      # fp32 accumulation of identical inputs


@pytest.mark.parametrize("dt", DT)
def test_gemm_fused_epilogues(ops, dt):
    M, N, K = 300, 264, 136
    a, b = rnd(M, K, dtype=dt, seed=3), rnd(N, K, dtype=dt, seed=4, scale=0.1)
    rs, rs2 = torch.rand(M, device="cuda") + 0.5, torch.rand(M, device="cuda") + 0.5
    bias = rnd(N, seed=5)
    res, aux, cx = rnd(M, N, dtype=dt, seed=6), rnd(M, N, dtype=dt, seed=7), rnd(M, N, dtype=dt, seed=8)
    num, den = rnd(M, seed=9), torch.rand(M, device="cuda") + 1
    cases = {
        "ffn_in": dict(row_scale=rs, bias=bias, act=ops.ACT_GELU, want_out2=True),
        "resid": dict(bias=bias, res=res, want_rowsq=True),
        "dx": dict(res=res, cx=cx, coef_num=num, coef_den=den),
        "gelu_bwd": dict(act=ops.ACT_GELU_BWD, aux=aux, aux_bias=bias, row_scale2=rs2, want_out2=True, want_rowdot=True),
        "relu": dict(bias=bias, act=ops.ACT_RELU),
        "tanh": dict(bias=bias, act=ops.ACT_TANH),
    }
    for name, kw in cases.items():
        kw = dict(kw)
        o2, rq, rd = kw.pop("want_out2", False), kw.pop("want_rowsq", False), kw.pop("want_rowdot", False)
        out = torch.empty(M, N, device="cuda", dtype=dt)
        out2 = torch.empty(M, N, device="cuda", dtype=dt) if o2 else None
        rowsq = torch.zeros(M, device="cuda") if rq else None
        rowdot = torch.zeros(M, device="cuda") if rd else None
        ops.gemm_(a, b, out=out, out2=out2, row_sumsq=rowsq, row_dot=rowdot, act_dtype=dt, **kw)
        r_out, r_out2 = torch.empty(M, N), (torch.empty(M, N) if o2 else None)
        r_rowsq, r_rowdot = (torch.zeros(M) if rq else None), (torch.zeros(M) if rd else None)
        ckw = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in kw.items()}
        E.gemm_(*cpu(a, b), out=r_out, out2=r_out2, row_sumsq=r_rowsq, row_dot=r_rowdot, **ckw)
        assert rel(out, r_out) < TOL[dt], name
        if o2:
            assert rel(out2, r_out2) < TOL[dt], name
        if rq:
            assert rel(rowsq, r_rowsq) < 1e-4, name
        if rd:
            assert rel(rowdot, r_rowdot) < 1e-3, name


def test_gemm_tc_matches_ffma_on_bf16(ops):
    """The tcgen05 kernel and the FFMA kernel are independent implementations: same bf16 inputs, same fp32 result."""
    M, N, K = 1000, 520, 1096
    a, b = rnd(M, K, dtype=torch.bfloat16, seed=11), rnd(N, K, dtype=torch.bfloat16, seed=12)
    o1 = torch.empty(M, N, device="cuda")
    o2 = torch.empty(M, N, device="cuda")
    ops.gemm_(a, b, out=o1)
    ops.gemm_(a, b, out=o2, force_simt=True)
    assert rel(o1, o2) < 1e-5
    acc = torch.full((M, N), 2.0, device="cuda")
    ops.gemm_(a, b, out=acc, accumulate=True)
    assert rel(acc - 2.0, o1) < 1e-5


@pytest.mark.parametrize("dt", DT)
def test_gemm_grouped_strided(ops, dt):
    """Grouped mode on the strided views the embedding uses: group g reads psi[:, g, :] (row stride G*d, group stride d)."""
    G, R, d, H = 5, 300, 24, 64
    psi = rnd(R, G, d, dtype=dt, seed=13)                       # [rows, group, d]
    a = psi.permute(1, 0, 2)                                    # [G, R, d] view: stride (d, G*d, 1)
    w = rnd(G, H, d, dtype=dt, seed=14, scale=0.3)              # per-group weight [H, d]
    bias = rnd(G, H, seed=15)
    out = torch.empty(G, R, H, device="cuda", dtype=dt)
    ops.gemm_(a, w, out=out, bias=bias, act_dtype=dt)
    ref = torch.einsum("grd,ghd->grh", a.float(), w.float()) + bias[:, None, :]
    assert rel(out, ref) < TOL[dt]
    # dW_g += A_g^T @ B_g with both operands MN-major, fp32 accumulate, strided A
    hn = rnd(G, R, H, dtype=dt, seed=16)
    dw = torch.ones(G, d, H, device="cuda")
    ops.gemm_(a, hn, a_mn=True, b_mn=True, out=dw, accumulate=True)
    assert rel(dw - 1.0, torch.einsum("grd,grh->gdh", a.float(), hn.float())) < 1e-4
    # output written through a strided view (scatter into psi-like layout)
    outp = torch.zeros(R, G, 32, device="cuda", dtype=dt)
    w2 = rnd(G, 32, H, dtype=dt, seed=17, scale=0.3)
    ops.gemm_(hn, w2, out=outp.permute(1, 0, 2), act_dtype=dt)
    assert rel(outp.permute(1, 0, 2), torch.einsum("grh,gdh->grd", hn.float(), w2.float())) < TOL[dt]


def test_gemm_rejects_bad_arguments(ops):
    a, b = rnd(64, 20, dtype=torch.bfloat16), rnd(32, 20, dtype=torch.bfloat16)     # lda = 20 is not 16 B aligned
    with pytest.raises(ops.L.DxError):
        ops.gemm_(a, b, out=torch.empty(64, 32, device="cuda"))
    with pytest.raises(ops.L.DxError):
        ops.gemm_(rnd(8, 4), rnd(8, 5), out=torch.empty(8, 8, device="cuda"))       # contraction mismatch


@pytest.mark.parametrize("dt", DT)
def test_relayout_fwd_bwd(ops, dt):
    B, P, Q, d = 3, 5, 7, 16
    src = rnd(B, P, Q, d, dtype=dt, seed=20)
    rowsq = (src.float() ** 2).sum((2, 3)).reshape(-1).contiguous()
    g = torch.tensor([1.3], device="cuda")
    pos_b, pos_n = rnd(Q, P * d, seed=21), rnd(B, Q, P * d, dtype=dt, seed=22)
    for kw in (dict(), dict(src_rowsq=rowsq, g=g, pos_bcast=pos_b), dict(src_rowsq=rowsq, g=g, pos_batched=pos_n)):
        dst, rq = ops.relayout_fwd(src, B, P, Q, d, **kw)
        ckw = {k: v.cpu() for k, v in kw.items()}
        rdst, rrq = E.relayout_fwd(src.cpu(), B, P, Q, d, **ckw)
        assert rel(dst, rdst) < TOL[dt] and rel(rq, rrq) < TOL[dt]
    gd = rnd(B, Q, P, d, dtype=dt, seed=23)
    dg = torch.zeros(1, device="cuda")
    ds = ops.relayout_bwd(gd, B, P, Q, d, src=src, src_rowsq=rowsq, g=g, dg=dg)
    rdg = torch.zeros(1)
    rds = E.relayout_bwd(gd.cpu(), B, P, Q, d, src=src.cpu(), src_rowsq=rowsq.cpu(), g=g.cpu(), dg=rdg)
    assert rel(ds, rds) < TOL[dt] and rel(dg, rdg) < TOL[dt]
    assert rel(ops.relayout_bwd(gd, B, P, Q, d), E.relayout_bwd(gd.cpu(), B, P, Q, d)) == 0.0     # pure permutation: exact


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("S,D,H", [(129, 128, 2), (33, 64, 2), (5, 8, 2), (35, 24, 2), (70, 256, 2),
                                   (513, 256, 2), (129, 256, 2), (200, 128, 2), (300, 512, 4)])     # tiled kernels: dh 128 / S > 144
def test_attention_fwd_bwd(ops, dt, S, D, H):
    B = 3
    qkv = rnd(B, S, 3 * D, dtype=dt, seed=30, scale=0.7)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    o, lse = ops.attn_fwd(q, k, v, H)
    ro, rlse = E.attn_fwd(*cpu(q, k, v), H)
    assert rel(o, ro) < TOL[dt] and rel(lse, rlse) < 1e-3
    go = rnd(B, S, D, dtype=dt, seed=31)
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(q, k, v, o, go, lse, H, dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:])
    rq, rk, rv = torch.empty(B, S, D), torch.empty(B, S, D), torch.empty(B, S, D)
    E.attn_bwd(*cpu(q, k, v, o, go, lse), H, rq, rk, rv)
    assert rel(dqkv, torch.cat([rq, rk, rv], 2)) < TOL[dt]


@pytest.mark.parametrize("B,Sq,Sk,D,H,p", [(3, 129, 129, 128, 2, 0.0), (1, 64, 64, 64, 1, 0.0), (2, 128, 128, 128, 2, 0.2),
                                          (2, 144, 144, 192, 3, 0.0), (2, 256, 256, 128, 2, 0.1), (2, 130, 77, 128, 2, 0.0),
                                          (2, 300, 17, 64, 1, 0.3), (5, 193, 193, 128, 2, 0.0), (2, 129, 129, 128, 2, 0.25),
                                          (2, 257, 96, 64, 1, 0.1), (3, 129, 33, 128, 2, 0.0)])
def test_attention_tcgen05_forward(ops, monkeypatch, B, Sq, Sk, D, H, p):
    """tcgen05 / TMEM forward (dx_attention_tc.cu: dh = 64, Sk <= 256, Sq >= 64 — the event axis): against the emulator and
    against the mma.sync kernel it replaces (DX_ATTN_TC=0) on the same dropout mask; tile edges (Sq = 128k, and 128k + 1 where
    the last row runs on the CTA's CUDA-core warp), register-resident (Sk <= 160) and two-pass softmax, key counts that are not
    multiples of 16 / 32 / 64, cross attention, a single batch row."""
    bf = torch.bfloat16
    drop = (p, 24680) if p else None
    if Sq == Sk:
        qkv = rnd(B, Sq, 3 * D, dtype=bf, seed=130, scale=0.7)          # packed projections, strided head views
        q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    else:
        q, k, v = rnd(B, Sq, D, dtype=bf, seed=131, scale=0.7), rnd(B, Sk, D, dtype=bf, seed=132, scale=0.7), rnd(B, Sk, D, dtype=bf, seed=133)
    args = (q, k, v, H) + ((drop,) if drop else ())
    o, lse = ops.attn_fwd(*args)
    ro, rlse = E.attn_fwd(*cpu(q, k, v), H, *((drop,) if drop else ()))
    assert rel(o, ro) < TOL[bf] * (2 if p else 1) and rel(lse, rlse) < 1e-3
    monkeypatch.setenv("DX_ATTN_TC", "0")
    o2, lse2 = ops.attn_fwd(*args)
    monkeypatch.delenv("DX_ATTN_TC")
    assert rel(o, o2) < TOL[bf] and rel(lse, lse2) < 1e-4
    # the mma.sync backward consumes the tcgen05 forward's o / lse and regenerates the same mask
    go = rnd(B, Sq, D, dtype=bf, seed=134)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attn_bwd(q, k, v, o, go, lse, H, dq, dk, dv, *((drop,) if drop else ()))
    rq, rk, rv = torch.empty(B, Sq, D), torch.empty(B, Sk, D), torch.empty(B, Sk, D)
    E.attn_bwd(*cpu(q, k, v, o, go, lse), H, rq, rk, rv, *((drop,) if drop else ()))
    assert rel(dq, rq) < TOL[bf] * 2 and rel(dk, rk) < TOL[bf] * 2 and rel(dv, rv) < TOL[bf] * 2


def test_cross_attention_few_queries(ops):
    B, Sq, Sk, D, H = 2, 7, 200, 256, 4
    q, k, v = rnd(B, Sq, D, seed=33), rnd(B, Sk, D, seed=34), rnd(B, Sk, D, seed=35)
    o, lse = ops.attn_fwd(q, k, v, H)
    ro, _ = E.attn_fwd(*cpu(q, k, v), H)
    assert rel(o, ro) < 1e-4
    go = rnd(B, Sq, D, seed=36)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attn_bwd(q, k, v, o, go, lse, H, dq, dk, dv)
    rq, rk, rv = torch.empty(B, Sq, D), torch.empty(B, Sk, D), torch.empty(B, Sk, D)
    E.attn_bwd(*cpu(q, k, v, o, go, lse), H, rq, rk, rv)
    assert rel(dq, rq) < 1e-4 and rel(dk, rk) < 1e-4 and rel(dv, rv) < 1e-4


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,Sq,Sk,D,H", [(2, 7, 1369, 256, 4), (3, 7, 24, 32, 4), (1, 129, 129, 128, 2), (2, 4, 130, 48, 4)])
def test_attention_probs_mean(ops, dt, B, Sq, Sk, D, H):
    """dx_attn_probs_mean (head-averaged attention maps of nn.MultiheadAttention(need_weights=True)) from the lse of
    dx_attn_fwd: the perceiver's shapes (7 queries over 1369 patches / 24 hours), strided k views, a key count that is not a
    multiple of the 128-thread block; rows sum to one."""
    q = rnd(B, Sq, D, dtype=dt, seed=37, scale=0.7)
    kv = rnd(B, Sk, 2 * D, dtype=dt, seed=38, scale=0.7)
    k, v = kv[:, :, :D], kv[:, :, D:]
    o, lse = ops.attn_fwd(q, k, v, H)
    w = ops.attn_probs_mean(q, k, lse, H)
    assert w.shape == (B, Sq, Sk) and w.dtype == torch.float32
    rw = E.attn_probs_mean(*cpu(q, k), E.attn_fwd(*cpu(q, k, v), H)[1], H)
    assert rel(w, rw) < (1e-4 if dt == torch.float32 else 5e-3)
    assert float((w.sum(-1) - 1).abs().max()) < (1e-4 if dt == torch.float32 else 5e-3)


def _embed_inputs(B, T, V, d, seed=40, ssl=True):
    g = torch.Generator().manual_seed(seed)
    obs = (torch.rand(B, T, V, generator=g) < 0.3).float()
    xs = torch.cat((torch.randn(B, T, V, generator=g) * obs, obs * torch.randint(1, 20, (B, T, V), generator=g).float(),
                    torch.zeros(B, T, 1)), 2)
    if ssl:
        for b in range(B):
            xs[b, b % T, :] = 0; xs[b, b % T, -1] = 1
            xs[b, :, (3 * b) % V] = 0; xs[b, :, V + (3 * b) % V] = -1
    P = dict(W0=torch.randn(V, 64, 2, generator=g) * 0.7, b0=torch.randn(V, 64, generator=g) * 0.3,
             gamma=1 + 0.1 * torch.randn(V, 64, generator=g), beta=0.1 * torch.randn(V, 64, generator=g),
             W4=torch.randn(V, d, 64, generator=g) * 0.12, b4=torch.randn(V, d, generator=g) * 0.1,
             nobs=torch.randn(16, generator=g), special=torch.randn(8, d, generator=g), tab=torch.randn(B, d, generator=g))
    return xs, P


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("B,T,V,d", [(6, 4, 5, 8), (16, 32, 34, 24), (8, 24, 40, 128)])
def test_embedding_fwd_bwd(ops, dt, B, T, V, d):
    xs, P = _embed_inputs(B, T, V, d)
    c = lambda t: t.cuda().contiguous()
    for training in (True, False):
        rm, rv = torch.zeros(V, 64) + 0.05, torch.ones(V, 64) * 0.9
        rmc, rvc = c(rm), c(rv)
        psi, mean, rstd = ops.embed_fwd(c(xs), V, d, c(P["W0"]), c(P["b0"]), c(P["gamma"]), c(P["beta"]), rmc, rvc, c(P["W4"]),
                                        c(P["b4"]), c(P["nobs"]), c(P["special"]), c(P["tab"]), dt, training)
        rpsi, rmean, rrstd = E.embed_fwd(xs, V, d, P["W0"], P["b0"], P["gamma"], P["beta"], rm, rv, P["W4"], P["b4"], P["nobs"],
                                         P["special"], P["tab"], dt, training)
        assert rel(psi, rpsi) < TOL[dt], training
        assert rel(mean, rmean) < 1e-4 and rel(rstd, rrstd) < 1e-3
        assert rel(rmc, rm) < 1e-4 and rel(rvc, rv) < 1e-3                      # running statistics
        dpsi = rnd(B, T + 1, V + 1, d, dtype=dt, seed=41)
        names = dict(dW0=P["W0"], db0=P["b0"], dgamma=P["gamma"], dbeta=P["beta"], dW4=P["W4"], db4=P["b4"], dnobs=P["nobs"],
                     dspecial=P["special"])
        G = {k: torch.zeros_like(v).cuda() for k, v in names.items()}
        RG = {k: torch.zeros_like(v) for k, v in names.items()}
        dtab = ops.embed_bwd(c(xs), V, d, c(P["W0"]), c(P["b0"]), c(P["gamma"]), c(P["beta"]), c(P["W4"]), c(P["nobs"]), mean, rstd,
                             dpsi.clone(), G, training)      # the op consumes dpsi (zeroes its special cells in place)
        rdtab = E.embed_bwd(xs, V, d, P["W0"], P["b0"], P["gamma"], P["beta"], P["W4"], P["nobs"], rmean, rrstd, dpsi.cpu(), RG,
                            training)
        assert rel(dtab, rdtab) < TOL[dt]
        for k in G:
            assert rel(G[k], RG[k]) < (2e-3 if dt == torch.float32 else 2e-2), (k, training)   # fp32: atomics re-association


def test_norm_kernels(ops):
    R, C = 300, 90
    x, w, b = rnd(R, C, seed=50), rnd(C, seed=51), rnd(C, seed=52)
    for training in (True, False):
        rm, rv = torch.zeros(C, device="cuda") + 0.1, torch.ones(C, device="cuda")
        crm, crv = rm.cpu().clone(), rv.cpu().clone()
        y, mean, rstd = ops.bn2d_fwd(x, w, b, rm, rv, training)
        ry, rmean, rrstd = E.bn2d_fwd(x.cpu(), w.cpu(), b.cpu(), crm, crv, training)
        assert rel(y, ry) < 1e-4 and rel(rm, crm) < 1e-5 and rel(rv, crv) < 1e-4
        dy = rnd(R, C, seed=53)
        dw, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        dx = ops.bn2d_bwd(dy, x, w, mean, rstd, dw, db, training)
        rdw, rdb = torch.zeros(C), torch.zeros(C)
        rdx = E.bn2d_bwd(dy.cpu(), x.cpu(), w.cpu(), rmean, rrstd, rdw, rdb, training)
        assert rel(dx, rdx) < 1e-4 and rel(dw, rdw) < 1e-4 and rel(db, rdb) < 1e-4
    for dt in DT:
        x = rnd(500, 256, dtype=dt, seed=54)
        w, b = rnd(256, seed=55), rnd(256, seed=56)
        y, mean, rstd = ops.layernorm_fwd(x, w, b)
        ry, rmean, rrstd = E.layernorm_fwd(x.cpu(), w.cpu(), b.cpu())
        assert rel(y, ry) < TOL[dt]
        dy = rnd(500, 256, dtype=dt, seed=57)
        dw, db = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
        dx = ops.layernorm_bwd(dy, x, w, mean, rstd, dw, db)
        rdw, rdb = torch.zeros(256), torch.zeros(256)
        rdx = E.layernorm_bwd(dy.cpu(), x.cpu(), w.cpu(), rmean, rrstd, rdw, rdb)
        assert rel(dx, rdx) < TOL[dt] and rel(dw, rdw) < TOL[dt] and rel(db, rdb) < TOL[dt]


def test_small_kernels(ops):
    x = rnd(37, 1000, dtype=torch.bfloat16, seed=60)
    out = torch.zeros(1000, device="cuda")
    ops.colsum(x, out, accumulate=True)
    assert rel(out, x.float().sum(0)) < 1e-5
    x3 = rnd(4, 9, 64, dtype=torch.bfloat16, seed=61)
    assert rel(ops.mean_rows(x3, 8), E.mean_rows(x3.cpu(), 8)) < 1e-6
    dy = rnd(4, 64, seed=62)
    assert rel(ops.mean_rows_bwd(dy, 9, 8, torch.float32), E.mean_rows_bwd(dy.cpu(), 9, 8, torch.float32)) < 1e-6
    off = torch.tensor([0, 128, 320], device="cuda")
    assert torch.equal(ops.gather_vec(x3, off, 64).cpu(), E.gather_vec(x3.cpu(), off.cpu(), 64))
    g, aux = rnd(100, 40, seed=63), rnd(100, 40, seed=64)
    for code in (E.ACT_GELU_BWD, E.ACT_RELU_BWD, E.ACT_TANH_BWD):
        assert rel(ops.act_bwd(g, aux, code), E.act_bwd(g.cpu(), aux.cpu(), code)) < 1e-5
    rowsq, gg = torch.rand(50, device="cuda") + 0.1, torch.tensor([0.7], device="cuda")
    assert rel(ops.scalenorm_scale(rowsq, gg, 96), E.scalenorm_scale(rowsq.cpu(), gg.cpu(), 96)) < 1e-6
    a, gr = rnd(50, 24, seed=65), rnd(50, 24, seed=66)
    gr2 = gr.clone()
    rd = ops.rowdot_scale(a, gr2, rowsq)
    cgr = gr.cpu().clone()
    rrd = E.rowdot_scale(a.cpu(), cgr, rowsq.cpu())
    assert rel(rd, rrd) < 1e-5 and rel(gr2, cgr) < 1e-6


def test_gather_scatter_negative_offsets(ops):
    """A negative offset selects nothing (the zero-padded rows of pretrain_masked_steps > 1): zero row out of the gather,
    nothing written by the scatter."""
    for dt in DT:
        x = rnd(5, 7, 64, dtype=dt, seed=67)
        off = torch.tensor([64, -1, 0, 1984, -1, 640], device="cuda")
        g = ops.gather_vec(x, off, 64)
        assert torch.equal(g.cpu(), E.gather_vec(x.cpu(), off.cpu(), 64)) and float(g[1].abs().max()) == 0.0
        src = rnd(6, 64, seed=68)
        dst, rdst = torch.zeros_like(x), torch.zeros(5, 7, 64, dtype=dt)
        ops.scatter_vec(src, off, dst, accumulate=False)
        E.scatter_vec(src.cpu(), off.cpu(), rdst, accumulate=False)
        assert torch.equal(dst.cpu(), rdst)
        ops.scatter_vec(src, off, dst, accumulate=True)
        E.scatter_vec(src.cpu(), off.cpu(), rdst, accumulate=True)
        assert torch.equal(dst.cpu(), rdst)


@pytest.mark.parametrize("K", [1, 2, 3])
@pytest.mark.parametrize("with_ev,with_keep", [(True, True), (False, True), (True, False)])
def test_ssl_mask_kernel(ops, K, with_ev, with_keep):
    """dx_ssl_mask against the reference's index chain (emulator), bit-exact: one or several masked timesteps per sample
    (repeats included), masked variable, variable dropout, NaN cells and zero / fractional / multiple counts."""
    B, T, V = 9, 6, 7
    g = torch.Generator().manual_seed(70 + K)
    vals = torch.randn(B, T, V, generator=g)
    cnts = torch.randint(0, 4, (B, T, V), generator=g).float()
    cnts[0, 0, 0], vals[1, 2, 3] = 0.5, float("nan")
    xs = torch.cat([vals * (cnts > 0), cnts, torch.zeros(B, T, 1)], 2)
    step = torch.randint(0, T, (B, K), generator=g, dtype=torch.int32)
    if K > 1:
        step[0, 1] = step[0, 0]                                    # a repeated draw
    ev = torch.randint(0, V, (B,), generator=g, dtype=torch.int32) if with_ev else None
    keep = (torch.rand(B, V, generator=g) > 0.5).to(torch.uint8) if with_keep else None
    st = step[:, 0].contiguous() if K == 1 else step
    want = E.ssl_mask(xs, st, ev, keep)
    got = ops.ssl_mask(xs.cuda(), st.cuda(), None if ev is None else ev.cuda(), None if keep is None else keep.cuda())
    assert got[1].shape == ((B, V) if K == 1 else (B, K, V))
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(torch.nan_to_num(a.cpu(), nan=12345.0), torch.nan_to_num(b, nan=12345.0))


def test_loss_kernels(ops):
    B, K = 37, 7
    g = torch.Generator().manual_seed(70)
    zs, zt = torch.randn(B, generator=g) * 2, torch.randn(B, generator=g) * 2
    y = (torch.rand(B, generator=g) < 0.3).float()
    for pw in (None, 2.5):
        out, dz = ops.kd_loss(zs.cuda(), zt.cuda(), y.cuda(), 4.0, 0.5, pw)
        rout, rdz = E.kd_loss(zs, zt, y, 4.0, 0.5, pw)
        assert rel(out, rout) < 1e-5 and rel(dz, rdz) < 1e-4
    out, dz = ops.bce_logits(zs.cuda(), y.cuda(), 1.7, 0.6)
    rout, rdz = E.bce_logits(zs, y, 1.7, 0.6)
    assert rel(out, rout) < 1e-5 and rel(dz, rdz) < 1e-5
    yh, ph, yy = torch.randn(B, K, generator=g), torch.randn(B, K, generator=g), torch.randn(B, K, generator=g)
    m = (torch.rand(B, K, generator=g) < 0.4).float()
    o2, ro2 = torch.zeros(2, device="cuda"), torch.zeros(2)
    d1, d2 = ops.masked_mse_bce(yh.cuda(), ph.cuda(), yy.cuda(), m.cuda(), 0.2, o2)
    r1, r2 = E.masked_mse_bce(yh, ph, yy, m, 0.2, ro2)
    assert rel(o2, ro2) < 1e-5 and rel(d1, r1) < 1e-5 and rel(d2, r2) < 1e-5
    pwv = torch.rand(K, generator=g) + 1
    per, dz = ops.masked_bce_cols(yh.cuda(), y[:, None].expand(B, K).contiguous().cuda(), m.cuda(), pwv.cuda(), None, 1e-6)
    rper, rdz = E.masked_bce_cols(yh, y[:, None].expand(B, K).contiguous(), m, pwv, None, 1e-6)
    assert rel(per, rper) < 1e-5 and rel(dz, rdz) < 1e-5
    ym = (torch.rand(B, K, generator=g) < 0.2).float()
    out, dc = ops.aux_residual_kl(yh.cuda(), ph.cuda(), ym.cuda(), m.cuda(), 0.05)
    rout, rdc = E.aux_residual_kl(yh, ph, ym, m, 0.05)
    assert rel(out, rout) < 1e-5 and rel(dc, rdc) < 1e-4


def test_adamw_matches_torch(ops):
    n = 10007
    p = rnd(n, seed=80); g = rnd(n, seed=81)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        ref.grad = g.clone() * step
        opt.step()
        ops.adamw(p, g * step, m, v, 1e-3, (0.9, 0.999), 1e-8, 0.05, step)
    assert rel(p, ref) < 1e-6
    ss = torch.zeros(1, device="cuda")
    ops.sumsq(g, ss)
    assert rel(ss, (g * g).sum()) < 1e-5
    clip = torch.zeros(1, device="cuda")
    ops.clip_factor(ss, 1.0, clip)
    assert rel(clip, torch.clamp(1.0 / (g.norm() + 1e-6), max=1.0)) < 1e-5


@pytest.mark.parametrize("n", [1, 7, 1000, 4096, 5000])
def test_binary_auc_vs_sklearn(ops, n):
    """dx_binary_auc against sklearn (the reference's scorer, training_duett/evaluator.py:22-35): heavy ties, raw-score mode
    is exact to float64 rounding; sigmoid mode ranks fp32 sigmoid(logits) like the reference."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    g = torch.Generator().manual_seed(n)
    z = torch.randn(n, generator=g)
    zq = (z * 4).round() / 4                      # many exact ties
    y = (torch.rand(n, generator=g) < 0.35).float()
    if n > 1:
        y[0], y[1] = 1.0, 0.0                      # both classes present
    for scores, sig in ((zq, False), (z, True), (zq, True)):
        out = ops.binary_auc(scores.cuda(), y.cuda(), apply_sigmoid=sig).cpu()
        p = (torch.sigmoid(scores) if sig else scores).numpy()
        assert out[2] == y.sum() and out[3] == n
        if n == 1:
            assert torch.isnan(out[0])
            continue
        tol = 1e-6 if sig else 1e-12
        assert abs(float(out[0]) - roc_auc_score(y.numpy(), p)) < tol
        assert abs(float(out[1]) - average_precision_score(y.numpy(), p)) < tol
    ones = ops.binary_auc(z.cuda(), torch.ones(n).cuda()).cpu()
    assert torch.isnan(ones[0])                    # roc_auc_score raises on a single class; the reference maps that to NaN


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("k", [1, 4, 8])
def test_sum_n(ops, dt, k):
    xs = [rnd(3, 40, 16, dtype=dt, seed=70 + i) for i in range(k)]
    want = torch.stack([x.float().cpu() for x in xs]).sum(0)
    got = ops.sum_n(xs)
    assert got.dtype == dt and rel(got, want) < TOL[dt]


def test_bin_events_bit_exact_vs_reference_golden(ops):
    """dx_bin_events and the build_stay_tensor / build_batch_tensors mirror against the reference's own build_stay_tensor
    (tests/golden/g5_binning.npz, oracle/make_golden_binning.py): integer/byte-level work -> bit-exact, NaN cells included."""
    import os
    import numpy as np
    import pandas as pd
    from multimodal_edema_prediction_b200.duett import mimic_dataset as md
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g5_binning.npz"))
    T, rs = int(G["T"]), G["row_start"]
    slot = np.where(G["slot_raw"] < 0, G["slot_raw"] + T, G["slot_raw"]).astype(np.int32)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x = ops.bin_events(cu(slot), cu(G["vals"]), cu(G["cnts"]), cu(rs), cu(G["means"]), cu(G["stds"]), T).cpu().numpy()
    assert np.array_equal(x, G["x"], equal_nan=True)
    # through the reference-facing builders, from pandas frames
    V = G["vals"].shape[1]
    all_vars = [f"v{j}" for j in range(V)]
    all_counts = [f"count_v{int(k)}" for k in G["shared_count_of"]]
    means = {v: float(m) for v, m in zip(all_vars, G["means"])}
    stds = {v: float(s) for v, s in zip(all_vars, G["stds"])}
    frames = []
    for b in range(len(rs) - 1):
        sl = slice(int(rs[b]), int(rs[b + 1]))
        df = pd.DataFrame({"slot_idx": G["slot_raw"][sl]})
        for j, v in enumerate(all_vars):
            df[v] = G["vals"][sl, j]
        for j, c in enumerate(all_counts):
            df[c] = G["cnts"][sl, j]
        frames.append(df)
    xb = md.build_batch_tensors(frames, means, stds, T, all_vars, all_counts).cpu().numpy()
    assert np.array_equal(xb, G["x"], equal_nan=True)
    x3 = md.build_stay_tensor(frames[3], means, stds, T, all_vars, all_counts)
    assert x3.is_cuda and np.array_equal(x3.cpu().numpy(), G["x"][3], equal_nan=True)
    bad = frames[1].copy()
    bad.loc[0, "slot_idx"] = -T - 1
    with pytest.raises(IndexError):
        md.build_stay_tensor(bad, means, stds, T, all_vars, all_counts)
    # The Dataset as the reference's training scripts use it: DataLoader workers + pin_memory + the collate function
    # (duett/train_duett_ssl.py:137).  __getitem__ is host-only (StayRows); the batch is binned by ONE launch in the main
    # process when the model's feats_to_input meets it.
    from multimodal_edema_prediction_b200.duett.duett import Model
    keep = [b for b, f in enumerate(frames) if len(f)]          # like the reference's Dataset, a stay needs at least one event row
    B = len(keep)
    icu = pd.concat([frames[b].assign(stay_id=100 + i) for i, b in enumerate(keep)], ignore_index=True)
    static = pd.DataFrame({"stay_id": [100 + b for b in range(B)], "age_at_intime": np.linspace(30, 80, B), "s0": 1.0, "s1": 0.0,
                           "label": [float(b % 2) for b in range(B)]})
    meta = {"N_TIMESTEPS": T, "LABEL_COL": "label", "age_mean": 55.0, "age_std": 10.0, "ONEHOT_STATIC": ["s0", "s1"],
            "means": means, "stds": stds, "ALL_VARS": all_vars, "ALL_COUNTS": all_counts, "D_STATIC": 3}
    ds = md.MIMICDataset([100 + b for b in range(B)], icu, static, meta)
    item = ds[0]
    assert isinstance(item[0][0], md.StayRows) and not torch.is_tensor(item[0][0])
    dl = torch.utils.data.DataLoader(ds, batch_size=B, shuffle=False, num_workers=2, pin_memory=True,
                                     collate_fn=md.collate_into_seqs)
    (xs_ts, xs_static, times), ys = next(iter(dl))
    model = Model(3, V, 1, d_embedding=8, masked_transform_timesteps=T, max_len=T, n_duett_layers=1, pretrain=False,
                  precision="fp32").cuda().eval()
    x_static_d, x_ts_d, x_times_d, n_ts = model.feats_to_input((xs_ts, xs_static, times), B)
    assert x_ts_d.is_cuda and np.array_equal(x_ts_d[:, :, :-1].cpu().numpy(), G["x"][keep], equal_nan=True)
    assert x_static_d.shape == (B, 3) and x_times_d.shape == (B, T) and n_ts == [T] * B and len(ys) == B


@pytest.mark.parametrize("dt", DT)
def test_dropout_matches_generator_spec(ops, dt):
    """dx_dropout against the numpy restatement of the generator (tests/ops_emulator.keep_factor): identical mask bits,
    keep rate ~ 1-p, backward = same call on the gradient."""
    x = rnd(37, 129, dtype=dt, seed=80)
    for p, seed in ((0.1, 12345), (0.5, 2 ** 63 + 7)):
        y = ops.dropout(x, p, seed)
        want = E.dropout(x.cpu(), p, seed)
        assert rel(y, want) < (1e-6 if dt == torch.float32 else 4e-3)
        assert torch.equal((y == 0).cpu(), (want == 0))
        keep = float((y != 0).float().mean())
        assert abs(keep - (1 - p)) < 0.03
    a, b, bias = rnd(50, 96, dtype=dt, seed=81), rnd(50, 96, dtype=dt, seed=82), rnd(96, seed=83)
    assert rel(ops.rowdot_bias(a, b, bias), E.rowdot_bias(*cpu(a, b, bias))) < TOL[dt]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("S,D,H", [(129, 128, 2), (33, 64, 2), (35, 24, 2), (513, 256, 2), (150, 128, 2)])
def test_attention_dropout_fwd_bwd(ops, dt, S, D, H):
    """Attention-probability dropout in the tensor-core (bf16, dh 32/64) and SIMT kernels against the emulator running on
    the same generator; the backward regenerates the mask."""
    B, drop = 3, (0.3, 987654321)
    qkv = rnd(B, S, 3 * D, dtype=dt, seed=84, scale=0.7)
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    o, lse = ops.attn_fwd(q, k, v, H, drop)
    ro, rlse = E.attn_fwd(*cpu(q, k, v), H, drop)
    assert rel(o, ro) < TOL[dt] * 2 and rel(lse, rlse) < 1e-3
    go = rnd(B, S, D, dtype=dt, seed=85)
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(q, k, v, o, go, lse, H, dqkv[:, :, :D], dqkv[:, :, D:2 * D], dqkv[:, :, 2 * D:], drop)
    rq, rk, rv = torch.empty(B, S, D), torch.empty(B, S, D), torch.empty(B, S, D)
    E.attn_bwd(*cpu(q, k, v, o, go, lse), H, rq, rk, rv, drop)
    assert rel(dqkv, torch.cat([rq, rk, rv], 2)) < TOL[dt] * 2


@pytest.mark.parametrize("M,N,K", [(520, 392, 1160), (136, 256, 2056), (1000, 520, 1096), (264, 264, 1032)])
def test_gemm_cta_pair_ragged_shapes(ops, monkeypatch, M, N, K):
    """CTA-pair (cluster of 2, tcgen05.mma.cta_group::2) kernel on shapes whose last 256-row tile is partly or (for the
    second CTA) wholly outside the matrix, N tails, and every operand layout that has a pair instance: staged epilogues with
    K-major and MN-major B, and the fp32-accumulating dW form — forced on (DX_GEMM_PAIR=1) and against the single-CTA
    kernel (DX_GEMM_PAIR=0) and the emulator."""
    bf = torch.bfloat16
    a, b = rnd(M, K, dtype=bf, seed=90), rnd(N, K, dtype=bf, seed=91, scale=0.1)
    bt = b.t().contiguous()                                   # [K,N]: MN-major B
    bias, rs = rnd(N, seed=92), torch.rand(M, device="cuda") + 0.5
    res = rnd(M, N, dtype=bf, seed=93)
    at_, bt_ = rnd(K, M, dtype=bf, seed=94), rnd(K, N, dtype=bf, seed=95, scale=0.1)    # dW form: A^T B over K rows

    def run():
        o1, o1b = torch.empty(M, N, device="cuda", dtype=bf), torch.empty(M, N, device="cuda", dtype=bf)
        ops.gemm_(a, b, out=o1, out2=o1b, row_scale=rs, bias=bias, act=ops.ACT_GELU, act_dtype=bf)
        o2, rq = torch.empty(M, N, device="cuda", dtype=bf), torch.zeros(M, device="cuda")
        ops.gemm_(a, bt, b_mn=True, out=o2, res=res, row_sumsq=rq, act_dtype=bf)
        o3 = torch.ones(M, N, device="cuda")
        ops.gemm_(at_, bt_, a_mn=True, b_mn=True, out=o3, accumulate=True)
        return o1, o1b, o2, rq, o3

    monkeypatch.setenv("DX_GEMM_PAIR", "1")
    pair = run()
    monkeypatch.setenv("DX_GEMM_PAIR", "0")
    single = run()
    for p_, s_ in zip(pair, single):
        assert rel(p_, s_) < 1e-5          # same bf16 products, fp32 accumulation: only the summation order differs
    r1, r1b = torch.empty(M, N), torch.empty(M, N)
    E.gemm_(*cpu(a, b), out=r1, out2=r1b, row_scale=rs.cpu(), bias=bias.cpu(), act=ops.ACT_GELU)
    assert rel(pair[0], r1) < TOL[bf] and rel(pair[1], r1b) < TOL[bf]
    r3 = torch.ones(M, N)
    E.gemm_(*cpu(at_, bt_), a_mn=True, b_mn=True, out=r3, accumulate=True)
    assert rel(pair[4], r3) < 1e-4


def test_gemm_b_multicast_cluster_mode(ops, monkeypatch):
    """Opt-in cluster mode of the staged kernels (two single-CTA MMAs on neighbouring M tiles, each CTA fetching half of the
    B block and TMA-multicasting it into both rings): same results as the default kernel, K-major and MN-major B."""
    bf = torch.bfloat16
    M, N, K = 1000, 520, 264
    a, b = rnd(M, K, dtype=bf, seed=96), rnd(N, K, dtype=bf, seed=97, scale=0.1)
    bt = b.t().contiguous()
    res, cx = rnd(M, N, dtype=bf, seed=98), rnd(M, N, dtype=bf, seed=99)
    num, den, bias = rnd(M, seed=100), torch.rand(M, device="cuda") + 1, rnd(N, seed=101)

    def run():
        o1, rq = torch.empty(M, N, device="cuda", dtype=bf), torch.zeros(M, device="cuda")
        ops.gemm_(a, b, out=o1, bias=bias, res=res, row_sumsq=rq, act_dtype=bf)
        o2 = torch.empty(M, N, device="cuda", dtype=bf)
        ops.gemm_(a, bt, b_mn=True, out=o2, res=res, cx=cx, coef_num=num, coef_den=den, act_dtype=bf)
        return o1, rq, o2

    monkeypatch.setenv("DX_GEMM_MCAST", "1")
    mc = run()
    monkeypatch.setenv("DX_GEMM_MCAST", "0")
    base = run()
    for x, y in zip(mc, base):
        assert rel(x, y) < 1e-5


def test_evaluate_dual_pathology_on_device(ops):
    """evaluate_dual_pathology with the real ranking kernel: masked per-label AUROC / AUPRC against sklearn."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from multimodal_edema_prediction_b200.training_duett import evaluator
    K = 3
    g = torch.Generator().manual_seed(5)

    class Teacher(torch.nn.Module):
        def forward(self, x_ts, x_static, bin_ends, pv):
            return {"img_logits": pv[:, :K], "ts_logits": pv[:, K:2 * K], "fusion_logits": pv[:, :K] + 0.3 * pv[:, 2 * K:],
                    "scaled_correction": 0.3 * pv[:, 2 * K:]}

    batches = []
    for _ in range(5):
        n = 40
        batches.append({"x_ts": tuple(torch.zeros(1) for _ in range(n)), "x_static": tuple(torch.zeros(1) for _ in range(n)),
                        "bin_ends": tuple(torch.zeros(1) for _ in range(n)), "y": torch.zeros(n),
                        "pixel_values": torch.randn(n, 3 * K, generator=g), "y_multi": (torch.rand(n, K, generator=g) < 0.4).float(),
                        "y_multi_mask": (torch.rand(n, K, generator=g) < 0.8).float()})
    res = evaluator.evaluate_dual_pathology(Teacher().cuda(), batches, torch.device("cuda"), ("a", "b", "c"))
    pv = torch.cat([b["pixel_values"] for b in batches]).numpy()
    y = torch.cat([b["y_multi"] for b in batches]).numpy()
    mk = torch.cat([b["y_multi_mask"] for b in batches]).numpy().astype(bool)
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    for k in range(K):
        m = mk[:, k]
        r = res["per_label"][k]
        assert r["n_valid"] == int(m.sum())
        assert abs(r["ts_auroc"] - roc_auc_score(y[m, k], sig(pv[m, K + k]))) < 1e-6
        assert abs(r["img_auprc"] - average_precision_score(y[m, k], sig(pv[m, k]))) < 1e-6
    assert res["n"] == 200 and res["main_auroc"] == res["main_auroc"]


# ---- fp32 operands on the tensor cores (tcgen05 kind::tf32): the reference's SSL / fine-tune precision ---------------------
TF32_TOL = 3e-3     # tf32 keeps 10 mantissa bits of each operand (relative 2^-11..2^-10 per product), fp32 accumulate


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_tf32_layouts(ops, a_mn, b_mn):
    M, N, K = 520, 392, 264
    a = rnd(*((K, M) if a_mn else (M, K)), seed=101)
    b = rnd(*((K, N) if b_mn else (N, K)), seed=102)
    out = torch.empty(M, N, device="cuda")
    ops.gemm_(a, b, a_mn=a_mn, b_mn=b_mn, out=out, tf32=True)
    ref = torch.empty(M, N)
    E.gemm_(*cpu(a, b), a_mn=a_mn, b_mn=b_mn, out=ref)
    e = rel(out, ref)
    assert 1e-6 < e < TF32_TOL, e               # > 1e-6: the tensor-core path really ran (FFMA would be ~1e-7)
    exact = torch.empty(M, N, device="cuda")
    ops.gemm_(a, b, a_mn=a_mn, b_mn=b_mn, out=exact)          # default fp32 mode stays exact
    assert rel(exact, ref) < 1e-5
    acc = torch.full((M, N), 2.0, device="cuda")
    ops.gemm_(a, b, a_mn=a_mn, b_mn=b_mn, out=acc, accumulate=True, tf32=True)
    assert rel(acc - 2.0, ref) < TF32_TOL


def test_gemm_tf32_fused_epilogues_and_grouped(ops):
    M, N, K = 300, 264, 136
    a, b = rnd(M, K, seed=103), rnd(N, K, seed=104, scale=0.1)
    rs = torch.rand(M, device="cuda") + 0.5
    bias, res = rnd(N, seed=105), rnd(M, N, seed=106)
    out, out2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    ops.gemm_(a, b, out=out, out2=out2, row_scale=rs, bias=bias, act=ops.ACT_GELU, act_dtype=torch.float32, tf32=True)
    r_out, r_out2 = torch.empty(M, N), torch.empty(M, N)
    E.gemm_(*cpu(a, b), out=r_out, out2=r_out2, row_scale=rs.cpu(), bias=bias.cpu(), act=ops.ACT_GELU)
    assert rel(out, r_out) < TF32_TOL and rel(out2, r_out2) < TF32_TOL
    rowsq = torch.zeros(M, device="cuda")
    ops.gemm_(a, b, out=out, bias=bias, res=res, row_sumsq=rowsq, act_dtype=torch.float32, tf32=True)
    r_rowsq = torch.zeros(M)
    E.gemm_(*cpu(a, b), out=r_out, bias=bias.cpu(), res=res.cpu(), row_sumsq=r_rowsq)
    assert rel(out, r_out) < TF32_TOL and rel(rowsq, r_rowsq) < TF32_TOL
    # grouped mode on the strided views of the embedding (64 -> d and its dW)
    G, R, d, H = 5, 300, 24, 64
    psi = rnd(R, G, d, seed=107)
    av = psi.permute(1, 0, 2)
    hn, w = rnd(G, R, H, seed=108), rnd(G, d, H, seed=109, scale=0.3)
    outp = torch.zeros(R, G, d, device="cuda")
    ops.gemm_(hn, w, out=outp.permute(1, 0, 2), bias=rnd(G, d, seed=110), act_dtype=torch.float32, tf32=True)
    want = torch.einsum("grh,gdh->grd", hn, w) + rnd(G, d, seed=110)[:, None, :]
    assert rel(outp.permute(1, 0, 2), want) < TF32_TOL
    dw = torch.ones(G, d, H, device="cuda")
    ops.gemm_(av, hn, a_mn=True, b_mn=True, out=dw, accumulate=True, tf32=True)
    assert rel(dw - 1.0, torch.einsum("grd,grh->gdh", av, hn)) < TF32_TOL


def test_gemm_tf32_production_shape_multi_tile(ops):
    """33 024 x 512 x 4 224 (FFN-in, event axis) and its dW in fp32 storage: several tiles per persistent CTA, checked per
    128 x 256 tile against torch's fp32 matmul."""
    torch.backends.cuda.matmul.allow_tf32 = False
    M, N, K = 33024, 512, 4224
    a, b = rnd(M, K, seed=111), rnd(N, K, seed=112, scale=0.02)
    out = torch.empty(M, N, device="cuda")
    ops.gemm_(a, b, out=out, bias=rnd(N, seed=113), act_dtype=torch.float32, tf32=True)
    ref = a @ b.t() + rnd(N, seed=113)[None]
    assert rel(out, ref) < TF32_TOL
    d = ((out - ref) ** 2).reshape(M // 128, 128, N // 256, 256).sum((1, 3))
    r = (ref ** 2).reshape(M // 128, 128, N // 256, 256).sum((1, 3))
    assert float((d / r).sqrt().max()) < 2 * TF32_TOL
    dw = torch.zeros(N, K, device="cuda")
    ops.gemm_(out, a, a_mn=True, b_mn=True, out=dw, accumulate=True, tf32=True)
    assert rel(dw, out.t() @ a) < TF32_TOL
