"""GPU bring-up for dx_gemm: each group runs in its own subprocess so a trapped kernel cannot poison the rest.
Usage (on the GPU box): python tools/gemm_bringup.py [group ...]   -> gpurun_out/gemm_bringup.json
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(a, b):
    import torch
    a = a.float(); b = b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_case(name, M, N, K, a_mn, b_mn, dtype, mode, bn=0, stages=0, lbos=(-1, -1, -1, -1), epi="plain"):
    import ctypes as C
    import torch
    from multimodal_edema_prediction_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g).to(dt)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g).to(dt)
    Af = (A.float().t() if a_mn else A.float())
    Bf = (B.float().t() if b_mn else B.float())
    ref = Af.double() @ Bf.double().t()
    kw = {}
    out_dt = torch.float32
    extra = {}
    if epi == "plain":
        pass
    elif epi == "ffn_in":   # rowscale + bias + gelu, out2 = pre
        rs = torch.rand(M, device="cuda", generator=g) + 0.5
        bias = torch.randn(N, device="cuda", generator=g)
        pre = ref * rs.double()[:, None] + bias.double()[None]
        ref = torch.nn.functional.gelu(pre)
        out2 = torch.empty(M, N, device="cuda", dtype=dt)
        kw = dict(row_scale=rs, bias=bias, act=L.ACT_GELU, out2=out2)
        extra["out2"] = (out2, pre)
        out_dt = dt
    elif epi == "resid":    # residual + bias + rowsq
        res = torch.randn(M, N, device="cuda", generator=g).to(dt)
        bias = torch.randn(N, device="cuda", generator=g)
        ref = ref + bias.double()[None] + res.double()
        rowsq = torch.zeros(M, device="cuda")
        kw = dict(res=res, bias=bias, row_sumsq=rowsq)
        extra["rowsq"] = (rowsq, (ref ** 2).sum(1))
        out_dt = dt
    elif epi == "dx":       # res + acc - cx*num/den
        res = torch.randn(M, N, device="cuda", generator=g).to(dt)
        cx = torch.randn(M, N, device="cuda", generator=g).to(dt)
        num = torch.randn(M, device="cuda", generator=g)
        den = torch.rand(M, device="cuda", generator=g) + 1.0
        ref = ref + res.double() - cx.double() * (num / den).double()[:, None]
        kw = dict(res=res, cx=cx, coef_num=num, coef_den=den)
        out_dt = dt
    elif epi == "gelu_bwd":
        aux = torch.randn(M, N, device="cuda", generator=g).to(dt)
        ab = torch.randn(N, device="cuda", generator=g)
        rs2 = torch.rand(M, device="cuda", generator=g) + 0.5
        x = aux.double()
        gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * 3.141592653589793) ** 0.5
        dv = ref * gp
        rd_ref = (dv * (x - ab.double()[None])).sum(1)
        ref = dv * rs2.double()[:, None]
        out2 = torch.empty(M, N, device="cuda", dtype=dt)
        rd = torch.zeros(M, device="cuda")
        kw = dict(aux=aux, aux_bias=ab, row_scale2=rs2, out2=out2, row_dot=rd, act=L.ACT_GELU_BWD)
        extra["out2"] = (out2, dv)
        extra["rowdot"] = (rd, rd_ref)
        out_dt = dt
    elif epi == "accum":
        kw = dict(accumulate=True)
    out = torch.full((M, N), 0.5 if epi == "accum" else float("nan"), device="cuda", dtype=out_dt)
    if epi == "accum":
        ref = ref + 0.5
    d = L.make_gemm_desc(A, B, a_mn=a_mn, b_mn=b_mn, out=out, force_simt=(mode == "simt"), **kw)
    if mode == "tc_debug":
        rc = L.lib().dx_gemm_tc_debug(C.byref(d), bn, stages, *lbos, L.stream_ptr())
    else:
        rc = L.lib().dx_gemm(C.byref(d), L.stream_ptr())
    L.check(rc)
    torch.cuda.synchronize()
    res = {"name": name, "shape": [M, N, K], "a_mn": a_mn, "b_mn": b_mn, "dtype": dtype, "mode": mode, "bn": bn,
           "stages": stages, "lbos": list(lbos), "epi": epi, "err": rel_err(out, ref)}
    for k, (got, want) in extra.items():
        res["err_" + k] = rel_err(got, want)
    return res


def group_cases(group):
    cs = []
    if group == "simt":
        for a_mn in (False, True):
            for b_mn in (False, True):
                cs.append(dict(name="simt_f32", M=130, N=200, K=77, a_mn=a_mn, b_mn=b_mn, dtype="f32", mode="simt"))
        for epi in ("ffn_in", "resid", "dx", "gelu_bwd", "accum"):
            cs.append(dict(name="simt_epi", M=130, N=200, K=77, a_mn=False, b_mn=False, dtype="f32", mode="simt", epi=epi))
        cs.append(dict(name="simt_bf16", M=130, N=200, K=80, a_mn=False, b_mn=True, dtype="bf16", mode="simt"))
    elif group == "tc_kk":
        for (M, N, K, bn, st) in ((128, 128, 64, 128, 3), (128, 128, 256, 128, 3), (256, 384, 512, 128, 3),
                                  (300, 200, 600, 128, 3), (256, 512, 1024, 256, 4), (1000, 72, 840, 64, 4),
                                  (512, 256, 2048, 128, 6)):
            cs.append(dict(name="tc_kk", M=M, N=N, K=K, a_mn=False, b_mn=False, dtype="bf16", mode="tc_debug", bn=bn, stages=st))
    elif group == "tc_kmn":
        for (M, N, K, bn, st) in ((128, 128, 64, 128, 3), (256, 384, 512, 128, 3), (300, 200, 600, 256, 4)):
            cs.append(dict(name="tc_kmn", M=M, N=N, K=K, a_mn=False, b_mn=True, dtype="bf16", mode="tc_debug", bn=bn, stages=st))
    elif group == "tc_mnk":
        for (M, N, K, bn, st) in ((128, 128, 64, 128, 3), (256, 384, 512, 128, 3), (300, 200, 600, 256, 4)):
            cs.append(dict(name="tc_mnk", M=M, N=N, K=K, a_mn=True, b_mn=False, dtype="bf16", mode="tc_debug", bn=bn, stages=st))
    elif group == "tc_mnmn":
        for (M, N, K, bn, st) in ((128, 128, 64, 128, 3), (256, 384, 512, 128, 3), (520, 512, 1000, 256, 4)):
            cs.append(dict(name="tc_mnmn", M=M, N=N, K=K, a_mn=True, b_mn=True, dtype="bf16", mode="tc_debug", bn=bn, stages=st))
    elif group == "tc_epi":
        for epi in ("ffn_in", "resid", "dx", "gelu_bwd", "accum"):
            cs.append(dict(name="tc_epi", M=300, N=520, K=256, a_mn=False, b_mn=False, dtype="bf16", mode="tc", epi=epi))
    elif group.startswith("sweep_"):
        # descriptor sweep for MN-major operands, only needed if the built-in values fail
        which = group.split("_")[1]
        cands = [(8192, 1024), (1024, 8192), (16, 1024), (1024, 16), (8192, 128), (128, 8192), (1024, 1024), (2048, 1024)]
        for (lbo, sbo) in cands:
            if which == "a":
                cs.append(dict(name=group, M=128, N=128, K=64, a_mn=True, b_mn=False, dtype="bf16", mode="tc_debug", bn=128, stages=3, lbos=(lbo, sbo, -1, -1)))
            else:
                cs.append(dict(name=group, M=128, N=128, K=64, a_mn=False, b_mn=True, dtype="bf16", mode="tc_debug", bn=128, stages=3, lbos=(-1, -1, lbo, sbo)))
    return cs


def child(group):
    out = []
    for c in group_cases(group):
        try:
            r = run_case(**c)
        except Exception as ex:  # a CUDA fault is sticky: stop this group
            out.append({**c, "error": repr(ex)[:300]})
            print(json.dumps(out[-1]), flush=True)
            break
        out.append(r)
        print(json.dumps(r), flush=True)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    groups = sys.argv[1:] or ["simt", "tc_kk", "tc_kmn", "tc_mnk", "tc_mnmn", "tc_epi"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = {}
    for g in groups:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--child", g], capture_output=True, text=True, timeout=240)
            lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
            results[g] = {"rc": p.returncode, "cases": lines, "stderr_tail": p.stderr[-1500:], "secs": time.time() - t0}
        except subprocess.TimeoutExpired as ex:
            results[g] = {"rc": "timeout", "stdout": (ex.stdout or b"")[-2000:].decode(errors="replace") if isinstance(ex.stdout, bytes) else str(ex.stdout)[-2000:]}
        print(g, json.dumps(results[g])[:3000], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "gemm_bringup.json"), "w") as f:
            json.dump(results, f, indent=1)
    bad = [g for g, r in results.items() if r.get("rc") != 0 or any(("error" in c) or c.get("err", 1) > 2e-2 for c in r.get("cases", []))]
    print("BAD GROUPS:", bad)
