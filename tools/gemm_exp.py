"""Timing experiments on the production GEMM shapes: runs each (shape, epilogue) under a list of environment settings of
the launcher's experiment knobs (DX_GEMM_* are read at every launch) and prints one table.
usage: python tools/gemm_exp.py [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from multimodal_edema_prediction_b200 import ops  # noqa: E402

bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
flush = torch.empty(512 << 20, device="cuda", dtype=torch.uint8)


def make(M, N, K, epi):
    kw = {}
    if epi == "dw":
        a, b = rn(K, M).to(bf), rn(K, N).to(bf)
        kw = dict(a_mn=True, b_mn=True, out=torch.zeros(M, N, device="cuda"), accumulate=True)
        nbytes = (K * M + K * N) * 2 + M * N * 8
        return a, b, kw, nbytes
    a = rn(M, K).to(bf)
    bm = epi in ("dx", "gelu_bwd", "plain_b")
    b = rn(K, N).to(bf) if bm else rn(N, K).to(bf)
    nbytes = (M * K + N * K + M * N) * 2
    kw = dict(out=torch.empty(M, N, device="cuda", dtype=bf), act_dtype=bf, b_mn=bm)
    if epi == "dx":
        kw.update(res=rn(M, N).to(bf), cx=rn(M, N).to(bf), coef_num=rn(M), coef_den=torch.rand(M, device="cuda") + 1)
        nbytes += 2 * M * N * 2
    elif epi == "resid":
        kw.update(res=rn(M, N).to(bf), bias=rn(N), row_sumsq=torch.zeros(M, device="cuda"))
        nbytes += M * N * 2
    elif epi == "res_rowsq":
        kw.update(res=rn(M, N).to(bf), row_sumsq=torch.zeros(M, device="cuda"))
        nbytes += M * N * 2
    elif epi == "rs":
        kw.update(row_scale=torch.rand(M, device="cuda"))
    elif epi == "ffn_in":
        kw.update(row_scale=torch.rand(M, device="cuda"), bias=rn(N), act=ops.ACT_GELU, out2=torch.empty(M, N, device="cuda", dtype=bf))
        nbytes += M * N * 2
    elif epi == "gelu_bwd":
        kw.update(act=ops.ACT_GELU_BWD, aux=rn(M, N).to(bf), aux_bias=rn(N), row_scale2=torch.rand(M, device="cuda"),
                  out2=torch.empty(M, N, device="cuda", dtype=bf), row_dot=torch.zeros(M, device="cuda"))
        nbytes += 2 * M * N * 2
    return a, b, kw, nbytes


def time_it(a, b, kw, iters=12):
    for _ in range(2):
        ops.gemm_(a, b, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_(a, b, **kw); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


SHAPES = [
    (33024, 4224, 512, "dx"), (8448, 16512, 512, "dx"), (33024, 4224, 384, "dx"), (8448, 16512, 384, "dx"),
    (33024, 4224, 512, "resid"), (8448, 16512, 512, "resid"), (33024, 4224, 128, "res_rowsq"), (8448, 16512, 128, "res_rowsq"),
    (33024, 512, 4224, "ffn_in"), (8448, 512, 16512, "ffn_in"), (33024, 512, 4224, "gelu_bwd"), (8448, 512, 16512, "gelu_bwd"),
    (33024, 384, 4224, "rs"), (8448, 384, 16512, "rs"), (8448, 128, 16512, "plain_b"), (33024, 128, 4224, "plain_b"),
    (16512, 512, 8448, "dw"), (512, 16512, 8448, "dw"), (4224, 512, 33024, "dw"), (512, 4224, 33024, "dw"),
    (384, 16512, 8448, "dw"), (384, 4224, 33024, "dw"), (16512, 128, 8448, "dw"), (4224, 128, 33024, "dw"),
]
ENVS = [{}] + [json.loads(x) for x in os.environ.get("GEMM_EXP_ENVS", "").split(";") if x.strip()]
only = os.environ.get("GEMM_EXP_ONLY")          # comma list of epilogue names
rows = []
for (M, N, K, epi) in SHAPES:
    if only and epi not in only.split(","):
        continue
    a, b, kw, nbytes = make(M, N, K, epi)
    fl = 2.0 * M * N * K
    ideal = max(fl / 1408.4e12, nbytes / 6462.1e9) * 1e3
    r = {"shape": f"{M}x{N}x{K}", "epi": epi, "ideal_ms": ideal, "runs": []}
    for env in ENVS:
        for k, v in env.items():
            os.environ[k] = str(v)
        try:
            ms = time_it(a, b, kw)
        except Exception as ex:  # noqa: BLE001
            ms = float("nan")
            print("ERR", env, repr(ex)[:200])
        for k in env:
            del os.environ[k]
        r["runs"].append({"env": env, "ms": ms, "tflops": fl / ms / 1e9, "gbs": nbytes / ms / 1e6, "frac": ideal / ms})
    rows.append(r)
    print(f"{r['shape']:>18} {epi:>9} ideal {ideal*1e3:6.1f}us | " +
          " | ".join(f"{json.dumps(x['env']) if x['env'] else 'default'}: {x['ms']*1e3:6.1f}us {x['frac']:.2f}" for x in r["runs"]), flush=True)
    del a, b, kw
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
