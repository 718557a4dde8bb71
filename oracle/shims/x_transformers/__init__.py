"""ORACLE / TEST INFRASTRUCTURE ONLY — stand-in for the third-party `x_transformers` package.

`x-transformers` (PyPI, lucidrains) is imported by the reference at duett/duett.py:7 with NO version pin anywhere in
the repository and is not installed here (no network).  This file restates the published algorithm of the single
call signature the reference uses (duett/duett.py:95-105):

    Encoder(dim, depth=1, heads=h, pre_norm=True, use_scalenorm=True, attn_dim_head=d//h, ff_glu=False,
            ff_mult=d_ff/dim, attn_dropout=p, ff_dropout=p);   forward(x[B,N,dim]) — no mask, no context
            (call sites duett/duett.py:276,279 and models/main_architecture_duett.py:81,91)

following x-transformers 1.x/2.x semantics: pre-norm ScaleNorm branches, bias-free
q/k/v/out projections, fp32 softmax, erf-GELU feed-forward with inner = int(dim * ff_mult), and a final ScaleNorm.
PARITY UNPINNED: the reference holds no test or golden vector for this boundary.  Module/parameter names follow
the library's state-dict layout (layers.{0,1}.0.0.g, layers.0.1.to_{q,k,v,out}.weight, layers.1.1.ff.0.0.*,
layers.1.1.ff.2.*, final_norm.g) so real checkpoints would load.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ScaleNorm(nn.Module):
    def __init__(self, dim, eps_mode="normalize"):
        super().__init__()
        self.scale = dim ** 0.5
        self.g = nn.Parameter(torch.ones(1))
        self.eps_mode = eps_mode

    def forward(self, x):
        if self.eps_mode == "normalize":          # recent vintage: F.normalize (eps 1e-12)
            return F.normalize(x, dim=-1) * self.scale * self.g
        norm = torch.norm(x, dim=-1, keepdim=True) * (self.scale ** -1)   # pre-2023 vintage: clamp 1e-5
        return x / norm.clamp(min=1e-5) * self.g


class Attention(nn.Module):
    def __init__(self, dim, heads, dim_head, dropout=0.0):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.attn_dropout = nn.Dropout(dropout)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x):
        b, n, _ = x.shape
        h = self.heads
        q, k, v = (t.view(b, n, h, -1).transpose(1, 2) for t in (self.to_q(x), self.to_k(x), self.to_v(x)))
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        attn = sim.softmax(dim=-1, dtype=torch.float32).to(sim.dtype)
        attn = self.attn_dropout(attn)
        out = torch.einsum("bhij,bhjd->bhid", attn, v)
        return self.to_out(out.transpose(1, 2).reshape(b, n, -1))


class FeedForward(nn.Module):
    def __init__(self, dim, mult, dropout=0.0):
        super().__init__()
        inner = int(dim * mult)   # float expression kept verbatim (SURVEY §8 "Note on F")
        self.ff = nn.Sequential(nn.Sequential(nn.Linear(dim, inner), nn.GELU()), nn.Dropout(dropout),
                                nn.Linear(inner, dim))

    def forward(self, x):
        return self.ff(x)


class Encoder(nn.Module):
    def __init__(self, dim, depth=1, heads=8, pre_norm=True, use_scalenorm=False, attn_dim_head=64, ff_glu=False,
                 ff_mult=4, attn_dropout=0.0, ff_dropout=0.0, final_norm=True, scalenorm_eps_mode="normalize", **kw):
        super().__init__()
        if not (pre_norm and use_scalenorm) or ff_glu or kw:
            raise NotImplementedError("shim covers only the reference's call signature (duett/duett.py:95-105)")
        mk = lambda: ScaleNorm(dim, scalenorm_eps_mode)
        self.layers = nn.ModuleList()
        for _ in range(depth):
            self.layers.append(nn.ModuleList([nn.ModuleList([mk()]), Attention(dim, heads, attn_dim_head, attn_dropout)]))
            self.layers.append(nn.ModuleList([nn.ModuleList([mk()]), FeedForward(dim, ff_mult, ff_dropout)]))
        self.final_norm = mk() if final_norm else nn.Identity()

    def forward(self, x):
        for norms, block in self.layers:
            x = x + block(norms[0](x))
        return self.final_norm(x)
