"""Single-kernel parity checks through the C ABI test entry points, shared by
tests/test_kernels_gpu.py (pytest, -m gpu) and tools/gpu_check.py (one process per check).
Each check compares a CUDA kernel with a plain PyTorch fp32 computation of the same op on the
same (fp16-rounded) inputs and returns a dict of error metrics; `assert_ok` applies the
tolerances written next to each check."""
from __future__ import annotations

import math

import torch

from zipvoice_b200 import _lib
from zipvoice_b200.weights import pack_pos_table, pack_pos_table_tc

DEV = "cuda"
H16 = torch.float16          # storage type of every activation / weight on the CUDA path


def _s():
    return torch.cuda.current_stream().cuda_stream


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _swoosh_l(x):
    return torch.logaddexp(torch.zeros((), device=x.device), x - 4.0) - 0.08 * x - 0.035


def _swoosh_r(x):
    return torch.logaddexp(torch.zeros((), device=x.device), x - 1.0) - 0.08 * x - 0.313261687


def check_linear(M=300, K=512, N=272, block_n=0, act=0, resid=False, bypass=False, out_mode=0, seed=0):
    """tolerance: rel-L2 <= 8e-4 for fp16 outputs (rounding 2^-12), <= 2e-5 for fp32 outputs.
    out_mode 0: fp16, 1: fp32; resid: fp16 residual-stream tile added in the epilogue (TMA aux ring);
    bypass: out = orig + (out - orig) * scale with an fp16 `orig` tile riding the same ring."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    kp = (K + 7) // 8 * 8
    A = torch.zeros(M, kp, dtype=H16)
    A[:, :K] = (torch.randn(M, K, generator=g) * 1.0).to(H16)
    W = torch.zeros(N, kp, dtype=H16)
    W[:, :K] = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(H16)
    b = torch.randn(N, generator=g)
    R = torch.randn(M, N, generator=g).to(H16).to(DEV) if resid else None
    O = torch.randn(M, N, generator=g).to(H16).to(DEV) if bypass else None
    sc = (torch.rand(N, generator=g) * 0.6 + 0.3).to(DEV) if bypass else None
    A, W, b = A.to(DEV), W.to(DEV), b.to(DEV)
    out = torch.full((M, N), float("nan"), dtype=H16 if out_mode == 0 else torch.float32, device=DEV)
    _lib.check(lib.zvb_test_linear(A.data_ptr(), M, K, kp, W.data_ptr(), b.data_ptr(), N, kp, block_n, act,
                                   R.data_ptr() if resid else None, O.data_ptr() if bypass else None,
                                   sc.data_ptr() if bypass else None, out.data_ptr(), N, out_mode, _s()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + b
    if act == 1:
        ref = _swoosh_l(ref)
    elif act == 2:
        ref = _swoosh_r(ref)
    if resid:
        ref = ref + R.float()
    if bypass:
        ref = O.float() + (ref - O.float()) * sc
    err = (out.float() - ref).abs().nan_to_num(1e9)
    am = int(err.argmax())
    return dict(rel=_rel(out.float().nan_to_num(0.0), ref), nan=int(torch.isnan(out.float()).sum()),
                tol=8e-4 if out_mode == 0 else 2e-5,
                worst=[am // N, am % N, float(out.float().flatten()[am]), float(ref.flatten()[am])],
                bad_frac=float((err > 0.1).float().mean()),
                bad_rows=int((err.max(dim=1).values > 0.1).sum()), bad_cols=int((err.max(dim=0).values > 0.1).sum()))


def check_gated(M=300, K=512, n_out=384, mode=1, masked=False, seed=0):
    """Gated projection on tile-packed weights: mode 1 x*tanh(s) (rows [s|x]), mode 2 GLU x*sigmoid(s)
    (rows [x|s]) with optional zeroed rows.  tolerance: rel-L2 <= 8e-3 (tanh.approx / fast sigmoid + fp16)."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(H16)
    Wa = (torch.randn(n_out, K, generator=g) / math.sqrt(K)).to(H16)
    Wb = (torch.randn(n_out, K, generator=g) / math.sqrt(K)).to(H16)
    ba, bb = torch.randn(n_out, generator=g) * 0.3, torch.randn(n_out, generator=g) * 0.3
    tiles = (n_out + 127) // 128
    W = torch.zeros(tiles * 256, K, dtype=H16)
    Bv = torch.zeros(tiles * 256)
    for t in range(tiles):
        r = min(128, n_out - t * 128)
        W[t * 256: t * 256 + r] = Wa[t * 128: t * 128 + r]
        W[t * 256 + 128: t * 256 + 128 + r] = Wb[t * 128: t * 128 + r]
        Bv[t * 256: t * 256 + r] = ba[t * 128: t * 128 + r]
        Bv[t * 256 + 128: t * 256 + 128 + r] = bb[t * 128: t * 128 + r]
    rm = (torch.rand(M, generator=g) < 0.3) if masked else None
    A, W, Bv = A.to(DEV), W.to(DEV), Bv.to(DEV)
    rm8 = rm.to(torch.uint8).to(DEV) if masked else None
    out = torch.full((M, n_out), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_gated(A.data_ptr(), M, K, K, W.data_ptr(), Bv.data_ptr(), tiles * 256, n_out, K, mode,
                                  rm8.data_ptr() if masked else None, out.data_ptr(), n_out, _s()))
    torch.cuda.synchronize()
    a = A.float() @ Wa.float().to(DEV).t() + ba.to(DEV)
    b = A.float() @ Wb.float().to(DEV).t() + bb.to(DEV)
    ref = b * torch.tanh(a) if mode == 1 else a * torch.sigmoid(b)
    if masked:
        ref = ref * (~rm).to(DEV).unsqueeze(-1)
    return dict(rel=_rel(out.float().nan_to_num(0.0), ref), nan=int(torch.isnan(out.float()).sum()), tol=8e-3)


def _attn_inputs(N, H, L, seed, masked, pos_scale=0.5):
    g = torch.Generator(device="cpu").manual_seed(seed)
    ld = H * 68
    qkp = torch.randn(N, L, ld, generator=g)
    qkp[..., : 2 * H * 32] *= 0.45          # q.k std ~ 1.1 ... a few units of score range
    qkp = qkp.to(H16)
    E = torch.randn(H, 2 * L - 1, 4, generator=g) * pos_scale
    mask = torch.zeros(N, L, dtype=torch.bool)
    if masked:
        lens = torch.randint(max(1, L // 2), L + 1, (N,), generator=g)
        lens[0] = L
        mask = torch.arange(L)[None, :] >= lens[:, None]
    return qkp.to(DEV), E.to(DEV), mask.to(DEV)


def _attn_ref(qkp, E, mask, H):
    N, L, _ = qkp.shape
    x = qkp.float()
    q = x[..., : H * 32].reshape(N, L, H, 32).permute(0, 2, 1, 3)
    k = x[..., H * 32: 2 * H * 32].reshape(N, L, H, 32).permute(0, 2, 1, 3)
    p = x[..., 2 * H * 32:].reshape(N, L, H, 4).permute(0, 2, 1, 3)
    s = q @ k.transpose(-1, -2)
    idx = (L - 1) - torch.arange(L, device=x.device)[:, None] + torch.arange(L, device=x.device)[None, :]
    pos = torch.einsum("nhic,hrc->nhir", p, E)                 # (N,H,L,2L-1)
    s = s + torch.gather(pos, 3, idx.expand(N, H, L, L))
    s = s.masked_fill(mask[:, None, None, :], -1000.0)
    return s.softmax(-1)


def check_attn(N=2, H=4, L=200, masked=True, seed=0, pos_scale=0.5, tol=2.5e-3, tc=True):
    """tolerance: max-abs <= 2.5e-3 on probabilities (the rel-pos bias and the exponent are evaluated in
    packed fp16: ~2^-9 absolute on an exponent of a few units; fp16 storage of P)"""
    lib = _lib.load()
    qkp, E, mask = _attn_inputs(N, H, L, seed, masked, pos_scale)
    Lk = (L + 7) // 8 * 8
    P = torch.full((N, H, L, Lk), float("nan"), dtype=H16, device=DEV)
    inv_l = torch.full((N, H, L), float("nan"), dtype=torch.float32, device=DEV)
    m8 = mask.to(torch.uint8).contiguous()
    Ex = pack_pos_table_tc(E) if tc else pack_pos_table(E)
    scratch = torch.zeros(N * 4 * ((L + 127) // 128), dtype=torch.int32, device=DEV)
    fn = lib.zvb_test_attn_weights_tc if tc else lib.zvb_test_attn_weights
    _lib.check(fn(qkp.data_ptr(), H * 68, Ex.data_ptr(), m8.data_ptr(), scratch.data_ptr(),
                  P.data_ptr(), inv_l.data_ptr(), N, H, L, Lk, _s()))
    torch.cuda.synchronize()
    ref = _attn_ref(qkp, E, mask, H)
    pmax = float(P.float()[..., :L].max())
    assert 0 < pmax <= 4096.0 * 1.05, pmax                 # unnormalised weights live in (0, 2^12]
    got = P.float() * inv_l.unsqueeze(-1)
    pad = got[..., L:]
    return dict(maxabs=float((got[..., :L] - ref).abs().max()), rel=_rel(got[..., :L], ref),
                nan=int(torch.isnan(got).sum()), pad_nonzero=int((pad != 0).sum()),
                rowsum_err=float((got[..., :L].sum(-1) - 1).abs().max()), tol=tol)


def check_pv(N=2, H=4, L=200, hd=12, hp=16, per_head=True, mul=False, seed=0):
    """tolerance: rel-L2 <= 8e-4 (fp16 output)"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    Lk = (L + 7) // 8 * 8
    P = torch.zeros(N, H, L, Lk)
    P[..., :L] = torch.rand(N, H, L, L, generator=g).pow(4)
    P = (P / P.sum(-1, keepdim=True)).to(H16).to(DEV)
    if per_head:
        V = torch.randn(N, H, hd, L, generator=g).to(H16)
        Vt = torch.zeros(N, H, hp, Lk, dtype=H16)
        Vt[:, :, :hd, :L] = V
        out = torch.full((N, L, H * hd), float("nan"), dtype=H16, device=DEV)
        ref = torch.einsum("nhij,nhdj->nihd", P.float()[..., :L], V.float().to(DEV)).reshape(N, L, H * hd)
    else:
        V = torch.randn(N, hd, L, generator=g).to(H16)
        Vt = torch.zeros(N, hd, Lk, dtype=H16)
        Vt[:, :, :L] = V
        out = torch.full((N, L, hd), float("nan"), dtype=H16, device=DEV)
        ref = torch.einsum("nij,ndj->nid", P.float()[:, 0, :, :L], V.float().to(DEV))
    inv_l = (torch.rand(N, H, L, generator=g) + 0.5).to(DEV)
    ref = ref * (inv_l.permute(0, 2, 1).repeat_interleave(hd, dim=2) if per_head else inv_l[:, 0].unsqueeze(-1))
    Y = None
    if mul:
        Y = torch.randn(N, L, hd, generator=g).to(H16).to(DEV)
        ref = ref * Y.float()
    Vt = Vt.to(DEV)
    _lib.check(lib.zvb_test_pv(P.data_ptr(), inv_l.data_ptr(), Vt.data_ptr(), out.data_ptr(), N, H, L, Lk, hd, hp,
                               1 if per_head else 0, Y.data_ptr() if mul else None, _s()))
    torch.cuda.synchronize()
    return dict(rel=_rel(out.float(), ref), nan=int(torch.isnan(out.float()).sum()), tol=8e-4)


def check_biasnorm(rows=1000, C=512, seed=0):
    """fp16 stream in/out + fp16 time-embedded copy, fp32 arithmetic.
    tolerance: rel-L2 <= 6e-4 (fp16 outputs)"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = 37
    src = torch.randn(rows, C, generator=g).mul(2).to(H16).to(DEV)
    orig = torch.randn(rows, C, generator=g).to(H16).to(DEV)
    nb = (torch.randn(C, generator=g) * 0.1).to(DEV)
    ls = torch.tensor([0.4], device=DEV)
    bs = (torch.rand(C, generator=g) * 0.6 + 0.3).to(DEV)
    temb = torch.randn((rows + L - 1) // L, C, generator=g).to(DEV)
    out = torch.full((rows, C), float("nan"), dtype=H16, device=DEV)
    ot = torch.full((rows, C), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_biasnorm_bypass(src.data_ptr(), orig.data_ptr(), out.data_ptr(), ot.data_ptr(),
                                            temb.data_ptr(), L, nb.data_ptr(), ls.data_ptr(),
                                            bs.data_ptr(), rows, C, _s()))
    torch.cuda.synchronize()
    x, o = src.float(), orig.float()
    y = x * (((x - nb) ** 2).mean(-1, keepdim=True) ** -0.5) * ls.exp()
    ref = o + (y - o) * bs
    ref_t = ref + temb[torch.arange(rows, device=DEV) // L]
    rel = max(_rel(out.float(), ref), _rel(ot.float(), ref_t))
    return dict(rel=rel, nan=int(torch.isnan(out.float()).sum() + torch.isnan(ot.float()).sum()), tol=6e-4)


def check_dwconv(N=2, L=150, C=512, K=31, seed=0):
    """tolerance: rel-L2 <= 8e-4 (fp16 output)"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, L, C, generator=g).to(H16).to(DEV)
    w = (torch.randn(C, 1, K, generator=g) / math.sqrt(K)).to(DEV)
    b = (torch.randn(C, generator=g) * 0.2).to(DEV)
    wt = w.reshape(C, K).t().contiguous()
    out = torch.full((N, L, C), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_dwconv(x.data_ptr(), out.data_ptr(), wt.data_ptr(), b.data_ptr(), N, L, C, K, _s()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv1d(x.float().permute(0, 2, 1), w, b, padding=K // 2, groups=C).permute(0, 2, 1)
    ref = _swoosh_r(ref)
    return dict(rel=_rel(out.float(), ref), nan=int(torch.isnan(out.float()).sum()), tol=8e-4)


def check_cfg_euler(B=3, T=50, F=100, cfg=1, seed=0):
    """tolerance: exact to fp32 rounding (rel-L2 <= 1e-6)"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, T, F, generator=g).to(DEV)
    v = torch.randn(2 * B if cfg else B, T, F, generator=g).to(DEV)
    gd = torch.tensor([0.0, 0.7, 2.0])[:B].to(DEV)
    ts = torch.tensor([0.1, 0.35, 0.8], device=DEV)
    x0 = x.clone()
    _lib.check(lib.zvb_test_cfg_euler(x.data_ptr(), v.data_ptr(), gd.data_ptr(), 2.0, ts.data_ptr(), 0, B, T * F, cfg,
                                      _s()))
    torch.cuda.synchronize()
    if cfg:
        gg = (2.0 * gd)[:, None, None]
        vel = (1 + gg) * v[B:] - gg * v[:B]
    else:
        vel = v
    ref = x0 + vel * (ts[1] - ts[0])
    return dict(rel=_rel(x, ref), tol=1e-6)


def check_linear_t(N=3, L=77, K=512, H=4, hd=12, hp=16, seed=0):
    """Value projection with the transposed store (SelfAttention: heads padded hd -> hp rows; NonlinAttention:
    hd = hp = 1 i.e. plain transpose): out[n][drow(col)][l].  tolerance: rel-L2 <= 8e-4; pad rows/cols stay 0."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    M, n_out = N * L, H * hd
    Lk = (L + 7) // 8 * 8
    A = torch.randn(M, K, generator=g).to(H16).to(DEV)
    W = (torch.randn(n_out, K, generator=g) / math.sqrt(K)).to(H16).to(DEV)
    b = torch.randn(n_out, generator=g).to(DEV)
    rows = H * hp
    out = torch.zeros(N, rows, Lk, dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_linear_t(A.data_ptr(), M, K, K, W.data_ptr(), b.data_ptr(), n_out, K, out.data_ptr(),
                                     L, Lk, rows, hd, hp, _s()))
    torch.cuda.synchronize()
    ref = (A.float() @ W.float().t() + b).reshape(N, L, H, hd).permute(0, 2, 3, 1)          # (N, H, hd, L)
    got = out.float().reshape(N, H, hp, Lk)
    pad = int((got[:, :, hd:] != 0).sum() + (got[..., L:] != 0).sum())
    return dict(rel=_rel(got[:, :, :hd, :L], ref), nan=int(torch.isnan(got).sum()), pad_bad=pad, tol=8e-4)


def check_resample(N=3, L=333, ds=2, C=512, seed=0):
    """SimpleDownsample (weighted sum over groups of ds frames, last frame repeated) and SimpleUpsample +
    out_combiner (reference: zipformer.py:873-935).  tolerance: rel-L2 <= 6e-4 (fp16 outputs)."""
    import ctypes as C_
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, L, C, generator=g).to(H16).to(DEV)
    w = torch.softmax(torch.randn(ds, generator=g), 0)
    w4 = (C_.c_float * 4)(*(w.tolist() + [0.0] * (4 - ds)))
    Ld = (L + ds - 1) // ds
    down = torch.full((N, Ld, C), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_downsample(x.data_ptr(), down.data_ptr(), N, L, ds, w4, C, _s()))
    xf = x.float()
    pad = Ld * ds - L
    xp = torch.cat([xf, xf[:, -1:].expand(N, pad, C)], dim=1) if pad else xf
    ref_d = (xp.reshape(N, Ld, ds, C) * w.to(DEV)[None, None, :, None]).sum(2)
    y = torch.randn(N, Ld, C, generator=g).to(H16).to(DEV)
    sc = (torch.rand(C, generator=g) * 0.6 + 0.3).to(DEV)
    up = torch.full((N, L, C), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_upsample_combine(x.data_ptr(), y.data_ptr(), up.data_ptr(), sc.data_ptr(), N, L, ds, C, _s()))
    torch.cuda.synchronize()
    yu = y.float().repeat_interleave(ds, dim=1)[:, :L]
    ref_u = xf + (yu - xf) * sc
    return dict(rel=max(_rel(down.float(), ref_d), _rel(up.float(), ref_u)),
                nan=int(torch.isnan(down.float()).sum() + torch.isnan(up.float()).sum()), tol=6e-4)


def check_stream_prep(rows=1001, C=512, L=37, seed=0):
    """xt = x + temb[row / L].  tolerance: rel-L2 <= 6e-4 (fp16 output)."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(rows, C, generator=g).to(H16).to(DEV)
    temb = torch.randn((rows + L - 1) // L, C, generator=g).to(DEV)
    out = torch.full((rows, C), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_stream_prep(x.data_ptr(), out.data_ptr(), temb.data_ptr(), L, rows, C, _s()))
    torch.cuda.synchronize()
    ref = x.float() + temb[torch.arange(rows, device=DEV) // L]
    return dict(rel=_rel(out.float(), ref), nan=int(torch.isnan(out.float()).sum()), tol=6e-4)


def check_assemble(B=3, T=41, F=100, Ft=100, cfg=1, drop=0, seed=0):
    """Decoder input [x | text | speech | 0] as fp16 with the CFG doubling [uncond ; cond] (reference:
    solver.py:83-98, zipvoice.py:163).  Bit-exact against torch's fp32 -> fp16 rounding."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, T, F, generator=g).to(DEV)
    text = torch.randn(B, T, Ft, generator=g).to(DEV)
    speech = torch.randn(B, T, F, generator=g).to(DEV)
    ldx = (2 * F + Ft + 7) // 8 * 8
    N = 2 * B if cfg else B
    out = torch.full((N, T, ldx), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_assemble_input(x.data_ptr(), text.data_ptr(), speech.data_ptr(), out.data_ptr(), B, T, F, Ft,
                                           ldx, cfg, drop, _s()))
    torch.cuda.synchronize()
    z = torch.zeros(B, T, ldx - 2 * F - Ft, device=DEV)
    cond = torch.cat([x, text, speech, z], dim=2)
    if cfg:
        unc = torch.cat([x, torch.zeros_like(text), torch.zeros_like(speech) if drop else speech, z], dim=2)
        ref = torch.cat([unc, cond], dim=0)
    else:
        ref = cond
    return dict(rel=0.0 if torch.equal(out, ref.to(H16)) else 1.0, nan=int(torch.isnan(out.float()).sum()), tol=0.0)


def check_small_linear(N=5, K=192, O=384, act_in=0, act_out=2, addend=False, seed=0):
    """fp32 time-embedding MLP layer.  tolerance: rel-L2 <= 2e-5 (SwooshR uses the fast exp/log path)."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, K, generator=g).to(DEV)
    W = (torch.randn(O, K, generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn(O, generator=g).to(DEV)
    add = torch.randn(N, O, generator=g).to(DEV) if addend else None
    out = torch.full((N, O), float("nan"), device=DEV)
    _lib.check(lib.zvb_test_small_linear(x.data_ptr(), W.data_ptr(), b.data_ptr(), add.data_ptr() if addend else None,
                                         out.data_ptr(), N, K, O, act_in, act_out, _s()))
    torch.cuda.synchronize()
    xi = _swoosh_r(x) if act_in == 2 else x
    ref = xi @ W.t() + b
    if addend:
        ref = ref + add
    if act_out == 2:
        ref = _swoosh_r(ref)
    return dict(rel=_rel(out, ref), nan=int(torch.isnan(out).sum()), tol=2e-5)


def check_timestep_embedding(N=7, dim=192):
    """[cos(t f) | sin(t f)], f_i = 10000^(-i/half) (reference: zipformer.py:47-69).  tolerance: max-abs <= 2e-5
    (arguments up to ~10 rad in fp32; guidance scales reach 3)."""
    lib = _lib.load()
    t = torch.linspace(0.0, 3.0, N, device=DEV)
    out = torch.full((N, dim), float("nan"), device=DEV)
    _lib.check(lib.zvb_test_timestep_embedding(t.data_ptr(), out.data_ptr(), N, dim, _s()))
    torch.cuda.synchronize()
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, device=DEV, dtype=torch.float32) / half)
    a = t[:, None].double() * f[None].double()
    ref = torch.cat([a.cos(), a.sin()], dim=-1)
    return dict(rel=float((out.double() - ref).abs().max()), nan=int(torch.isnan(out).sum()), tol=2e-5)


def check_masks(N=3, T=333, ds=2, seed=0):
    """mask[:, ::ds] and the attention kernel's excluded-key bit words.  Bit-exact."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = torch.randint(T // 2, T + 1, (N,), generator=g)
    mask = (torch.arange(T)[None, :] >= lens[:, None]).to(torch.uint8).to(DEV)
    L = (T + ds - 1) // ds
    words = 4 * ((L + 127) // 128)
    st = torch.full((N, L), 7, dtype=torch.uint8, device=DEV)
    wd = torch.zeros(N, words, dtype=torch.int32, device=DEV)
    _lib.check(lib.zvb_test_masks(mask.data_ptr(), N, T, ds, st.data_ptr() if ds != 1 else None, wd.data_ptr(), _s()))
    torch.cuda.synchronize()
    ref_m = mask[:, ::ds]
    ok = ds == 1 or torch.equal(st, ref_m)
    bits = torch.ones(N, words * 32, dtype=torch.int64, device=DEV)
    bits[:, :L] = ref_m.long()
    ref_w = (bits.reshape(N, words, 32) << torch.arange(32, device=DEV)).sum(-1)
    got_w = wd.long() & 0xFFFFFFFF
    ok = ok and torch.equal(got_w, ref_w)
    return dict(rel=0.0 if ok else 1.0, tol=0.0)


def assert_ok(name, r):
    assert r.get("nan", 0) == 0, (name, r)
    assert r.get("pad_bad", 0) == 0, (name, r)
    if "maxabs" in r:
        assert r["maxabs"] <= r["tol"], (name, r)
        assert r.get("pad_nonzero", 0) == 0, (name, r)
    else:
        assert r["rel"] <= r["tol"], (name, r)



# ------------------------------------------------------------------------------------------ vocoder pieces
def check_layernorm(rows=301, C=512, masked=True, seed=0):
    """tolerance: rel-L2 <= 8e-4 (fp16 output); masked rows exactly zero"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(rows, C, generator=g) * 2.0 + 0.5).to(H16).to(DEV)
    w = (1.0 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    b = (0.3 * torch.randn(C, generator=g)).to(DEV)
    m = (torch.rand(rows, generator=g) < 0.2).to(torch.uint8).to(DEV) if masked else None
    out = torch.full((rows, C), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_layernorm(x.data_ptr(), out.data_ptr(), w.data_ptr(), b.data_ptr(),
                                      m.data_ptr() if masked else None, rows, C, 1e-6, _s()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w, b, 1e-6)
    if masked:
        ref = ref * (m == 0).unsqueeze(1)
        assert float(out[m != 0].float().abs().max()) == 0.0
    return dict(rel=_rel(out.float(), ref), nan=int(torch.isnan(out.float()).sum()), tol=8e-4)


def check_dwconv_linear(N=3, L=150, C=512, K=7, seed=0):
    """depthwise convolution + bias without activation; tolerance: rel-L2 <= 8e-4 (fp16 output)"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, L, C, generator=g).to(H16).to(DEV)
    w = (torch.randn(C, 1, K, generator=g) / math.sqrt(K)).to(DEV)
    b = (torch.randn(C, generator=g) * 0.2).to(DEV)
    wt = w.reshape(C, K).t().contiguous()
    out = torch.full((N, L, C), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_dwconv_linear(x.data_ptr(), out.data_ptr(), wt.data_ptr(), b.data_ptr(), N, L, C, K, _s()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv1d(x.float().permute(0, 2, 1), w, b, padding=K // 2, groups=C).permute(0, 2, 1)
    return dict(rel=_rel(out.float(), ref), nan=int(torch.isnan(out.float()).sum()), tol=8e-4)


def check_linear_masked(M=700, K=1536, N=512, act=0, resid=True, seed=0):
    """linear (+ exact GELU) + residual with a row mask: masked rows exactly zero; rel-L2 <= 8e-4"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(H16).to(DEV)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(H16).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    R = torch.randn(M, N, generator=g).to(H16).to(DEV) if resid else None
    m = (torch.rand(M, generator=g) < 0.3).to(torch.uint8).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_linear_masked(A.data_ptr(), M, K, K, W.data_ptr(), b.data_ptr(), N, K, act,
                                          R.data_ptr() if resid else None, m.data_ptr(), out.data_ptr(), N, _s()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + b
    if act == 3:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + R.float()
    ref = ref * (m == 0).unsqueeze(1)
    assert float(out[m != 0].float().abs().max()) == 0.0
    return dict(rel=_rel(out.float(), ref), nan=int(torch.isnan(out.float()).sum()), tol=8e-4)


def check_linear_gelu(M=1000, K=512, N=1536, seed=0):
    """pwconv1 of a ConvNeXt block: linear + exact (erf) GELU on the lean epilogue; rel-L2 <= 8e-4"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(H16).to(DEV)
    W = (torch.randn(N, K, generator=g) * 2.0 / math.sqrt(K)).to(H16).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=H16, device=DEV)
    _lib.check(lib.zvb_test_linear(A.data_ptr(), M, K, K, W.data_ptr(), b.data_ptr(), N, K, 0, 3, None, None, None,
                                   out.data_ptr(), N, 0, _s()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(A.float() @ W.float().t() + b)
    return dict(rel=_rel(out.float(), ref), nan=int(torch.isnan(out.float()).sum()), tol=8e-4)


def check_istft(N=3, T=40, seed=0):
    """inverse STFT head (exp / clip / cos / sin -> irfft -> window -> overlap-add / envelope) against torch.istft per
    utterance, ragged lengths; fp32: max-abs <= 2e-5 of the signal peak"""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    lens = torch.tensor([T, max(2, T // 2 + 1), 2][:N], dtype=torch.int32)
    ld = 1028
    S = torch.zeros(N, T, ld)
    S[..., :513] = torch.randn(N, T, 513, generator=g) * 1.5 + 1.0          # some log-magnitudes beyond the clip at log(100)
    S[..., 513:1026] = torch.randn(N, T, 513, generator=g) * 3.0
    win = torch.hann_window(1024)
    Sd, ld_, wd = S.to(DEV), lens.to(DEV), win.to(DEV)
    frames = torch.zeros(N * T, 1024, device=DEV)
    mask = torch.zeros(N * T, dtype=torch.uint8, device=DEV)
    wav = torch.full((N, 256 * (T - 1)), float("nan"), device=DEV)
    _lib.check(lib.zvb_test_istft(Sd.data_ptr(), ld, ld_.data_ptr(), wd.data_ptr(), frames.data_ptr(), mask.data_ptr(),
                                  wav.data_ptr(), N, T, 256, 0, _s()))
    torch.cuda.synchronize()
    err, peak = 0.0, 0.0
    for n in range(N):
        L = int(lens[n])
        mag = torch.clip(torch.exp(S[n, :L, :513].double()), max=1e2)
        p = S[n, :L, 513:1026].double()
        spec = (mag * (torch.cos(p) + 1j * torch.sin(p))).t().unsqueeze(0)
        ref = torch.istft(spec, 1024, 256, 1024, win.double(), center=True)[0]
        got = wav[n].cpu().double()
        err = max(err, float((got[: ref.numel()] - ref).abs().max()))
        peak = max(peak, float(ref.abs().max()))
        assert float(got[ref.numel():].abs().max() if got.numel() > ref.numel() else 0.0) == 0.0
    return dict(rel=err / peak, nan=int(torch.isnan(wav).sum()), tol=2e-5)


ALL = {
    "linear_basic": lambda: check_linear(M=300, K=512, N=272),
    "linear_k48": lambda: check_linear(M=257, K=48, N=512, resid=True),
    "linear_k300": lambda: check_linear(M=130, K=300, N=512),
    "linear_swoosh": lambda: check_linear(M=1000, K=512, N=1152, act=1),
    "linear_big": lambda: check_linear(M=20000, K=1536, N=512, resid=True),
    "linear_bypass": lambda: check_linear(M=3000, K=1536, N=512, resid=True, bypass=True),
    "linear_resid_n192": lambda: check_linear(M=777, K=384, N=192, resid=True),
    "linear_bypass_n192": lambda: check_linear(M=130, K=96, N=192, resid=True, bypass=True),
    "linear_f32_n100": lambda: check_linear(M=333, K=512, N=100, out_mode=1),
    "linear_n1920_swoosh": lambda: check_linear(M=5000, K=512, N=1920, act=1),
    # CTA pairs with the A-stationary tile order (K = 512, several n-tiles, aux-less TMA-store epilogue): odd tile
    # counts, chunk boundaries inside an m-group, a last pair with one all-padding m-tile
    "linear_resident_n384": lambda: check_linear(M=6001, K=512, N=384),
    "linear_resident_swoosh_n1152": lambda: check_linear(M=9000, K=512, N=1152, act=1),
    "linear_resident_n1536": lambda: check_linear(M=40000, K=512, N=1536, act=1),
    "gated_glu_resident": lambda: check_gated(M=6001, n_out=512, mode=2, masked=True),
    # CTA pairs are only used from 2 x 148 m-tiles on (engine.cu: set_grid): the residual, bypass, gated and row-scaled
    # (P.V) epilogues as pairs, with a last pair whose second m-tile is all padding
    "linear_pair_resid": lambda: check_linear(M=38100, K=1536, N=512, resid=True),
    "linear_pair_bypass": lambda: check_linear(M=38100, K=1536, N=512, resid=True, bypass=True),
    "linear_pair_k48": lambda: check_linear(M=38100, K=48, N=512, resid=True),
    "gated_glu_pair": lambda: check_gated(M=38100, n_out=512, mode=2, masked=True),
    "pv_wide_mul_pair": lambda: check_pv(N=30, H=4, L=1219, hd=384, hp=384, per_head=False, mul=True),
    "gated_tanh": lambda: check_gated(M=300, n_out=384, mode=1),
    "gated_tanh96": lambda: check_gated(M=81, K=128, n_out=96, mode=1),
    "gated_glu_masked": lambda: check_gated(M=1000, n_out=512, mode=2, masked=True),
    "linear_bn64": lambda: check_linear(M=128, K=64, N=64, block_n=64),
    # tile widths the wave model picks for single-utterance shapes (no multiple of 64: per-thread store path)
    "linear_bn80_resid": lambda: check_linear(M=2438, K=1920, N=512, block_n=80, resid=True),
    "linear_bn112_bypass": lambda: check_linear(M=2438, K=1536, N=512, block_n=112, resid=True, bypass=True),
    "linear_bn224_swoosh": lambda: check_linear(M=2438, K=512, N=1536, block_n=224, act=1),
    "linear_bn160_swoosh": lambda: check_linear(M=2438, K=512, N=1920, block_n=160, act=1),
    "linear_bn48_f32": lambda: check_linear(M=2438, K=512, N=100, block_n=48, out_mode=1),
    "linear_small_auto": lambda: check_linear(M=2438, K=1152, N=512, resid=True),
    "linear_tail": lambda: check_linear(M=77, K=192, N=640, act=2),
    "attn_small": lambda: check_attn(N=2, H=4, L=100, masked=False),
    "attn_masked": lambda: check_attn(N=3, H=4, L=333, masked=True),
    "attn_long": lambda: check_attn(N=1, H=4, L=1219, masked=True),
    # strong rel-pos bias: |p|.max|E| ~ 15, the softmax shift bound is loose by up to ~2^40 and most rows
    # sit far below it; weights must stay normal fp16 numbers thanks to the 2^12 head-room (looser
    # tolerance: the fp16 bias sum carries ~2^-6 absolute error at magnitudes of 16..32)
    "attn_strong_pos": lambda: check_attn(N=2, H=4, L=333, masked=True, pos_scale=2.5, tol=3e-2),
    # the CUDA-core-bias kernel (attn.cuh, the default)
    "attn_v2_masked": lambda: check_attn(N=3, H=4, L=333, masked=True, tc=False),
    "attn_v2_strong_pos": lambda: check_attn(N=2, H=4, L=333, masked=True, pos_scale=2.5, tol=3e-2, tc=False),
    # small grids split the key tiles of a query tile over a cluster (attn.cuh, CS = 2 / 4; the two cases above already run
    # as pairs with an uneven 1 + 2 split): four CTAs with 2 + 3 + 2 + 3 key tiles, the exact-maximum path over four CTAs,
    # and a grid large enough to stay on one CTA per query tile
    "attn_v2_split4_long": lambda: check_attn(N=1, H=4, L=1219, masked=True, tc=False),
    "attn_v2_split4_strong": lambda: check_attn(N=1, H=4, L=610, masked=True, pos_scale=2.5, tol=3e-2, tc=False),
    "attn_v2_split2_long": lambda: check_attn(N=2, H=4, L=1219, masked=True, tc=False),
    "attn_v2_nosplit": lambda: check_attn(N=16, H=4, L=333, masked=True, tc=False),
    # tensor-core bias: odd / even L (window alignment in the two table copies), one tile, tile boundaries
    "attn_tc_L128": lambda: check_attn(N=2, H=4, L=128, masked=False),
    "attn_tc_L129": lambda: check_attn(N=2, H=4, L=129, masked=True),
    "attn_tc_L610": lambda: check_attn(N=2, H=4, L=610, masked=True),
    "attn_tc_L77": lambda: check_attn(N=3, H=2, L=77, masked=True),
    "pv_heads": lambda: check_pv(N=2, H=4, L=333, per_head=True),
    "pv_wide": lambda: check_pv(N=2, H=4, L=333, hd=384, hp=384, per_head=False),
    "pv_wide96": lambda: check_pv(N=2, H=4, L=81, hd=96, hp=96, per_head=False),
    "pv_wide_mul": lambda: check_pv(N=3, H=4, L=333, hd=384, hp=384, per_head=False, mul=True),
    "pv_wide144_mul": lambda: check_pv(N=2, H=4, L=130, hd=144, hp=144, per_head=False, mul=True),
    "biasnorm": lambda: check_biasnorm(),
    "biasnorm_c192": lambda: check_biasnorm(rows=77, C=192),
    "dwconv31": lambda: check_dwconv(K=31),
    "dwconv15": lambda: check_dwconv(K=15, L=77),
    "dwconv7": lambda: check_dwconv(K=7, L=64, C=128),
    "dwconv9": lambda: check_dwconv(K=9, L=20, C=192),
    "linear_t_heads": lambda: check_linear_t(N=3, L=77, H=4, hd=12, hp=16),
    "linear_t_plain": lambda: check_linear_t(N=2, L=333, H=1, hd=48, hp=48),
    "resample_ds2": lambda: check_resample(N=3, L=333, ds=2),
    "resample_ds4": lambda: check_resample(N=2, L=1219, ds=4),
    "resample_ds4_c128": lambda: check_resample(N=2, L=50, ds=4, C=128),
    "stream_prep": lambda: check_stream_prep(),
    "assemble_cfg_keep": lambda: check_assemble(cfg=1, drop=0),
    "assemble_cfg_drop": lambda: check_assemble(cfg=1, drop=1),
    "assemble_stereo_nocfg": lambda: check_assemble(B=2, T=30, F=200, Ft=100, cfg=0),
    "small_linear_swoosh_out": lambda: check_small_linear(),
    "small_linear_swoosh_in": lambda: check_small_linear(K=192, O=512, act_in=2, act_out=0),
    "small_linear_addend": lambda: check_small_linear(K=192, O=192, act_in=0, act_out=0, addend=True),
    "timestep_embedding": lambda: check_timestep_embedding(),
    "masks_ds1": lambda: check_masks(ds=1),
    "masks_ds2": lambda: check_masks(ds=2),
    "masks_ds4": lambda: check_masks(N=2, T=1219, ds=4),
    "layernorm": lambda: check_layernorm(),
    "layernorm_c256": lambda: check_layernorm(rows=77, C=256, masked=False),
    "dwconv7_linear": lambda: check_dwconv_linear(),
    "linear_masked_resid": lambda: check_linear_masked(),
    "linear_masked_gelu": lambda: check_linear_masked(M=300, K=512, N=768, act=3, resid=False),
    "linear_gelu": lambda: check_linear_gelu(),
    "istft": lambda: check_istft(),
    "istft_long": lambda: check_istft(N=2, T=333, seed=1),
    "cfg_euler": lambda: check_cfg_euler(cfg=1),
    "euler_nocfg": lambda: check_cfg_euler(cfg=0),
}
