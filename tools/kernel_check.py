"""GPU bring-up checks for the individual kernels (run under gpurun; one case per process so a
trap in one kernel cannot poison the others).  Usage: python tools/kernel_check.py <case>|list"""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kosmos-x_b200"))
from kosmosx import _abi, ops  # noqa: E402

dev = "cuda"


def blockmap(err, br=8, bc=8):
    M, N = err.shape
    rs, cs = max(1, M // br), max(1, N // bc)
    out = []
    for i in range(0, M, rs):
        out.append(" ".join(f"{err[i:i+rs, j:j+cs].max().item():8.2e}" for j in range(0, N, cs)))
    return "\n".join(out[:br + 1])


def report(name, got, ref, tol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bad = ~torch.isfinite(got)
    mx = err[~bad].max().item() if (~bad).any() else float("nan")
    ok = (not bad.any()) and mx <= tol
    print(f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={mx:.4e} tol={tol:.1e} nonfinite={int(bad.sum())} ref_absmax={ref.abs().max().item():.3f}")
    if not ok and err.ndim == 2:
        print(blockmap(torch.where(bad, torch.full_like(err, float('inf')), err)))
    return ok


def gemm_case(cg, bn, M=512, N=512, K=256, **kw):
    torch.manual_seed(0)
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
    ops.gemm(a, w, out, cta_group=cg, block_n=bn, **kw)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    return report(f"gemm cg={cg} bn={bn} M={M} N={N} K={K}", out, ref, 2e-3 * math.sqrt(K))


def case_gemm_basic():
    ok = True
    for cg, bn in ((1, 128), (1, 256), (2, 128), (2, 256)):
        ok &= gemm_case(cg, bn, 256, 256, 64)
        ok &= gemm_case(cg, bn, 512, 512, 256)
    return ok


def case_gemm_shapes():
    ok = True
    for cg, bn in ((1, 256), (2, 256), (1, 128), (2, 128)):
        ok &= gemm_case(cg, bn, 2056, 1024, 1024)        # ViT rows (ragged M)
        ok &= gemm_case(cg, bn, 300, 1002, 128)          # ragged M and N
        ok &= gemm_case(cg, bn, 4096, 2048, 2048)
        ok &= gemm_case(cg, bn, 1024, 770, 640)          # N tail, K not multiple of 64*? (640 = 10 blocks)
        ok &= gemm_case(cg, bn, 640, 512, 72)            # K tail (72 = 64 + 8)
    return ok


def case_gemm_epilogue():
    torch.manual_seed(1)
    ok = True
    M, N, K = 1000, 768, 512
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev)
    ref0 = a.float() @ w.float().T + bias
    for cg in (1, 2):
        # bias + gelu -> bf16
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, w, out, bias=bias, act=_abi.KX_ACT_GELU, cta_group=cg)
        ok &= report(f"epi bias+gelu bf16 cg={cg}", out, torch.nn.functional.gelu(ref0), 2e-2)
        # bias + quick gelu
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, w, out, bias=bias, act=_abi.KX_ACT_QUICK_GELU, cta_group=cg)
        ok &= report(f"epi bias+quickgelu bf16 cg={cg}", out, ref0 * torch.sigmoid(1.702 * ref0), 2e-2)
        # bias + residual in place (fp32)
        x = torch.randn(M, N, device=dev)
        x0 = x.clone()
        ops.gemm(a, w, x, bias=bias, res=x, cta_group=cg)
        ok &= report(f"epi bias+res fp32 in-place cg={cg}", x, ref0 + x0, 1e-2)
        # scatter rows + positional add: groups of 250 rows -> stride 300, offset 7; table row = r + 3
        G, stride, off = 250, 300, 7
        tab = torch.randn(G + 3, N, device=dev)
        outs = torch.zeros(4 * stride + off, N, device=dev)
        ops.gemm(a, w, outs, grp=(G, stride, off), add_tab=tab, add_off=3, cta_group=cg)
        ref = torch.zeros_like(outs)
        base = a.float() @ w.float().T
        for g in range(4):
            ref[g * stride + off: g * stride + off + G] = base[g * G:(g + 1) * G] + tab[3:3 + G]
        ok &= report(f"epi scatter+pos fp32 cg={cg}", outs, ref, 1e-2)
        # odd ld fp32 (LM-head style N=1002, ld=1002)
        w2 = (torch.randn(1002, K, device=dev) / math.sqrt(K)).bfloat16()
        o2 = torch.zeros(M, 1002, device=dev)
        ops.gemm(a, w2, o2, cta_group=cg)
        ok &= report(f"epi fp32 N=1002 cg={cg}", o2, a.float() @ w2.float().T, 1e-2)
    # staged (TMA-store) epilogue vs direct stores: ragged M and N, bf16 and fp32, residual in place, both tile widths
    for cg, bn in ((1, 128), (2, 128), (1, 256), (2, 256)):
        for (M2, N2, K2) in ((300, 1000, 192), (2056, 1024, 256), (515, 40, 64)):
            a2 = torch.randn(M2, K2, device=dev).bfloat16()
            w3 = (torch.randn(N2, K2, device=dev) / math.sqrt(K2)).bfloat16()
            b3 = torch.randn(N2, device=dev)
            base = a2.float() @ w3.float().T + b3
            for mode in (0, 1):
                ob = torch.full((M2 + 3, N2), 7.0, device=dev, dtype=torch.bfloat16)     # 3 guard rows must stay untouched
                ops.gemm(a2, w3, ob, bias=b3, act=_abi.KX_ACT_GELU, cta_group=cg, block_n=bn, epi_mode=mode, M=M2)
                good = report(f"epi mode={mode} bf16 gelu cg={cg} bn={bn} {M2}x{N2}x{K2}", ob[:M2],
                              torch.nn.functional.gelu(base), 2e-2)
                ok &= good and bool((ob[M2:] == 7.0).all())
                xr = torch.randn(M2 + 3, N2, device=dev)
                x0 = xr.clone()
                ops.gemm(a2, w3, xr, bias=b3, res=xr, cta_group=cg, block_n=bn, epi_mode=mode, M=M2)
                good = report(f"epi mode={mode} fp32 res cg={cg} bn={bn} {M2}x{N2}x{K2}", xr[:M2], base + x0[:M2], 1e-2)
                ok &= good and bool(torch.equal(xr[M2:], x0[M2:]))
    # fast erf-GELU against torch's exact erf GELU in fp32 (identity weights isolate the activation)
    K3 = 256
    eye = torch.eye(K3, device=dev).bfloat16()
    xs = torch.linspace(-9, 9, 4096 * K3, device=dev).view(4096, K3).bfloat16()
    og = torch.empty(4096, K3, device=dev)
    ops.gemm(xs, eye, og, act=_abi.KX_ACT_GELU)
    ok &= report("erf-GELU approximation (fp32 out)", og, torch.nn.functional.gelu(xs.float()), 6e-6)
    return ok


def xpos_ref(T, device):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import kosmos_oracle as ko
    xp = ko.XPOS(64)
    scale, sin, cos = xp.tables(T)
    return xp, scale.to(device), sin.to(device), cos.to(device)


def case_gemm_trans():
    """Backward GEMMs on operands where they lie: dgrad = dY . W (W as [K', N']), wgrad = dY^T . X (both [K', *])."""
    torch.manual_seed(11)
    ok = True
    for (M, N, K) in ((512, 256, 384), (2048, 2048, 1024), (300, 200, 136), (4096, 1024, 4096), (130, 64, 72)):
        dy = torch.randn(M, N, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(N)).bfloat16()
        x = torch.randn(M, K, device=dev).bfloat16()
        for cg, bn in ((2, 256), (1, 128)):
            dx = torch.full((M + 2, K), 7.0, device=dev, dtype=torch.bfloat16)
            ops.gemm(dy, w, dx, b_trans=True, cta_group=cg, block_n=bn, M=M)           # dX[M,K] = dY[M,N] . W[N,K]
            good = report(f"dgrad bf16 cg={cg} bn={bn} {M}x{N}x{K}", dx[:M], dy.float() @ w.float(), 3e-2)
            ok &= good and bool((dx[M:] == 7.0).all())
            dw = torch.zeros(N, K, device=dev)
            ops.gemm(dy, x, dw, a_trans=True, b_trans=True, cta_group=cg, block_n=bn)  # dW[N,K] = dY^T . X
            ref = dy.float().T @ x.float()
            ok &= report(f"wgrad fp32 cg={cg} bn={bn} {M}x{N}x{K}", dw / math.sqrt(M), ref / math.sqrt(M), 1e-2)
            ops.gemm(dy, x, dw, res=dw, a_trans=True, b_trans=True, cta_group=cg, block_n=bn)   # accumulate: dW += dY^T . X
            ok &= report(f"wgrad accumulate cg={cg} bn={bn} {M}x{N}x{K}", dw / math.sqrt(M), 2 * ref / math.sqrt(M), 2e-2)
    # LM-head shapes: vocab not a multiple of 64, dlogits rows padded to a multiple of 64 columns with zeros
    M, V, D = 384, 1002, 256
    Vp = (V + 63) // 64 * 64
    dl = torch.zeros(M, Vp, device=dev, dtype=torch.bfloat16)
    dl[:, :V] = torch.randn(M, V, device=dev).bfloat16()
    wout = (torch.randn(V, D, device=dev) / math.sqrt(D)).bfloat16()
    h = torch.randn(M, D, device=dev).bfloat16()
    dh = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(dl[:, :V], wout, dh, b_trans=True)
    ok &= report("LM head dgrad (K'=1002)", dh, dl[:, :V].float() @ wout.float(), 6e-2)
    dwo = torch.zeros(V, D, device=dev)
    ops.gemm(dl[:, :V], h, dwo, a_trans=True, b_trans=True)
    ok &= report("LM head wgrad (M'=1002)", dwo / math.sqrt(M), dl[:, :V].float().T @ h.float() / math.sqrt(M), 1e-2)
    return ok


def case_train_elementwise():
    """Training-step glue kernels against autograd / torch.optim on the same inputs."""
    torch.manual_seed(21)
    ok = True
    F = torch.nn.functional
    # ---- LayerNorm backward: fp32 residual form (dres add, bf16 copy, column sums) and bf16 forms, n = 2048 / 8192 / ragged 200
    # (the wide GELU form is its own kernel, two rows in flight per CTA: more rows than SMs, fewer, one, and rows narrower than 8192
    # whose second column chunk is partly / entirely dead)
    for (rows, n, xdt, act) in ((1000, 2048, torch.float32, 0), (515, 8192, torch.bfloat16, 1), (300, 200, torch.bfloat16, 0),
                                (64, 4096, torch.float32, 0), (100, 8192, torch.bfloat16, 1), (1, 8192, torch.bfloat16, 1),
                                (1200, 8192, torch.bfloat16, 1), (450, 6144, torch.bfloat16, 1), (333, 4096, torch.bfloat16, 1),
                                (200, 2056, torch.bfloat16, 1), (150, 1024, torch.bfloat16, 1)):
        x = torch.randn(rows, n, device=dev) * 1.5 + 0.3
        x = x.to(xdt)
        gamma = torch.rand(n, device=dev) + 0.5
        beta = torch.randn(n, device=dev)
        dy = (torch.randn(rows, n, device=dev) / math.sqrt(n)).bfloat16()
        xr = x.float().clone().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        a = F.gelu(xr) if act else xr
        y = F.layer_norm(a, (n,), gr, br, 1e-5)
        y.backward(dy.float())
        P = ops.ln_bwd_partials(rows)
        part = torch.empty(3, P, n, device=dev)
        dg, db = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        if xdt == torch.float32:
            dres = torch.randn(rows, n, device=dev) / math.sqrt(n)
            dx = dres.clone()
            dxb = torch.zeros(rows, n, device=dev, dtype=torch.bfloat16)
            dcol = torch.zeros(n, device=dev)
            ops.layernorm_bwd(x, dy, gamma, dx, dg, db, part, act=act, dres=dx, dxb=dxb, d_colsum=dcol)
            ok &= report(f"ln_bwd fp32 dx {rows}x{n}", dx * math.sqrt(n), (xr.grad + dres) * math.sqrt(n), 2e-3)
            ok &= report(f"ln_bwd bf16 copy {rows}x{n}", dxb.float() * math.sqrt(n), (xr.grad + dres) * math.sqrt(n), 3e-2)
            ok &= report(f"ln_bwd colsum {rows}x{n}", dcol, (xr.grad + dres).sum(0), 2e-3)
        else:
            dx = torch.zeros(rows, n, device=dev, dtype=torch.bfloat16)
            dcol = torch.zeros(n, device=dev)
            ops.layernorm_bwd(x, dy, gamma, dx, dg, db, part, act=act, d_colsum=dcol)
            ok &= report(f"ln_bwd bf16 dx act={act} {rows}x{n}", dx.float() * math.sqrt(n), xr.grad * math.sqrt(n), 3e-2)
            ok &= report(f"ln_bwd bf16 colsum {rows}x{n}", dcol, dx.float().sum(0), 2e-3)
        ok &= report(f"ln_bwd dgamma {rows}x{n}", dg, gr.grad, 3e-3)
        ok &= report(f"ln_bwd dbeta {rows}x{n}", db, br.grad, 3e-3)
        ops.layernorm_bwd(x, dy, gamma, dx if xdt != torch.float32 else dx.clone(), dg, db, part, act=act, accumulate=True)
        ok &= report(f"ln_bwd dgamma accumulate {rows}x{n}", dg, 2 * gr.grad, 6e-3)
    # ---- LN(gelu(u)) forward
    # (wide rows run the two-rows-in-flight kernel: more rows than SMs, fewer, one, a partly / entirely dead second column chunk)
    for (rows, n) in ((700, 8192), (333, 256), (100, 8192), (1, 8192), (1500, 8192), (450, 6144), (333, 4096), (200, 2056), (150, 2048)):
        u = torch.randn(rows, n, device=dev).bfloat16()
        gamma, beta = torch.rand(n, device=dev) + 0.5, torch.randn(n, device=dev)
        out = torch.zeros(rows, n, device=dev, dtype=torch.bfloat16)
        ops.act_layernorm(u, gamma, beta, out)
        ok &= report(f"act_layernorm {rows}x{n}", out, F.layer_norm(F.gelu(u.float()), (n,), gamma, beta, 1e-5), 7e-2)   # bf16 output, |y| up to 12
    # ---- column sums
    for (rows, n) in ((4096, 6144), (1000, 136), (77, 2048)):
        xx = torch.randn(rows, n, device=dev).bfloat16()
        out = torch.ones(n, device=dev)
        ops.colsum(xx, out)
        ok &= report(f"colsum {rows}x{n}", out, xx.float().sum(0) + 1, 2e-3)
    # ---- xPos backward = transpose of the forward rotation: <R x, y> == <x, R^T y>
    T, D, B = 96, 256, 2
    scale = (torch.arange(0, 64, 2, device=dev) + 0.4 * 64) / (1.4 * 64)
    inv_freq = 1.0 / (10000 ** (torch.arange(0, 32, device=dev) / 32))
    tabs = ops.xpos_tables(scale.float(), inv_freq.float(), T, (-T) // 2, 512.0, dev)
    def rot(z, c, s_):          # forward rotation in fp32 on [B*T, D]
        z = z.view(B, T, D // 64, 32, 2)
        c, s_ = c.view(1, T, 1, 32), s_.view(1, T, 1, 32)
        return torch.stack([z[..., 0] * c - z[..., 1] * s_, z[..., 1] * c + z[..., 0] * s_], -1).view(B * T, D)
    dqkv = torch.randn(B * T, 3 * D, device=dev).bfloat16()
    want_q = torch.autograd.functional.vjp(lambda z: rot(z, tabs[0], tabs[1]), torch.zeros(B * T, D, device=dev), dqkv[:, :D].float())[1]
    want_k = torch.autograd.functional.vjp(lambda z: rot(z, tabs[2], tabs[3]), torch.zeros(B * T, D, device=dev), dqkv[:, D:2 * D].float())[1]
    v_before = dqkv[:, 2 * D:].clone()
    ops.xpos_bwd(dqkv, D, T, tabs)
    ok &= report("xpos_bwd dq", dqkv[:, :D], want_q, 3e-2)
    ok &= report("xpos_bwd dk", dqkv[:, D:2 * D], want_k, 3e-2)
    ok &= bool(torch.equal(dqkv[:, 2 * D:], v_before))
    # ---- cross-entropy over the text rows (two images) + d(logits)
    Bc, t_text, n_img, V = 3, 20, 8, 1002
    rows_img = (2, 15)                                          # spliced starts: in front of text tokens 2 and 7
    Tc = t_text + 2 * n_img
    tok = torch.randint(0, V, (Bc, t_text), device=dev)
    logits = torch.randn(Bc * Tc, V, device=dev) * 2
    is_img = torch.zeros(Tc, dtype=torch.bool)
    for r0 in rows_img:
        is_img[r0:r0 + n_img] = True
    text_rows = (~is_img).nonzero().flatten().tolist()
    tgt = torch.full((Bc, Tc), -100, device=dev, dtype=torch.long)
    for ti, t in enumerate(text_rows):
        if ti + 1 < t_text and not (t + 1 < Tc and is_img[t + 1]):
            tgt[:, t] = tok[:, ti + 1]
    count = int((tgt >= 0).sum())
    lr_ = logits.clone().requires_grad_(True)
    loss = F.cross_entropy(lr_, tgt.view(-1), ignore_index=-100, reduction="sum")
    (loss / count).backward()
    Vp = (V + 63) // 64 * 64
    dl = torch.full((Bc * Tc, Vp), 7.0, device=dev, dtype=torch.bfloat16)
    acc = torch.zeros(2, device=dev)
    tg_dev = torch.empty(Bc, Tc, dtype=torch.long, device=dev)
    cnt = torch.zeros(1, device=dev)
    ops.loss_targets(tok, tg_dev, cnt, img_rows=rows_img, n_img=n_img, rule="next_token")
    ok &= bool(torch.equal(tg_dev, tgt)) and bool(int(cnt.item()) == count)
    ops.ce_fwd_bwd(logits, tg_dev.view(-1), acc, count=cnt, dlogits=dl)
    ok &= report("ce loss sum", acc[:1], loss.detach().view(1), 2e-3 * count)
    ok &= bool(int(acc[1].item()) == count)
    ok &= report("ce dlogits", dl[:, :V].float() * count, lr_.grad * count, 1e-2)
    ok &= bool((dl[:, V:] == 0).all())
    # ---- embedding / position backward
    D2 = 128
    dx0 = torch.randn(Bc, Tc, D2, device=dev)
    tok[0, 3] = 1                                               # a padding token: no gradient to its row
    d_emb, d_pos = torch.zeros(V, D2, device=dev), torch.zeros(Tc + 2, D2, device=dev)
    ops.embed_bwd(dx0.view(-1, D2), tok, d_emb, d_pos, img_rows=rows_img, n_img=n_img)
    want_e = torch.zeros(V, D2, device=dev)
    want_e.index_add_(0, tok.view(-1), dx0[:, ~is_img].reshape(-1, D2))
    want_e[1] = 0
    ok &= report("embed_bwd d_embed", d_emb, want_e, 1e-5)
    want_p = torch.zeros(Tc + 2, D2, device=dev)
    want_p[2:] = dx0.sum(0)
    ok &= report("embed_bwd d_pos", d_pos, want_p, 1e-5)
    # alias_positions: the row of text token i also feeds pos[i + 2] (torchscale's in-place `x += positions`)
    d_pos2 = torch.zeros(Tc + 2, D2, device=dev)
    ops.embed_bwd(dx0.view(-1, D2), tok, None, d_pos2, img_rows=rows_img, n_img=n_img, alias_positions=True)
    want_p2 = want_p.clone()
    want_p2[2:2 + t_text] += dx0[:, ~is_img].sum(0)
    ok &= report("embed_bwd d_pos (aliased positions)", d_pos2, want_p2, 1e-5)
    # ---- gradient norm, clipping scale, AdamW, Lion
    n = 1000003
    pw = torch.randn(n, device=dev); g = torch.randn(n, device=dev) * 3
    ss = torch.zeros(1, device=dev); sc = torch.zeros(1, device=dev); nrm = torch.zeros(1, device=dev)
    ops.sumsq(g, ss)
    ops.clip_scale(ss, 1.0, 0.5, sc, nrm)
    ok &= report("grad norm", nrm, (g * 0.5).norm().view(1), 1e-2)
    ok &= report("clip scale", sc, (0.5 * torch.clamp(1.0 / ((g * 0.5).norm() + 1e-6), max=1.0)).view(1), 1e-7)
    for name in ("adamw", "lion"):
        pr = torch.nn.Parameter(pw.clone())
        if name == "adamw":
            opt = torch.optim.AdamW([pr], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
        m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
        mine = pw.clone(); wb = torch.zeros(n, device=dev, dtype=torch.bfloat16)
        ref_m = torch.zeros(n, device=dev)
        for step in (1, 2, 3):
            gs = g * (0.5 + step) * sc
            if name == "adamw":
                pr.grad = gs.clone()
                opt.step()
                ops.adamw_step(mine, g * (0.5 + step), m, v, wb, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, step=step, grad_scale=sc)
            else:
                with torch.no_grad():                     # lion_pytorch.Lion update rule
                    pr.mul_(1 - 1e-3 * 0.1)
                    pr.add_(torch.sign(ref_m * 0.9 + gs * 0.1), alpha=-1e-3)
                    ref_m.mul_(0.99).add_(gs, alpha=0.01)
                ops.lion_step(mine, g * (0.5 + step), m, wb, lr=1e-3, betas=(0.9, 0.99), weight_decay=0.1, grad_scale=sc)
        ok &= report(f"{name} 3 steps", mine, pr.detach(), 2e-5)
        ok &= report(f"{name} bf16 copy", wb, mine.bfloat16(), 0.0)
    return ok


def case_perceiver_bwd():
    """Perceiver cross-attention backward and the small resampler helpers against autograd."""
    torch.manual_seed(41)
    ok = True
    F = torch.nn.functional
    for (B, H, nq, nkv) in ((2, 8, 64, 321), (3, 2, 64, 81), (1, 2, 40, 130)):
        inner = H * 64
        q0 = (torch.randn(B * nq, inner, device=dev) * 0.7).bfloat16()
        kv0 = (torch.randn(B * nkv, 2 * inner, device=dev) * 0.7).bfloat16()
        d_out = torch.randn(B * nq, inner, device=dev).bfloat16()
        q = q0.float().requires_grad_(True)
        kv = kv0.float().requires_grad_(True)
        qh = q.view(B, nq, H, 64).transpose(1, 2) * 0.125
        kh = kv[:, :inner].reshape(B, nkv, H, 64).transpose(1, 2)
        vh = kv[:, inner:].reshape(B, nkv, H, 64).transpose(1, 2)
        sim = qh @ kh.transpose(-1, -2)
        o_ref = ((sim - sim.amax(-1, keepdim=True).detach()).softmax(-1) @ vh).transpose(1, 2).reshape(B * nq, inner)
        o_ref.backward(d_out.float())
        out = torch.zeros(B * nq, inner, device=dev, dtype=torch.bfloat16)
        ops.perceiver_attention(q0, kv0, out, batch=B, heads=H, n_q=nq, n_kv=nkv, v_col_off=inner, scale=0.125)
        dq = torch.full((B * nq, inner), 9.0, device=dev, dtype=torch.bfloat16)
        dkv = torch.full((B * nkv, 2 * inner), 9.0, device=dev, dtype=torch.bfloat16)
        ops.perceiver_attention_bwd(q0, kv0, out, d_out, dq, dkv, batch=B, heads=H, n_q=nq, n_kv=nkv, v_col_off=inner, scale=0.125)
        tag = f"B={B} H={H} nq={nq} nkv={nkv}"
        ok &= report(f"perceiver xattn bwd dq {tag}", dq, q.grad, 2e-2)
        ok &= report(f"perceiver xattn bwd dk {tag}", dkv[:, :inner], kv.grad[:, :inner], 2e-2)
        ok &= report(f"perceiver xattn bwd dv {tag}", dkv[:, inner:], kv.grad[:, inner:], 3e-2)
    # GELU forward / backward (bf16)
    u = (torch.randn(300, 4096, device=dev) * 2).bfloat16()
    dm = torch.randn(300, 4096, device=dev).bfloat16()
    ur = u.float().requires_grad_(True)
    F.gelu(ur).backward(dm.float())
    mid = torch.zeros_like(u); du = torch.zeros_like(u)
    ops.gelu_fwd(u, mid); ops.gelu_bwd(u, dm, du)
    ok &= report("gelu_fwd", mid, F.gelu(u.float()), 4e-2)
    ok &= report("gelu_bwd", du, ur.grad, 3e-2)
    # row gather (fp32 and bf16 sources, accumulate) and row sums
    src = torch.randn(5 * 30, 256, device=dev)
    dst = torch.zeros(5 * 8, 256, device=dev, dtype=torch.bfloat16)
    ops.gather_rows(src, dst, grp=(8, 30, 7))
    want = src.view(5, 30, 256)[:, 7:15].reshape(40, 256)
    ok &= report("gather_rows fp32 -> bf16", dst, want.bfloat16(), 0.0)
    ops.gather_rows(src.bfloat16(), dst, grp=(8, 30, 2), accumulate=True)
    want2 = want.bfloat16().float() + src.bfloat16().float().view(5, 30, 256)[:, 2:10].reshape(40, 256)
    ok &= report("gather_rows bf16 accumulate", dst, want2.bfloat16(), 0.0)
    mat = torch.randn(6, 64 * 128, device=dev)
    outv = torch.zeros(64 * 128, device=dev)
    ops.sum_rows_f32(mat, outv)
    ok &= report("sum_rows_f32", outv.view(1, -1), mat.sum(0).view(1, -1), 1e-5)
    # LayerNorm backward with a pre-added row (x + media_pos_emb[i]) and its column-sum gradient
    rows, n = 514, 1024
    x = torch.randn(rows, n, device=dev); pa = torch.randn(n, device=dev)
    gamma = torch.rand(n, device=dev) + 0.5; beta = torch.randn(n, device=dev)
    dy = (torch.randn(rows, n, device=dev) / 32).bfloat16()
    par = pa.clone().requires_grad_(True); gr = gamma.clone().requires_grad_(True)
    F.layer_norm(x + par, (n,), gr, beta, 1e-5).backward(dy.float())
    part = torch.empty(3, ops.ln_bwd_partials(rows), n, device=dev)
    dxs = torch.empty(rows, n, device=dev); dg = torch.zeros(n, device=dev); db = torch.zeros(n, device=dev); dpa = torch.zeros(n, device=dev)
    ops.layernorm_bwd(x, dy, gamma, dxs, dg, db, part, pre_add=pa, d_colsum=dpa, accumulate=True)
    ok &= report("ln_bwd pre_add: d(pre_add) = colsum(dx)", dpa.view(1, -1), par.grad.view(1, -1), 2e-3)
    ok &= report("ln_bwd pre_add: dgamma", dg.view(1, -1), gr.grad.view(1, -1), 3e-3)
    return ok


def case_attn_bwd():
    """Flash attention backward (with and without the fused xPos transpose) against autograd on the eager formula."""
    torch.manual_seed(31)
    ok = True
    for (B, H, T, causal, use_xpos) in ((2, 2, 256, True, False), (1, 3, 200, True, True), (2, 2, 114, True, True),
                                        (1, 2, 384, False, False), (1, 32, 1024, True, True)):
        D = H * 64
        M = B * T
        qkv = (torch.randn(M, 3 * D, device=dev) * 0.8).bfloat16()
        d_out = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
        scale = 0.125
        tabs = None
        if use_xpos:
            sc_ = (torch.arange(0, 64, 2, device=dev) + 0.4 * 64) / (1.4 * 64)
            inv_freq = 1.0 / (10000 ** (torch.arange(0, 32, device=dev) / 32))
            tabs = ops.xpos_tables(sc_.float(), inv_freq.float(), T, (-T) // 2, 512.0, dev)

        def rot(z, c, s_):
            z = z.view(B, T, H, 32, 2)
            c, s_ = c.view(1, T, 1, 32), s_.view(1, T, 1, 32)
            return torch.stack([z[..., 0] * c - z[..., 1] * s_, z[..., 1] * c + z[..., 0] * s_], -1).view(M, D)

        # the kernel differentiates w.r.t. the UN-rotated q, k when tables are given; its inputs are the rotated ones
        q0 = qkv[:, :D].float().clone().requires_grad_(True)
        k0 = qkv[:, D:2 * D].float().clone().requires_grad_(True)
        v0 = qkv[:, 2 * D:].float().clone().requires_grad_(True)
        qr = rot(q0, tabs[0], tabs[1]) if use_xpos else q0
        kr = rot(k0, tabs[2], tabs[3]) if use_xpos else k0
        qkv_rot = torch.cat([qr.detach(), kr.detach(), v0.detach()], 1).bfloat16().contiguous()
        # reference on the bf16-rounded rotated operands (what the kernel reads), gradient routed through the rotation
        qb = qr + (qkv_rot[:, :D].float() - qr).detach()
        kb = kr + (qkv_rot[:, D:2 * D].float() - kr).detach()
        def heads(t):
            return t.view(B, T, H, 64).transpose(1, 2)
        sc = heads(qb) @ heads(kb).transpose(-1, -2) * scale
        if causal:
            sc = sc + torch.triu(torch.full((T, T), float("-inf"), device=dev), 1)
        o_ref = (sc.softmax(-1) @ heads(v0)).transpose(1, 2).reshape(M, D)
        o_ref.backward(d_out.float())
        out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
        lse = torch.full((H, B, ops.lse_pad(T)), float("nan"), device=dev)
        ops.attention(qkv_rot[:, :D], qkv_rot[:, D:2 * D], qkv_rot[:, 2 * D:], out, batch=B, heads=H, seq_len=T, causal=causal,
                      scale=scale, lse_out=lse)
        ok &= report(f"attn fwd (lse path) B={B} H={H} T={T}", out, o_ref.detach(), 2e-2)
        lse_ref = torch.logsumexp(sc.detach(), -1) * 1.4426950408889634          # [B, H, T] in log2 units
        ok &= report(f"attn lse B={B} H={H} T={T}", lse[:, :, :T], lse_ref.transpose(0, 1), 2e-2)
        dqkv = torch.full((M, 3 * D), 9.0, device=dev, dtype=torch.bfloat16)
        acc = torch.empty(M, D, device=dev)
        delta = torch.empty(*lse.shape, 2, device=dev)
        ops.attention_bwd(qkv_rot[:, :D], qkv_rot[:, D:2 * D], qkv_rot[:, 2 * D:], out, d_out, lse, dqkv[:, :D], dqkv[:, D:2 * D],
                          dqkv[:, 2 * D:], acc, delta, batch=B, heads=H, seq_len=T, causal=causal, scale=scale, xpos=tabs)
        torch.cuda.synchronize()
        tag = f"B={B} H={H} T={T} causal={causal} xpos={use_xpos}"
        ok &= report(f"attn_bwd dv {tag}", dqkv[:, 2 * D:], v0.grad, 3e-2)
        ok &= report(f"attn_bwd dk {tag}", dqkv[:, D:2 * D], k0.grad, 3e-2)
        ok &= report(f"attn_bwd dq {tag}", dqkv[:, :D], q0.grad, 3e-2)
    return ok


def case_attn_dropout():
    """Attention dropout: the mask generator (both layouts agree, keep fraction, determinism) and both flash kernels against
    autograd through the eager formula with the SAME mask: out = ((M o P) / keep) V, P = softmax over all scores."""
    torch.manual_seed(41)
    ok = True
    keep = 3686 / 4096
    for (B, H, T) in ((2, 2, 256), (1, 3, 200), (2, 2, 84), (1, 4, 640)):
        D, M = H * 64, B * T
        words = ops.attn_dropout_mask_words(B, H, T)
        rows = torch.zeros(words, dtype=torch.int32, device=dev); keys = torch.zeros(words, dtype=torch.int32, device=dev)
        ops.attn_dropout_masks(rows, keys, p=0.1, site=6, seed=99, batch=B, heads=H, seq_len=T)
        rows2 = torch.zeros_like(rows); keys2 = torch.zeros_like(keys)
        ops.attn_dropout_masks(rows2, keys2, p=0.1, site=6, seed=99, batch=B, heads=H, seq_len=T)
        mr, mk = ops.unpack_attn_row_mask(rows, B, H, T), ops.unpack_attn_dropout_mask(keys, B, H, T)
        tri = torch.tril(torch.ones(T, T, dtype=torch.bool, device=dev))
        same_layouts = bool(torch.equal(mr[:, :, tri], mk[:, :, tri]))
        frac = mr[:, :, tri].float().mean().item()
        ops.attn_dropout_masks(rows2, keys2, p=0.1, site=7, seed=99, batch=B, heads=H, seq_len=T)
        other = not torch.equal(ops.unpack_attn_row_mask(rows2, B, H, T)[:, :, tri], mr[:, :, tri])
        good = same_layouts and abs(frac - keep) < 0.01 and other
        print(f"[{'OK' if good else 'FAIL'}] attn dropout masks B={B} H={H} T={T}: layouts agree={same_layouts} keep fraction={frac:.4f} (want {keep:.4f}) site-dependent={other}")
        ok &= good
        qkv = (torch.randn(M, 3 * D, device=dev) * 0.8).bfloat16()
        d_out = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
        q0 = qkv[:, :D].float().clone().requires_grad_(True)
        k0 = qkv[:, D:2 * D].float().clone().requires_grad_(True)
        v0 = qkv[:, 2 * D:].float().clone().requires_grad_(True)
        heads = lambda t: t.view(B, T, H, 64).transpose(1, 2)
        sc = heads(q0) @ heads(k0).transpose(-1, -2) * 0.125 + torch.triu(torch.full((T, T), float("-inf"), device=dev), 1)
        pm = sc.softmax(-1) * (mr & tri).float() / keep
        o_ref = (pm @ heads(v0)).transpose(1, 2).reshape(M, D)
        o_ref.backward(d_out.float())
        out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
        lse = torch.full((H, B, ops.lse_pad(T)), float("nan"), device=dev)
        ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, batch=B, heads=H, seq_len=T, causal=True, scale=0.125,
                      lse_out=lse, drop_p=0.1, row_mask=rows)
        ok &= report(f"attn fwd with dropout B={B} H={H} T={T}", out, o_ref.detach(), 2e-2)
        dqkv = torch.full((M, 3 * D), 9.0, device=dev, dtype=torch.bfloat16)
        acc = torch.empty(M, D, device=dev); delta = torch.empty(*lse.shape, 2, device=dev)
        ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, d_out, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                          acc, delta, batch=B, heads=H, seq_len=T, causal=True, scale=0.125, drop_p=0.1, drop_mask=keys)
        torch.cuda.synchronize()
        ok &= report(f"attn_bwd with dropout dv B={B} H={H} T={T}", dqkv[:, 2 * D:], v0.grad, 3e-2)
        ok &= report(f"attn_bwd with dropout dk B={B} H={H} T={T}", dqkv[:, D:2 * D], k0.grad, 3e-2)
        ok &= report(f"attn_bwd with dropout dq B={B} H={H} T={T}", dqkv[:, :D], q0.grad, 3e-2)
    return ok


def case_xpos():
    ok = True
    for T in (5, 114, 2048):
        xp, S, sin, cos = xpos_ref(T, dev)
        inv_freq = (1.0 / (10000 ** (torch.arange(0, 32) / 32))).to(dev)
        tabs = ops.xpos_tables(xp.scale.to(dev), inv_freq, T, -(T // 2) if T % 2 == 0 else -((T + 1) // 2), 512.0, dev)
        # python floor: -(T)//2
        mp = (-T) // 2
        tabs = ops.xpos_tables(xp.scale.to(dev), inv_freq, T, mp, 512.0, dev)
        ok &= report(f"xpos q_cos T={T}", tabs[0], cos * S, 2e-6 * max(1.0, (cos * S).abs().max().item()))
        ok &= report(f"xpos q_sin T={T}", tabs[1], sin * S, 2e-6 * max(1.0, (sin * S).abs().max().item()))
        ok &= report(f"xpos k_cos T={T}", tabs[2], cos / S, 2e-6 * max(1.0, (cos / S).abs().max().item()))
        ok &= report(f"xpos k_sin T={T}", tabs[3], sin / S, 2e-6 * max(1.0, (sin / S).abs().max().item()))
    return ok


def case_gemm_qkv():
    torch.manual_seed(2)
    ok = True
    B, T, d = 2, 114, 256
    M = B * T
    a = torch.randn(M, d, device=dev).bfloat16()
    w = (torch.randn(3 * d, d, device=dev) / math.sqrt(d)).bfloat16()
    bias = torch.randn(3 * d, device=dev)
    xp, S, sin, cos = xpos_ref(T, dev)
    inv_freq = (1.0 / (10000 ** (torch.arange(0, 32) / 32))).to(dev)
    tabs = ops.xpos_tables(xp.scale.to(dev), inv_freq, T, (-T) // 2, 512.0, dev)
    for cg in (1, 2):
        out = torch.zeros(M, 3 * d, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, w, out, bias=bias, xpos=tuple(tabs), seq_len=T, cta_group=cg)
        ref = a.float() @ w.float().T + bias
        q, k, v = ref[:, :d], ref[:, d:2 * d], ref[:, 2 * d:]
        H = d // 64
        xpd = xp.to(dev)

        def rot(t, down):
            t = t.view(B, T, H, 64).transpose(1, 2).reshape(B * H, T, 64)
            t = xpd(t, offset=0, downscale=down)
            return t.view(B, H, T, 64).transpose(1, 2).reshape(M, d)

        ref2 = torch.cat([rot(q, False), rot(k, True), v], dim=1)
        ok &= report(f"qkv xpos epilogue cg={cg}", out, ref2, 3e-2)
    return ok


def attn_ref(q, k, v, B, H, T, causal, scale):
    qh = q.float().view(B, T, H, 64).transpose(1, 2)
    kh = k.float().view(B, T, H, 64).transpose(1, 2)
    vh = v.float().view(B, T, H, 64).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    if causal:
        s = s + torch.triu(torch.full((T, T), float("-inf"), device=q.device), 1)
    p = s.softmax(-1)
    return (p @ vh).transpose(1, 2).reshape(B * T, H * 64)


def case_attn():
    torch.manual_seed(3)
    ok = True
    for (B, H, T, causal, qs) in ((1, 1, 128, False, 1), (1, 1, 128, True, 1), (2, 2, 114, True, 1), (2, 3, 257, False, 1),
                                  (1, 2, 512, True, 1), (2, 4, 2048, True, 1), (1, 2, 300, True, 1),
                                  (1, 2, 640, True, 1), (2, 2, 256, False, 1), (3, 2, 1000, True, 1), (1, 1, 1, True, 1),
                                  # large scores: the running maximum keeps growing -> exercises the O/l rescale path
                                  (2, 2, 1024, True, 8), (1, 2, 777, False, 8), (1, 3, 2048, True, 4)):
        qkv = torch.randn(B * T, 3 * H * 64, device=dev)
        qkv[:, :H * 64] *= qs
        # make later keys systematically larger so the maximum rises block after block
        if qs > 1:
            qkv[:, H * 64:2 * H * 64] *= torch.linspace(0.2, 1.5, B * T, device=dev)[:, None]
        qkv = qkv.bfloat16()
        q, k, v = qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:]
        out = torch.full((B * T, H * 64), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.attention(q, k, v, out, batch=B, heads=H, seq_len=T, causal=causal, scale=0.125)
        torch.cuda.synchronize()
        ref = attn_ref(q, k, v, B, H, T, causal, 0.125)
        ok &= report(f"attn B={B} H={H} T={T} causal={causal} qscale={qs}", out, ref, 2e-2)
    return ok


def case_layernorm():
    torch.manual_seed(4)
    ok = True
    for (rows, n, dt) in ((100, 128, torch.float32), (2056, 1024, torch.float32), (1000, 2048, torch.float32),
                          (777, 2048, torch.bfloat16), (513, 8192, torch.bfloat16), (64, 256, torch.bfloat16),
                          (10, 16384, torch.float32)):
        x = (torch.randn(rows, n, device=dev) * 2 + 0.5).to(dt)
        g = torch.randn(n, device=dev)
        b = torch.randn(n, device=dev)
        out = torch.zeros(rows, n, device=dev, dtype=torch.bfloat16)
        ops.layernorm(x, g, b, out)
        ref = torch.nn.functional.layer_norm(x.float(), (n,), g, b, 1e-5)
        ok &= report(f"layernorm rows={rows} n={n} {dt}", out, ref, 4e-2)
    # pre_add + scatter
    rows, n = 514, 1024
    x = torch.randn(rows, n, device=dev)
    pa = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev); b = torch.randn(n, device=dev)
    out = torch.zeros(2 * 321, n, device=dev, dtype=torch.bfloat16)
    ops.layernorm(x, g, b, out, pre_add=pa, grp=(257, 321, 0))
    ref = torch.zeros(2 * 321, n, device=dev)
    r = torch.nn.functional.layer_norm(x + pa, (n,), g, b, 1e-5)
    ref[0:257] = r[0:257]; ref[321:321 + 257] = r[257:]
    ok &= report("layernorm pre_add+scatter", out, ref, 4e-2)
    return ok


def case_ln_fold():
    """LayerNorm folded into the consumer GEMM: producers emit a bf16 copy + per-row partial
    (sum, sumsq); the consumer applies rstd*(acc - mean*c) + d.  Checked against torch layer_norm + matmul."""
    import torch.nn.functional as F
    torch.manual_seed(7)
    ok = True
    for (M, K, N, N2) in ((1000, 512, 768, 320), (70, 128, 128, 256), (515, 2048, 296, 1002), (16384, 2048, 2048, 512)):
        x = torch.randn(M, K, device=dev) * 2 + 0.3
        gamma, beta = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev) * 0.1
        W = torch.randn(N, K, device=dev) / math.sqrt(K)
        b = torch.randn(N, device=dev) * 0.1
        xb = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(1, M, 2, device=dev)
        ops.rowstats_cast(x, xb, st)
        ok &= bool(torch.equal(xb, x.bfloat16()))
        xf = xb.float()
        ok &= report(f"rowstats sum M={M} K={K}", st[0, :, 0], xf.sum(1), 1e-3 * math.sqrt(K))
        ok &= report(f"rowstats sumsq M={M} K={K}", st[0, :, 1], (xf * xf).sum(1), 2e-5 * K * 5)
        wf = (W * gamma).bfloat16()
        c = wf.double().sum(1).float()
        d = (W.double() @ beta.double() + b.double()).float()
        # consumer 1: fp32 out + residual in place + bf16 copy + stats for the next fold
        res0 = torch.randn(M, N, device=dev)
        y = res0.clone()
        yb = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        st2 = torch.zeros((N + 127) // 128, M, 2, device=dev)
        ops.gemm(xb, wf, y, bias=d, res=y, ln=(st, c, K, 1e-5), stats_out=st2, out2=yb)
        mu, var = xf.mean(1, keepdim=True), xf.var(1, unbiased=False, keepdim=True)
        ref = ((xf - mu) * torch.rsqrt(var + 1e-5)) @ wf.float().T + d + res0
        ok &= report(f"fold gemm (same rounded W') M={M} K={K} N={N}", y, ref, 3e-3)
        ref32 = F.layer_norm(xf, (K,), gamma, beta, 1e-5) @ W.T + b + res0
        ok &= report(f"fold gemm vs layer_norm+linear fp32", y, ref32, 4e-2)
        good = bool(torch.equal(yb, y.bfloat16()))
        print(f"[{'OK' if good else 'FAIL'}] bf16 copy equals bf16(out)")
        ok &= good
        ybf = yb.float()
        ok &= report("producer stats sum", st2[:, :, 0].sum(0), ybf.sum(1), 1e-3 * math.sqrt(N) + 1e-3)
        ok &= report("producer stats sumsq", st2[:, :, 1].sum(0), (ybf * ybf).sum(1), 1e-4 * N)
        # consumer 2: chained fold over multi-tile partials, bf16 out + GELU + stats
        g2, b2 = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev) * 0.1
        W2 = torch.randn(N2, N, device=dev) / math.sqrt(N)
        wf2 = (W2 * g2).bfloat16()
        c2 = wf2.double().sum(1).float()
        d2 = (W2.double() @ b2.double()).float()
        z = torch.zeros(M, N2, device=dev, dtype=torch.bfloat16) if (N2 * 2) % 16 == 0 else torch.zeros(M, N2, device=dev)
        st3 = torch.zeros((N2 + 127) // 128, M, 2, device=dev) if z.dtype == torch.bfloat16 else None
        ops.gemm(yb, wf2, z, bias=d2, act=_abi.KX_ACT_GELU, ln=(st2, c2, N, 1e-5), stats_out=st3)
        mu2, var2 = ybf.mean(1, keepdim=True), ybf.var(1, unbiased=False, keepdim=True)
        ref2 = F.gelu(((ybf - mu2) * torch.rsqrt(var2 + 1e-5)) @ wf2.float().T + d2)
        ok &= report(f"chained fold + gelu N2={N2} ({z.dtype})", z, ref2, 2e-2 if z.dtype == torch.bfloat16 else 3e-3)
        if st3 is not None:
            zf = z.float()
            ok &= report("bf16 producer stats sum", st3[:, :, 0].sum(0), zf.sum(1), 1e-3 * math.sqrt(N2) + 1e-3)
            ok &= report("bf16 producer stats sumsq", st3[:, :, 1].sum(0), (zf * zf).sum(1), 1e-4 * N2)
    # 128-wide tiles: one partial per 64 columns
    M3, K3, N3 = 2056, 256, 1024
    a3 = torch.randn(M3, K3, device=dev).bfloat16()
    w3 = (torch.randn(N3, K3, device=dev) / math.sqrt(K3)).bfloat16()
    y3 = torch.randn(M3, N3, device=dev)
    r3 = y3.clone()
    yb3 = torch.zeros(M3, N3, device=dev, dtype=torch.bfloat16)
    st4 = torch.zeros(N3 // 64, M3, 2, device=dev)
    ops.gemm(a3, w3, y3, res=y3, stats_out=st4, out2=yb3)
    ok &= report("bn=128 producer out", y3, a3.float() @ w3.float().T + r3, 1e-2)
    ok &= report("bn=128 producer stats sum", st4[:, :, 0].sum(0), yb3.float().sum(1), 5e-2)
    ok &= report("bn=128 producer stats sumsq", st4[:, :, 1].sum(0), (yb3.float() ** 2).sum(1), 2e-1)
    # attention: per-head partial statistics of the stored rows
    B, H, T = 2, 4, 300
    qkv = torch.randn(B * T, 3 * H * 64, device=dev).bfloat16()
    out = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16)
    sta = torch.zeros(H, B * T, 2, device=dev)
    ops.attention(qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:], out, batch=B, heads=H, seq_len=T,
                  causal=True, scale=0.125, stats_out=sta)
    of = out.float().view(B * T, H, 64)
    ok &= report("attention stats sum", sta[:, :, 0].T, of.sum(-1), 1e-4)
    ok &= report("attention stats sumsq", sta[:, :, 1].T, (of * of).sum(-1), 1e-4)
    return ok


def case_accurate():
    """Verification-precision kernels (csrc/accurate.cu): split-operand GEMM vs an fp64 product, fp32 attention, fp32
    xPos rotation, fp32 patch im2col — each against plain PyTorch in fp64/fp32."""
    torch.manual_seed(11)
    ok = True
    for M, N, K in ((114, 96, 128), (300, 1002, 2048), (257, 256, 588)):
        a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * K ** -0.5
        bias = torch.randn(N, device=dev)
        kp = (K + 63) // 64 * 64
        a3 = ops.split_bf16x3(a, n_pad=kp); w3 = ops.split_bf16x3(w, weights=True, n_pad=kp)
        hi = a.bfloat16().float()
        ok &= bool(torch.equal(a3[:, :K].float(), hi) and torch.equal(a3[:, kp:kp + K], a3[:, :K])
                   and torch.equal(a3[:, 2 * kp:2 * kp + K].float(), (a - hi).bfloat16().float()))
        ok &= bool(torch.equal(w3[:, kp:kp + K].float(), (w - w.bfloat16().float()).bfloat16().float()) and (a3[:, K:kp] == 0).all())
        out = torch.empty(M, N, device=dev)
        ops.gemm(a3, w3, out, bias=bias)
        want = (a.double() @ w.double().T + bias.double()).float()
        ok &= report(f"bf16x3 GEMM {M}x{N}x{K} vs fp64", out, want, 5e-5 * max(1.0, (K / 128) ** 0.5))
        if K % 8 == 0:
            plain = torch.empty(M, N, device=dev)
            ops.gemm(a.bfloat16(), w.bfloat16(), plain, bias=bias)
            print(f"      (plain bf16 operands: max err {(plain - want).abs().max().item():.2e})")
    for B, H, nq, nkv, causal in ((2, 3, 114, 114, True), (3, 2, 64, 321, False), (1, 2, 257, 257, False), (2, 2, 33, 33, True)):
        q = torch.randn(B * nq, H * 64 + 64, device=dev); kv = torch.randn(B * nkv, 2 * H * 64, device=dev)
        out = torch.zeros(B * nq, H * 64, device=dev)
        ops.attention_f32(q[:, :H * 64], kv[:, :H * 64], kv[:, H * 64:], out, batch=B, heads=H, n_q=nq, n_kv=nkv, causal=causal, scale=0.125)
        qh = q[:, :H * 64].view(B, nq, H, 64).transpose(1, 2).double()
        kh = kv[:, :H * 64].view(B, nkv, H, 64).transpose(1, 2).double()
        vh = kv[:, H * 64:].view(B, nkv, H, 64).transpose(1, 2).double()
        sc = qh @ kh.transpose(-1, -2) * 0.125
        if causal:
            sc = sc + torch.triu(torch.full((nq, nkv), float("-inf"), device=dev, dtype=torch.float64), 1)
        want = (torch.softmax(sc, -1) @ vh).transpose(1, 2).reshape(B * nq, H * 64).float()
        ok &= report(f"attn_f32 B={B} H={H} {nq}x{nkv} causal={causal}", out, want, 2e-6)
    T, D = 50, 128
    tabs = ops.xpos_tables((torch.arange(0, 64, 2, device=dev) + 25.6) / 89.6, 1.0 / (10000 ** (torch.arange(0, 32, device=dev) / 32)),
                           T, (-T) // 2, 512.0, dev)
    qkv = torch.randn(2 * T, 3 * D, device=dev)
    want = qkv.clone()
    for blk, (c, s_) in ((0, (tabs[0], tabs[1])), (1, (tabs[2], tabs[3]))):
        z = qkv[:, blk * D:(blk + 1) * D].view(2, T, D // 64, 32, 2)
        cc, ss = c.view(1, T, 1, 32), s_.view(1, T, 1, 32)
        want[:, blk * D:(blk + 1) * D] = torch.stack([z[..., 0] * cc - z[..., 1] * ss, z[..., 1] * cc + z[..., 0] * ss], -1).view(2 * T, D)
    ops.xpos_apply_f32(qkv, D, T, tabs)
    ok &= report("xpos_apply_f32", qkv, want, 1e-6)
    n, media, image, patch, Dv = 4, 2, 56, 14, 128
    px = torch.randn(n, 3, image, image, device=dev)
    P = (image // patch) ** 2
    kp = 640
    patches = torch.full((n * P, kp), 7.0, device=dev)
    cls = torch.randn(Dv, device=dev); vpos = torch.randn(P + 1, Dv, device=dev)
    x = torch.zeros(n, P + 1, Dv, device=dev)
    ops.im2col_patches_f32(px, patches, cls, vpos, x, image=image, patch=patch, media=media)
    un = torch.nn.functional.unfold(px, patch, stride=patch).transpose(1, 2)     # (n, P, 588)
    slot = torch.tensor([(i % media) * (n // media) + i // media for i in range(n)])
    want = torch.zeros(n, P, kp, device=dev)
    want[slot, :, :588] = un
    ok &= bool(torch.equal(patches.view(n, P, kp), want)) and bool(torch.equal(x[:, 0], (cls + vpos[0]).expand(n, -1)))
    print(f"[{'OK' if ok else 'FAIL'}] im2col_patches_f32 (media-major slots, CLS rows)")
    return ok


def case_embed():
    torch.manual_seed(5)
    ok = True
    B, t_text, V, D, n_img = 3, 50, 1002, 256, 64
    T = t_text + n_img
    tok = torch.randint(0, V, (B, t_text), device=dev)
    emb = torch.randn(V, D, device=dev); pos = torch.randn(T + 2, D, device=dev)
    x0 = torch.zeros(B, T, D, device=dev)
    ops.embed_splice_pos(tok, emb, pos, x0, img_rows=(2,), n_img=n_img)
    e = emb[tok]
    ref = torch.cat([e[:, :2], torch.zeros(B, n_img, D, device=dev), e[:, 2:]], 1) + pos[2:T + 2]
    ref[:, 2:2 + n_img] = 0
    ok &= report("embed_splice_pos", x0.view(B * T, D), ref.view(B * T, D), 1e-6)
    # alias_positions: text token i at row t gets (embed + pos[i + 2]) + pos[t + 2]  (torchscale's in-place `x += positions`)
    xa = torch.zeros(B, T, D, device=dev)
    ops.embed_splice_pos(tok, emb, pos, xa, img_rows=(2,), n_img=n_img, alias_positions=True)
    ea = e + pos[2:t_text + 2]
    refa = torch.cat([ea[:, :2], torch.zeros(B, n_img, D, device=dev), ea[:, 2:]], 1) + pos[2:T + 2]
    refa[:, 2:2 + n_img] = 0
    ok &= bool(torch.equal(xa, refa))
    print(f"[{'OK' if torch.equal(xa, refa) else 'FAIL'}] embed_splice_pos with aliased positions: bit-exact")
    # three images: in front of text tokens 0, 7, 7 (adjacent) -> spliced rows 0, 71, 135
    T3 = t_text + 3 * n_img
    pos3 = torch.randn(T3 + 2, D, device=dev)
    x3 = torch.zeros(B, T3, D, device=dev)
    ops.embed_splice_pos(tok, emb, pos3, x3, img_rows=(0, 7 + n_img, 7 + 2 * n_img), n_img=n_img)
    z = torch.zeros(B, n_img, D, device=dev)
    ref3 = torch.cat([z, e[:, :7], z, z, e[:, 7:]], 1) + pos3[2:T3 + 2]
    ref3[:, :n_img] = 0
    ref3[:, 7 + n_img:7 + 3 * n_img] = 0
    ok &= report("embed_splice_pos 3 images", x3.view(B * T3, D), ref3.view(B * T3, D), 1e-6)
    # im2col
    Bi, image, patch, dim = 2, 56, 14, 128
    px = torch.randn(Bi, 3, image, image, device=dev)
    g = image // patch
    kp = 640
    patches = torch.full((Bi * g * g, kp), 7.0, device=dev, dtype=torch.bfloat16)
    cls = torch.randn(dim, device=dev); p2 = torch.randn(g * g + 1, dim, device=dev)
    x = torch.zeros(Bi, g * g + 1, dim, device=dev)
    ops.im2col_patches(px, patches, cls, p2, x, image=image, patch=patch)
    ref = torch.nn.functional.unfold(px, patch, stride=patch).transpose(1, 2).reshape(Bi * g * g, 3 * patch * patch)
    ok &= report("im2col", patches[:, :588], ref.bfloat16(), 0.0)
    ok &= report("im2col pad", patches[:, 588:], torch.zeros(Bi * g * g, kp - 588, device=dev), 0.0)
    ok &= report("cls rows", x[:, 0], (cls + p2[0]).expand(Bi, dim), 1e-6)
    # media-major slots: pixels (2 sequences, 3 images) -> slot i*2 + s
    px = torch.randn(6, 3, image, image, device=dev)
    patches = torch.zeros(6 * g * g, kp, device=dev, dtype=torch.bfloat16)
    x = torch.zeros(6, g * g + 1, dim, device=dev)
    ops.im2col_patches(px, patches, cls, p2, x, image=image, patch=patch, media=3)
    pm = px.view(2, 3, 3, image, image).transpose(0, 1).reshape(6, 3, image, image)
    ref = torch.nn.functional.unfold(pm, patch, stride=patch).transpose(1, 2).reshape(6 * g * g, 3 * patch * patch)
    ok &= report("im2col media-major", patches[:, :588], ref.bfloat16(), 0.0)
    # cast / broadcast
    s = torch.randn(1000003, device=dev)
    ok &= report("cast bf16", ops.cast_bf16(s).view(1, -1), s.bfloat16().view(1, -1), 0.0)
    lat = torch.randn(64, 128, device=dev); dst = torch.zeros(3, 64, 128, device=dev)
    ops.broadcast_rows(lat, dst, 3)
    ok &= report("broadcast", dst.view(3, -1), lat.expand(3, 64, 128).reshape(3, -1), 0.0)
    return ok


def case_perceiver_attn():
    """Cross-attention through the tensor-core flash kernel (kv_len != seq_len): the model's shape, ragged query / key counts,
    a query count above one 128-row tile and above one tile pair; then the time of the model's launch (B = 8)."""
    torch.manual_seed(6)
    ok = True
    for B, H, nq, nkv in ((2, 8, 64, 321), (3, 2, 33, 130), (1, 4, 200, 77), (2, 3, 300, 321), (8, 8, 64, 321)):
        q = torch.randn(B * nq, H * 64, device=dev).bfloat16()
        kv = torch.randn(B * nkv, 2 * H * 64, device=dev).bfloat16()
        out = torch.zeros(B * nq, H * 64, device=dev, dtype=torch.bfloat16)
        ops.perceiver_attention(q, kv, out, batch=B, heads=H, n_q=nq, n_kv=nkv, v_col_off=H * 64, scale=0.125)
        qh = q.float().view(B, nq, H, 64).transpose(1, 2)
        kh = kv[:, :H * 64].float().view(B, nkv, H, 64).transpose(1, 2)
        vh = kv[:, H * 64:].float().view(B, nkv, H, 64).transpose(1, 2)
        ref = ((qh @ kh.transpose(-1, -2)) * 0.125).softmax(-1) @ vh
        ok &= report(f"perceiver xattn B={B} H={H} {nq}x{nkv}", out, ref.transpose(1, 2).reshape(B * nq, H * 64), 2e-2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.perceiver_attention(q, kv, out, batch=B, heads=H, n_q=nq, n_kv=nkv, v_col_off=H * 64, scale=0.125)
    e0.record()
    for _ in range(50):
        ops.perceiver_attention(q, kv, out, batch=B, heads=H, n_q=nq, n_kv=nkv, v_col_off=H * 64, scale=0.125)
    e1.record()
    torch.cuda.synchronize()
    print(f"  perceiver xattn B=8 H=8 64x321: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch (back to back)")
    return ok


def bench_gemm(M, N, K, cg, bn, iters=20):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, out, cta_group=cg, block_n=bn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, out, cta_group=cg, block_n=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.T, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"gemm M={M} N={N} K={K} cg={cg} bn={bn}: {ms*1e3:.1f} us  {tf:.0f} TFLOP/s   (cuBLAS {2.0*M*N*K/ms2/1e9:.0f})")


def case_bench():
    for cg, bn in ((1, 256), (2, 256), (2, 128)):
        bench_gemm(16384, 6144, 2048, cg, bn)
        bench_gemm(16384, 2048, 2048, cg, bn)
        bench_gemm(16384, 8192, 2048, cg, bn)
        bench_gemm(16384, 2048, 8192, cg, bn)
    return case_bench_attn()


def case_bench_vit():
    """Tile-shape sweep for the ViT GEMMs (M = 8 x 257 rows, K = 1024 / 4096) and the out_proj shape."""
    for M, N, K in ((2056, 3072, 1024), (2056, 4096, 1024), (2056, 1024, 1024), (2056, 1024, 4096), (16384, 2048, 2048)):
        for cg, bn in ((1, 128), (1, 256), (2, 128), (2, 256)):
            bench_gemm(M, N, K, cg, bn, iters=50)
    return True


def case_bench_attn():
    for B, H, T, causal in ((8, 32, 2048, True), (8, 32, 1024, True), (8, 32, 4096, True), (8, 32, 256, True),
                            (8, 16, 257, False), (1, 32, 114, True)):
        bench_attn(B, H, T, causal)
    return True


def bench_attn(B, H, T, causal):
    qkv = torch.randn(B * T, 3 * H * 64, device=dev).bfloat16()
    q, k, v = qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:]
    out = torch.empty(B * T, H * 64, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(q, k, v, out, batch=B, heads=H, seq_len=T, causal=causal, scale=0.125)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention(q, k, v, out, batch=B, heads=H, seq_len=T, causal=causal, scale=0.125)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * H * T * T * 64 * 2 / (2 if causal else 1)
    print(f"attn causal={causal} B={B} H={H} T={T}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s "
          f"({'causal-halved' if causal else 'full'})")
    return True


def _time(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case_bench_train():
    """Stand-alone timings of the training-step kernels at configs[3] shapes (M = 8 x 2048 rows)."""
    B, H, T = 8, 32, 2048
    D, F, M = H * 64, 8192, B * T
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    out = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(H, B, ops.lse_pad(T), device=dev)
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, batch=B, heads=H, seq_len=T, causal=True, scale=0.125, lse_out=lse)
    d_out = torch.randn(M, D, device=dev).bfloat16()
    dqkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    acc = torch.empty(M, D, device=dev); delta = torch.empty(*lse.shape, 2, device=dev)
    ms = _time(lambda: ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, d_out, lse, dqkv[:, :D], dqkv[:, D:2 * D],
                                         dqkv[:, 2 * D:], acc, delta, batch=B, heads=H, seq_len=T, causal=True, scale=0.125))
    fl = 10.0 * B * H * T * T * 64 / 2
    print(f"attn_bwd B={B} H={H} T={T}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s (causal-halved, 5 MMAs)")
    # the same pair with attention dropout (keep bits drawn ahead by the mask kernel)
    words = ops.attn_dropout_mask_words(B, H, T)
    rows = torch.zeros(words, dtype=torch.int32, device=dev); keys = torch.zeros(words, dtype=torch.int32, device=dev)
    ms = _time(lambda: ops.attn_dropout_masks(rows, keys, p=0.1, site=6, seed=99, batch=B, heads=H, seq_len=T))
    print(f"attn_dropout_masks B={B} H={H} T={T}: {ms*1e3:.1f} us")
    ms = _time(lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, batch=B, heads=H, seq_len=T, causal=True, scale=0.125,
                                     lse_out=lse, drop_p=0.1, row_mask=rows))
    print(f"attn fwd + lse + dropout: {ms*1e3:.1f} us  {4.0 * B * H * T * T * 64 / 2 / ms / 1e9:.0f} TFLOP/s (causal-halved)")
    ms = _time(lambda: ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, batch=B, heads=H, seq_len=T, causal=True, scale=0.125,
                                     lse_out=lse))
    print(f"attn fwd + lse (no dropout): {ms*1e3:.1f} us")
    ms = _time(lambda: ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, d_out, lse, dqkv[:, :D], dqkv[:, D:2 * D],
                                         dqkv[:, 2 * D:], acc, delta, batch=B, heads=H, seq_len=T, causal=True, scale=0.125,
                                         drop_p=0.1, drop_mask=keys))
    print(f"attn_bwd + dropout: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s")
    # LayerNorm backward, residual form (n = 2048) and FFN form (n = 8192, GELU)
    x = torch.randn(M, D, device=dev); dy = torch.randn(M, D, device=dev).bfloat16()
    gamma = torch.ones(D, device=dev); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); dc = torch.zeros(D, device=dev)
    dx = torch.randn(M, D, device=dev); dxb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    part = torch.empty(3, ops.ln_bwd_partials(M), D, device=dev)
    ms = _time(lambda: ops.layernorm_bwd(x, dy, gamma, dx, dg, db, part, dres=dx, dxb=dxb, d_colsum=dc))
    by = M * D * (4 + 2 + 4 + 4 + 2)
    print(f"layernorm_bwd fp32 residual form {M}x{D}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s")
    u = torch.randn(M, F, device=dev).bfloat16(); dgl = torch.randn(M, F, device=dev).bfloat16()
    gf = torch.ones(F, device=dev); dgf = torch.zeros(F, device=dev); dbf = torch.zeros(F, device=dev); dcf = torch.zeros(F, device=dev)
    du = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    partf = torch.empty(3, ops.ln_bwd_partials(M), F, device=dev)
    ms = _time(lambda: ops.layernorm_bwd(u, dgl, gf, du, dgf, dbf, partf, act=_abi.KX_ACT_GELU, d_colsum=dcf))
    print(f"layernorm_bwd GELU form {M}x{F}: {ms*1e3:.1f} us  {M*F*6/ms/1e6:.0f} GB/s")
    gl = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    ms = _time(lambda: ops.act_layernorm(u, gf, dbf, gl))
    print(f"act_layernorm {M}x{F}: {ms*1e3:.1f} us  {M*F*4/ms/1e6:.0f} GB/s")
    h = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    ms = _time(lambda: ops.layernorm(x, gamma, db, h))
    print(f"layernorm fwd {M}x{D}: {ms*1e3:.1f} us  {M*D*6/ms/1e6:.0f} GB/s")
    ms = _time(lambda: ops.colsum(dqkv, torch.zeros(3 * D, device=dev)))
    print(f"colsum {M}x{3*D}: {ms*1e3:.1f} us")
    # backward GEMMs
    w1 = torch.randn(F, D, device=dev).bfloat16(); h2 = torch.randn(M, D, device=dev).bfloat16()
    dh = torch.empty(M, D, device=dev, dtype=torch.bfloat16); dw = torch.empty(F, D, device=dev)
    ms = _time(lambda: ops.gemm(du, w1, dh, b_trans=True))
    print(f"dgrad {M}x{D}x{F}: {ms*1e3:.1f} us  {2.0*M*D*F/ms/1e9:.0f} TFLOP/s")
    ms = _time(lambda: ops.gemm(du, h2, dw, a_trans=True, b_trans=True))
    print(f"wgrad {F}x{D}x{M}: {ms*1e3:.1f} us  {2.0*M*D*F/ms/1e9:.0f} TFLOP/s")
    wo = torch.randn(D, D, device=dev).bfloat16(); dwo = torch.empty(D, D, device=dev)
    ms = _time(lambda: ops.gemm(dxb, h2, dwo, a_trans=True, b_trans=True))
    print(f"wgrad {D}x{D}x{M}: {ms*1e3:.1f} us  {2.0*M*D*D/ms/1e9:.0f} TFLOP/s")
    return True


def case_decode():
    """Incremental-decoding kernels against plain fp32 PyTorch on the same bf16 inputs."""
    ok = True
    F = torch.nn.functional
    g = torch.Generator(device=dev).manual_seed(0)
    # --- decode_linear: plain / LayerNorm fold / GELU / ragged N / every batch-group width
    for (B, N, K, ln, act, f32o) in ((8, 2048, 2048, True, 0, False), (3, 1002, 128, True, 1, True), (16, 512, 8192, True, 1, False),
                                     (32, 320, 256, False, 0, True), (1, 32002, 2048, True, 0, True), (8, 8192, 2048, True, 1, False)):
        a = (torch.randn(B + 2, K, device=dev, generator=g) * 1.5 + 0.3).bfloat16()[1:B + 1]        # offset rows
        w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=dev, generator=g)
        c = w.double().sum(1).float()
        out = torch.full((B, N + 8), float("nan"), device=dev, dtype=torch.float32 if f32o else torch.bfloat16)[:, :N]
        ops.decode_linear(a, w, bias=bias, ln_c=c if ln else None, act=_abi.KX_ACT_GELU if act else _abi.KX_ACT_NONE, out=out)
        af = a.float()
        h = F.layer_norm(af, (K,), eps=1e-5) if ln else af
        ref = h @ w.float().T + bias
        if act:
            ref = F.gelu(ref)
        tol = 3e-3 * max(1.0, ref.abs().max().item()) if f32o else 1.2e-2 * max(1.0, ref.abs().max().item())
        ok &= report(f"decode_linear B={B} N={N} K={K} ln={ln} act={act} f32={f32o}", out, ref, tol)
    # --- residual mode
    B, N, K = 5, 256, 512
    a = torch.randn(B, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    x = torch.randn(B, N, device=dev, generator=g)
    x0 = x.clone()
    xb = torch.empty(B, N, device=dev, dtype=torch.bfloat16)
    ops.decode_linear(a, w, bias=bias, ln_c=w.double().sum(1).float(), res=(x, xb))
    ref = x0 + F.layer_norm(a.float(), (K,)) @ w.float().T + bias
    ok &= report("decode_linear residual x", x, ref, 3e-3)
    ok &= report("decode_linear residual xb", xb, ref, 2e-2)
    # --- q|k|v mode: xPos at *pos, cache write
    B, D, T_MAX, P = 4, 128, 40, 17
    xp, scale, sin, cos = xpos_ref(T_MAX, dev)
    inv_freq = (1.0 / (10000 ** (torch.arange(0, 32) / 32))).to(dev)
    tabs = ops.xpos_tables(xp.scale.to(dev), inv_freq, T_MAX, (-T_MAX) // 2, 512.0, dev)
    a = torch.randn(B, D, device=dev, generator=g).bfloat16()
    w = (torch.randn(3 * D, D, device=dev, generator=g) / math.sqrt(D)).bfloat16()
    bias = torch.randn(3 * D, device=dev, generator=g)
    qo = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
    kc = torch.zeros(B, D // 64, T_MAX, 64, device=dev, dtype=torch.bfloat16)      # head-major cache
    vc = torch.zeros(B, D // 64, T_MAX, 64, device=dev, dtype=torch.bfloat16)
    pos = torch.tensor([P], device=dev, dtype=torch.int32)
    ops.decode_linear(a, w, bias=bias, ln_c=w.double().sum(1).float(), qkv=(qo, kc, vc, T_MAX, pos, tabs))
    y = F.layer_norm(a.float(), (D,)) @ w.float().T + bias
    q, k, v = y[:, :D], y[:, D:2 * D], y[:, 2 * D:]

    def rot(t, up):
        th = t.view(B, D // 64, 32, 2)
        s_ = scale[P] if up else 1.0 / scale[P]
        c_, n_ = (cos[P] * s_)[None, None], (sin[P] * s_)[None, None]
        o0 = th[..., 0] * c_ - th[..., 1] * n_
        o1 = th[..., 1] * c_ + th[..., 0] * n_
        return torch.stack([o0, o1], -1).view(B, D)

    ok &= report("decode qkv: q rotated", qo, rot(q, True), 3e-2)
    ok &= report("decode qkv: k rotated -> cache row", kc[:, :, P].reshape(B, D), rot(k, False), 3e-2)
    ok &= report("decode qkv: v -> cache row", vc[:, :, P].reshape(B, D), v, 3e-2)
    kc[:, :, P] = 0; vc[:, :, P] = 0
    ok &= bool((kc == 0).all() and (vc == 0).all())          # nothing else was touched
    # --- decode attention against eager softmax, several cache fills (1 chunk, ragged, many chunks)
    for (B, H, T_MAX, n_keys) in ((2, 4, 64, 1), (3, 2, 300, 131), (2, 32, 2048, 2048), (8, 32, 640, 517)):
        D = H * 64
        q = torch.randn(B, D, device=dev, generator=g).bfloat16()
        kc = torch.randn(B, H, T_MAX, 64, device=dev, generator=g).bfloat16()
        vc = torch.randn(B, H, T_MAX, 64, device=dev, generator=g).bfloat16()
        out = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
        scratch, counters = ops.decode_attn_scratch(B, H, T_MAX, dev)
        pos = torch.tensor([n_keys - 1], device=dev, dtype=torch.int32)
        for _ in range(2):                                   # twice: the counters must come back to zero
            ops.decode_attention(q, kc, vc, out, t_max=T_MAX, heads=H, pos=pos, scale=0.125, scratch=scratch, counters=counters)
        qh = q.float().view(B, H, 1, 64)
        kh = kc[:, :, :n_keys].float()
        vh = vc[:, :, :n_keys].float()
        ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).view(B, D)
        ok &= report(f"decode_attn B={B} H={H} keys={n_keys}/{T_MAX}", out, ref, 1.5e-2)
        ok &= bool((counters == 0).all())
    # --- cache fill from a q|k|v matrix, embed, greedy choice
    B, T, D, T_MAX = 3, 21, 128, 30
    qkv = torch.randn(B * T, 3 * D, device=dev, generator=g).bfloat16()
    kc = torch.zeros(B, D // 64, T_MAX, 64, device=dev, dtype=torch.bfloat16); vc = torch.zeros_like(kc)
    ops.kv_cache_store(qkv, kc, vc, batch=B, seq_len=T, d_model=D, t_max=T_MAX)
    ok &= bool(torch.equal(kc[:, :, :T], qkv[:, D:2 * D].view(B, T, D // 64, 64).transpose(1, 2))
               and torch.equal(vc[:, :, :T], qkv[:, 2 * D:].view(B, T, D // 64, 64).transpose(1, 2)))
    ok &= bool((kc[:, :, T:] == 0).all())
    V = 1002
    emb = torch.randn(V, D, device=dev, generator=g); ptab = torch.randn(64, D, device=dev, generator=g)
    tok = torch.tensor([5, 1001, 77], device=dev)
    pos = torch.tensor([9], device=dev, dtype=torch.int32)
    x = torch.empty(B, D, device=dev); xb = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    ops.decode_embed(tok, emb, ptab, pos, x, xb, err)
    ok &= bool(torch.equal(x, emb[tok] + ptab[11]) and torch.equal(xb, x.bfloat16()) and int(err.item()) == 0)
    ops.decode_embed(tok, emb, ptab, pos, x, xb, err, text_index_off=4)       # + the text-index position 11 - 4
    ok &= bool(torch.equal(x, (emb[tok] + ptab[7]) + ptab[11]) and int(err.item()) == 0)
    ops.decode_embed(torch.tensor([5, 1002, 77], device=dev), emb, ptab, pos, x, xb, err)
    ok &= int(err.item()) == 1
    logits = torch.randn(B, V, device=dev, generator=g)
    logits[1, 700] = logits[1, 300] = 50.0                   # tie -> lowest index
    hist = torch.full((B, 4), -1, device=dev, dtype=torch.int64)
    step = torch.zeros(1, device=dev, dtype=torch.int32); counter = torch.zeros(1, device=dev, dtype=torch.int32)
    tok_out = torch.zeros(B, device=dev, dtype=torch.int64)
    ops.argmax_advance(logits, tok_out, step=step, counter=counter, pos=pos, history=hist)
    want = logits.argmax(-1); want[1] = 300
    ok &= bool(torch.equal(tok_out, want) and torch.equal(hist[:, 0], want) and int(step.item()) == 1 and int(pos.item()) == 10)
    forced = torch.arange(B * 4, device=dev).view(B, 4)
    ops.argmax_advance(logits, tok_out, step=step, counter=counter, pos=None, history=hist, forced=forced)
    ok &= bool(torch.equal(tok_out, forced[:, 1]) and torch.equal(hist[:, 1], forced[:, 1]) and int(pos.item()) == 10
               and int(counter.item()) == 0)
    # greedy choice fused into the LM-head launch: (value, lowest index) keys
    Bk, Nk, Kk = 5, 1002, 128
    a = torch.randn(Bk, Kk, device=dev, generator=g).bfloat16()
    w = (torch.randn(Nk, Kk, device=dev, generator=g) / math.sqrt(Kk)).bfloat16()
    w[900] = w[40]                                           # an exact tie between two vocabulary entries ...
    a[2] = (w[40].float() * 3).bfloat16()                    # ... that is the row maximum for sequence 2
    lg = torch.empty(Bk, Nk, device=dev)
    keys = torch.zeros(Bk, device=dev, dtype=torch.int64)
    ops.decode_linear(a, w, out=lg, argmax_keys=keys)
    tok2 = torch.zeros(Bk, device=dev, dtype=torch.int64)
    ops.argmax_advance(lg, tok2, step=step, counter=counter, pos=pos, keys=keys)
    want2 = torch.stack([(lg[i] == lg[i].max()).nonzero()[0, 0] for i in range(Bk)])
    ok &= bool(torch.equal(tok2, want2) and int(tok2[2]) == 40 and (keys == 0).all() and int(pos.item()) == 11)
    print(f"[{'OK' if ok else 'FAIL'}] cache fill / embed / greedy choice")
    return ok


def _clip_ref(u8, channels_last, mean=ops.CLIP_MEAN, std=ops.CLIP_STD):
    """HF image_transforms.rescale + normalize in plain torch: float64 multiply -> float32, float32 subtract / divide."""
    x = u8.permute(0, 3, 1, 2) if channels_last else u8
    x = (x.double() * (1 / 255)).float()
    m = torch.tensor(mean, dtype=torch.float32, device=u8.device).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32, device=u8.device).view(1, 3, 1, 1)
    return ((x - m) / s).contiguous()


def case_preprocess():
    """CLIP rescale + normalise of raw uint8 pixels: stand-alone and fused into the patch pack, bit-exact."""
    ok = True
    g = torch.Generator(device="cpu").manual_seed(0)
    for image, patch, n, media in ((224, 14, 4, 1), (56, 14, 6, 3), (28, 7, 2, 2)):
        for cl in (False, True):
            u8 = torch.randint(0, 256, (n, image, image, 3) if cl else (n, 3, image, image), dtype=torch.uint8, generator=g).to(dev)
            ref = _clip_ref(u8, cl)
            got = ops.clip_normalize_u8(u8, image=image)
            ok &= report(f"clip_normalize_u8 image={image} channels_last={cl}", got, ref, 0.0)
            gp = image // patch
            P, dim, k_pad = gp * gp, 64, (3 * patch * patch + 63) // 64 * 64
            cls, pos = torch.randn(dim, device=dev), torch.randn(P + 1, dim, device=dev)
            outs = []
            for fused in (False, True):
                patches = torch.full((n * P, k_pad), 3.0, dtype=torch.bfloat16, device=dev)
                x = torch.zeros(n, P + 1, dim, device=dev)
                if fused:
                    ops.im2col_patches_u8(u8, patches, cls, pos, x, image=image, patch=patch, media=media)
                else:
                    ops.im2col_patches(ref, patches, cls, pos, x, image=image, patch=patch, media=media)
                outs.append((patches, x))
            ok &= report(f"im2col_patches_u8 image={image} media={media} channels_last={cl}", outs[1][0], outs[0][0], 0.0)
            pm = ref.view(n // media, media, 3, image, image).transpose(0, 1).reshape(n, 3, image, image)      # media-major slots
            want = torch.nn.functional.unfold(pm, patch, stride=patch).transpose(1, 2).reshape(n * P, 3 * patch * patch)
            ok &= report(f"  vs unfold of the normalised pixels", outs[1][0][:, :3 * patch * patch], want.bfloat16(), 0.0)
            ok &= report(f"  padding columns", outs[1][0][:, 3 * patch * patch:], torch.zeros(n * P, k_pad - 3 * patch * patch, device=dev), 0.0)
            ok &= report(f"  CLS rows", outs[1][1].view(n, -1), outs[0][1].view(n, -1), 0.0)
    return ok


def case_bench_preprocess():
    """Achieved HBM bandwidth of the preprocessing kernels at the model's size (algorithmic bytes: uint8 in + output)."""
    image, patch, dim, k_pad = 224, 14, 1024, 640
    P = (image // patch) ** 2
    for n in (8, 64, 512):
        for cl in (False, True):
            u8 = torch.randint(0, 256, (n, image, image, 3) if cl else (n, 3, image, image), dtype=torch.uint8, device=dev)
            out = torch.empty(n, 3, image, image, device=dev)
            patches = torch.empty(n * P, k_pad, dtype=torch.bfloat16, device=dev)
            x = torch.empty(n, P + 1, dim, device=dev)
            cls, pos = torch.randn(dim, device=dev), torch.randn(P + 1, dim, device=dev)
            t_n = _time(lambda: ops.clip_normalize_u8(u8, out, image=image), iters=20)
            t_f = _time(lambda: ops.im2col_patches_u8(u8, patches, cls, pos, x, image=image, patch=patch), iters=20)
            t_p = _time(lambda: ops.im2col_patches(out, patches, cls, pos, x, image=image, patch=patch), iters=20)
            b_n = u8.numel() * 5
            b_f = u8.numel() + patches.numel() * 2
            b_p = out.numel() * 4 + patches.numel() * 2
            print(f"preprocess N={n} channels_last={cl}: normalize {t_n*1e3:7.1f} us ({b_n/t_n/1e6:6.0f} GB/s)  "
                  f"fused u8 pack {t_f*1e3:7.1f} us ({b_f/t_f/1e6:6.0f} GB/s)  fp32 pack {t_p*1e3:7.1f} us ({b_p/t_p/1e6:6.0f} GB/s)")
    return True


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "list"
    if name == "list":
        print(" ".join(CASES))
        sys.exit(0)
    t0 = time.time()
    _abi.check(_abi.lib.kx_device_check(), "kx_device_check")
    ok = CASES[name]()
    torch.cuda.synchronize()
    print(f"== case {name}: {'PASS' if ok else 'FAIL'} in {time.time()-t0:.1f}s, launches={ops.launch_count()}")
    sys.exit(0 if ok else 1)
