"""Per-stage error table of the sm_100a path against the CPU oracle.

For one input, both precisions of ``kosmosx.Kosmos`` (bf16 = throughput mode, bf16x3 = verification mode) are compared
stage by stage — ViT output, decoder input x0, residual stream after decoder layer 1 / L/2 / L, logits — against the
fp32 oracle and against the oracle emulating the bf16 rounding points of the kernels (incl. the LayerNorm fold).
Usage:  python tools/stage_errors.py [tiny|c1] [--out profiles/r2_stage_errors_c1.md]
c1 = BASELINE.json configs[0]: the README example, 1 x (3,224,224) image + 50 text tokens at the reference's size.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "kosmos-x_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def err(a, b):
    d = a.float().cpu() - b.float().cpu()
    return d.abs().max().item(), d.pow(2).mean().sqrt().item(), b.float().pow(2).mean().sqrt().item()


def main():
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="tiny", choices=["tiny", "c1"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 8)
    if a.which == "tiny":
        oc = ko.OracleConfig.tiny()
        B, t_text = 2, 50
    else:
        oc = ko.OracleConfig(max_positions=2050)
        B, t_text = 1, 50
    ref = ko.build(oc, seed=0)
    kc = KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__})
    mine = Kosmos(config=kc)
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    text, images = ko.make_inputs(oc, B, t_text, seed=1)
    with torch.no_grad():
        st32 = ref.stages(text, images)
        ref.set_emulation(True)
        st16 = ref.stages(text, images)
        ref.set_emulation(False)
    L = oc.layers
    picks = [("vit (ViT-L/14 output)", lambda s: s["vit"]), ("x0 (decoder input)", lambda s: s["x0"]),
             ("after decoder layer 1", lambda s: s["inner_states"][1]),
             (f"after decoder layer {max(L // 2, 1)}", lambda s: s["inner_states"][max(L // 2, 1)]),
             (f"after decoder layer {L}", lambda s: s["inner_states"][L]), ("logits", lambda s: s["logits"])]
    lines = [f"# Per-stage error, {a.which} config (B={B}, T_text={t_text}, T={t_text + oc.p_latents}), "
             f"{torch.cuda.get_device_name(0)}", "",
             "max-abs / RMS of (sm_100a path - oracle); `rms(ref)` = RMS magnitude of the oracle tensor.", "",
             "| stage | rms(ref) | bf16 vs fp32 oracle | bf16 vs bf16-emulating oracle | bf16x3 vs fp32 oracle |", "|---|---|---|---|---|"]
    got16 = mine.stages(text.cuda(), images.cuda(), precision="bf16")
    got3 = mine.stages(text.cuda(), images.cuda(), precision="bf16x3")
    for name, pick in picks:
        e32 = err(pick(got16), pick(st32))
        e16 = err(pick(got16), pick(st16))
        e3 = err(pick(got3), pick(st32))
        lines.append(f"| {name} | {e32[2]:.3f} | {e32[0]:.2e} / {e32[1]:.2e} | {e16[0]:.2e} / {e16[1]:.2e} | {e3[0]:.2e} / {e3[1]:.2e} |")
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt + "\n")


def layer_breakdown():
    """Decoder layer 0 alone, fed the oracle's exact fp32 x0: every intermediate of the 5-launch layer against the
    bf16-emulating oracle computed from the same input (locates where kernels and emulation part ways)."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig, ops, _abi
    oc = ko.OracleConfig.tiny()
    ref = ko.build(oc, seed=0)
    mine = Kosmos(config=KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__}))
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    text, images = ko.make_inputs(oc, 2, 50, seed=1)
    with torch.no_grad():
        x0 = ref.embed_inputs(text, images)
    B, T, D = x0.shape
    H, Fd, M = oc.heads, oc.ffn, B * T
    e = ko._Emu(True, True)
    L = ref.decoder.layers[0]
    sa = L.self_attn
    with torch.no_grad():
        h = ko._ln(e, ko._inner(L.self_attn_layer_norm), x0)
        q = ko._linear(e, h, ko._inner(sa.q_proj)) * sa.scaling
        k = ko._linear(e, h, ko._inner(sa.k_proj))
        v = ko._linear(e, h, ko._inner(sa.v_proj))
        hd = lambda t: t.view(B, T, H, 64).transpose(1, 2).reshape(B * H, T, 64)
        qh, kh, vh = hd(q), hd(k), e.r(hd(v))
        kh = e.r(sa.xpos(kh, offset=0, downscale=True)); qh = e.r(sa.xpos(qh, offset=0, downscale=False))
        w = torch.bmm(qh, kh.transpose(1, 2)) + torch.triu(torch.full((T, T), float("-inf")), 1)[None]
        p = torch.exp(w - w.amax(-1, keepdim=True))
        a = e.r(torch.bmm(e.r(p), vh) / p.sum(-1, keepdim=True))
        a_exact = torch.bmm(p, vh) / p.sum(-1, keepdim=True)
        att = a.view(B, H, T, 64).transpose(1, 2).reshape(B, T, D)
        x_mid = x0 + ko._linear(e, ko._ln(e, ko._inner(sa.inner_attn_ln), att), ko._inner(sa.out_proj))
        ffn = ko._inner(L.ffn)
        u = ko._linear(e, ko._ln(e, ko._inner(L.final_layer_norm), x_mid), ffn.fc1)
        mid = e.r(torch.nn.functional.gelu(u))
        x_out = x_mid + ko._linear(e, ko._ln(e, ffn.ffn_layernorm, mid), ffn.fc2)
        unrot = lambda t: t.view(B, H, T, 64).transpose(1, 2).reshape(M, D)
    dec = mine.decoder
    pk = dec._pack()["layers"][0]
    dev = "cuda"
    bf, f32 = torch.bfloat16, torch.float32
    x = x0.reshape(M, D).to(dev).clone()
    xb = torch.empty(M, D, dtype=bf, device=dev); qkv = torch.empty(M, 3 * D, dtype=bf, device=dev)
    attg = torch.empty(M, D, dtype=bf, device=dev); midg = torch.empty(M, Fd, dtype=bf, device=dev)
    st_in = torch.empty(1, M, 2, device=dev); st_a = torch.empty((D + 127) // 128, M, 2, device=dev)
    st_b = torch.empty_like(st_a); st_att = torch.empty(H, M, 2, device=dev); st_mid = torch.empty((Fd + 127) // 128, M, 2, device=dev)
    tabs = dec._xpos(T, dev)
    ops.rowstats_cast(x, xb, st_in)
    wq, c, d = pk["qkv"]
    ops.gemm(xb, wq, qkv, bias=d, ln=(st_in, c, D, oc.eps), xpos=tuple(tabs), seq_len=T)
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], attg, batch=B, heads=H, seq_len=T, causal=True, scale=0.125, stats_out=st_att)
    wo, c, d = pk["o"]
    ops.gemm(attg, wo, x, bias=d, res=x, ln=(st_att, c, D, oc.eps), stats_out=st_a, out2=xb)
    x_mid_g = x.clone()
    w1, c, d = pk["fc1"]
    ops.gemm(xb, w1, midg, bias=d, act=_abi.KX_ACT_GELU, ln=(st_a, c, D, oc.eps), stats_out=st_mid)
    w2, c, d = pk["fc2"]
    ops.gemm(midg, w2, x, bias=d, res=x, ln=(st_mid, c, Fd, oc.eps), stats_out=st_b, out2=xb)
    rows = [("q (rotated, x 1/8)", qkv[:, :D].float() * 0.125, unrot(qh)), ("k (rotated)", qkv[:, D:2 * D], unrot(kh)),
            ("v", qkv[:, 2 * D:], unrot(vh)), ("attention out", attg, att.reshape(M, D)),
            ("attention out vs un-rounded-P oracle", attg, unrot(a_exact)),
            ("x after out_proj", x_mid_g, x_mid.reshape(M, D)), ("gelu(fc1) (bf16)", midg, mid.reshape(M, Fd)),
            ("x after fc2", x, x_out.reshape(M, D))]
    print("\n# Decoder layer 0 in isolation (input = the oracle's fp32 x0): kernels vs bf16-emulating oracle\n")
    print("| tensor | rms(ref) | max-abs / RMS |\n|---|---|---|")
    for name, g, r in rows:
        er = err(g, r)
        print(f"| {name} | {er[2]:.3f} | {er[0]:.2e} / {er[1]:.2e} |")


if __name__ == "__main__":
    if "--layer" in sys.argv:
        layer_breakdown()
    else:
        main()
