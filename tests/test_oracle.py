"""CPU tests of the oracle: pins against the installed third-party implementations the reference
calls (HF CLIP tower) or that port the same torchscale layer (HF Kosmos-2 text block), the
reference's structural facts, and the committed golden vectors."""
import math
import os

import pytest
import torch

import kosmos_oracle as ko

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tiny_golden.pt")


@pytest.fixture(scope="module")
def tiny():
    torch.set_num_threads(4)
    cfg = ko.OracleConfig.tiny()
    return cfg, ko.build(cfg, seed=0)


def test_vit_matches_hf_clip(tiny):
    """The reference's vision tower IS HF CLIPVisionTransformer (model.py:154-156,230)."""
    transformers = pytest.importorskip("transformers")
    cfg, model = tiny
    hc = transformers.CLIPVisionConfig(hidden_size=cfg.vit_dim, intermediate_size=cfg.vit_mlp,
                                       num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                                       patch_size=cfg.patch, image_size=cfg.image, hidden_act="gelu",
                                       attn_implementation="eager")
    hf = transformers.CLIPVisionModel(hc).vision_model.eval()
    tower = ko.ClipVisionTower(cfg, ko._Emu(False)).eval()
    tower.load_state_dict(hf.state_dict(), strict=True)          # identical parameter names
    _, images = ko.make_inputs(cfg, 2, 5)
    with torch.no_grad():
        a = hf(pixel_values=images)["last_hidden_state"]
        b = tower(images)
    assert a.shape == (2, cfg.vit_tokens, cfg.vit_dim)
    assert (a - b).abs().max() < 1e-5


def test_vit_quick_gelu_matches_hf():
    transformers = pytest.importorskip("transformers")
    cfg = ko.OracleConfig.tiny(vit_act="quick_gelu")
    hc = transformers.CLIPVisionConfig(hidden_size=cfg.vit_dim, intermediate_size=cfg.vit_mlp,
                                       num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                                       patch_size=cfg.patch, image_size=cfg.image, hidden_act="quick_gelu",
                                       attn_implementation="eager")
    hf = transformers.CLIPVisionModel(hc).vision_model.eval()
    tower = ko.ClipVisionTower(cfg, ko._Emu(False)).eval()
    tower.load_state_dict(hf.state_dict(), strict=True)
    _, images = ko.make_inputs(cfg, 1, 5)
    with torch.no_grad():
        assert (hf(pixel_values=images)["last_hidden_state"] - tower(images)).abs().max() < 1e-5


def test_vit_rejects_wrong_image_size(tiny):
    cfg, model = tiny
    with pytest.raises(ValueError):
        model.clip_model(torch.zeros(1, 3, cfg.image + 14, cfg.image))


def test_subln_layer_matches_hf_kosmos2_block(tiny):
    """torchscale's sub-LN DecoderLayer (A.4) vs HF's Kosmos-2 port of it (same parameter names),
    xPos switched off because Kosmos-2 uses sinusoidal absolute positions instead."""
    pytest.importorskip("transformers")
    from transformers.models.kosmos2.configuration_kosmos2 import Kosmos2TextConfig
    from transformers.models.kosmos2.modeling_kosmos2 import Kosmos2TextBlock
    cfg, _ = tiny
    torch.manual_seed(3)
    layer = ko.DecoderLayer(cfg, ko._Emu(False)).eval()
    layer.self_attn.use_xpos = False
    hc = Kosmos2TextConfig(embed_dim=cfg.dim, attention_heads=cfg.heads, ffn_dim=cfg.ffn, layers=1, dropout=0.0,
                           attention_dropout=0.0, activation_function="gelu")
    hc._attn_implementation = "eager"
    blk = Kosmos2TextBlock(hc, layer_idx=0).eval()
    sd = {k.replace(".A.", "."): v for k, v in layer.state_dict().items() if ".B." not in k and "xpos" not in k}
    blk.load_state_dict(sd, strict=True)
    T = 11
    x = torch.randn(2, T, cfg.dim)
    mask = torch.triu(torch.full((T, T), float("-inf")), 1)
    with torch.no_grad():
        a = layer(x, mask)
        b = blk(x, attention_mask=mask[None, None].expand(2, 1, T, T))
        b = b[0] if isinstance(b, tuple) else b
    assert (a - b).abs().max() < 1e-5


def test_xpos_is_relative():
    """q_t . k_u after xPos depends on t-u only."""
    xp = ko.XPOS(64)
    torch.manual_seed(0)
    q = torch.randn(1, 1, 64).expand(1, 33, 64)
    k = torch.randn(1, 1, 64).expand(1, 33, 64)
    s = xp(q, downscale=False)[0] @ xp(k, downscale=True)[0].T
    for d in (0, 1, 5, 17):
        diag = torch.diagonal(s, -d)
        assert (diag - diag[0]).abs().max() < 1e-4 * max(1.0, diag.abs().max().item())


def test_xpos_min_pos_uses_floor_division():
    xp = ko.XPOS(64)
    for T, mp in ((114, -57), (5, -3), (2048, -1024)):
        scale, _, _ = xp.tables(T)
        want = xp.scale ** (torch.tensor(float(mp)) / 512)
        assert torch.allclose(scale[0], want, rtol=1e-6)
        assert scale.shape == (T, 32)


def test_readme_example_shape_and_image_rows(tiny):
    """README.md:29-53 / example.py: (B, T_text) tokens + one image -> (B, T_text+64, vocab); image
    features occupy rows 2..65 (model.py:239-241)."""
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 1, 50)
    with torch.no_grad():
        st = model.stages(text, images)
        st2 = model.stages(text, images * 0.5)
    assert st["logits"].shape == (1, 114, cfg.vocab)
    changed = (st["x0"] - st2["x0"]).abs().amax(-1)[0] > 0
    assert changed[2:66].all() and not changed[:2].any() and not changed[66:].any()


def test_causality(tiny):
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 1, 20)
    text2 = text.clone()
    text2[0, 12:] = (text2[0, 12:] + 7) % cfg.vocab            # spliced rows >= 12+64 change
    with torch.no_grad():
        a, b = model(text, images), model(text2, images)
    assert (a[0, :76] - b[0, :76]).abs().max() < 1e-5
    assert (a[0, 76:] - b[0, 76:]).abs().max() > 1e-3


def test_batch_independence(tiny):
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 3, 12)
    with torch.no_grad():
        full = model(text, images)
        one = model(text[1:2], images[1:2])
    assert (full[1:2] - one).abs().max() < 1e-4


def test_multiway_b_branch_is_inert(tiny):
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 1, 8)
    with torch.no_grad():
        a = model(text, images)
        saved = {n: p.clone() for n, p in model.named_parameters() if ".B." in n}
        assert len(saved) > 0
        for n, p in model.named_parameters():
            if ".B." in n:
                p.add_(1.0)
        b = model(text, images)
        for n, p in model.named_parameters():
            if ".B." in n:
                p.copy_(saved[n])
    assert torch.equal(a, b)


def test_passed_x_skips_embedding(tiny):
    """README.md:179-193: with passed_x the decoder does not embed prev_output_tokens."""
    cfg, model = tiny
    x = torch.randn(1, 7, cfg.dim)
    with torch.no_grad():
        a = model.decoder(torch.zeros(1, 7, dtype=torch.long), passed_x=x)[0]
        b = model.decoder(torch.ones(1, 7, dtype=torch.long) * 5, passed_x=x)[0]
        c = model.decoder(torch.ones(1, 7, dtype=torch.long) * 5)[0]
    assert torch.equal(a, b) and not torch.allclose(a, c)


def test_position_table_limit(tiny):
    cfg, model = tiny
    T = cfg.max_positions - 1                                  # needs rows up to T+1 > table
    with pytest.raises((IndexError, RuntimeError)):
        model.decoder.forward_embedding(torch.zeros(1, T, dtype=torch.long))


def test_sub_ln_init_scale():
    cfg = ko.OracleConfig.tiny()
    torch.manual_seed(0)
    m = ko.KosmosOracle(cfg)
    s = math.sqrt(math.log(2 * cfg.layers))
    fc1 = m.decoder.layers[0].ffn.A.fc1.weight
    q = m.decoder.layers[0].self_attn.q_proj.A.weight
    bound = 1 / math.sqrt(cfg.dim)
    assert fc1.abs().max() > bound * 1.01 and fc1.abs().max() <= bound * s * 1.0001
    assert q.abs().max() <= bound * 1.0001


def test_language_model_oracle():
    cfg = ko.OracleConfig.tiny(vocab=777)
    torch.manual_seed(0)
    m = ko.KosmosLanguageOracle(cfg).eval()
    with torch.no_grad():
        y = m(torch.randint(0, 777, (2, 9)))
    assert y.shape == (2, 9, 777)


def test_incremental_decoding_equals_the_full_forward(tiny):
    """torchscale's incremental_state protocol restated in the oracle: K/V cached un-rotated,
    xPos re-applied each step with offset = src_len - 1 for the query.  Because xPos is relative, step i's logits must
    equal row T0-1+i of one full forward over prompt + generated tokens — the property that pins the incremental path."""
    oc, ref = tiny
    text, images = ko.make_inputs(oc, 2, 12, seed=1)
    n = 6
    toks, lg = ref.generate(text, images, n)
    assert toks.shape == (2, n) and lg.shape == (2, n, oc.vocab)
    with torch.no_grad():
        full = ref(torch.cat([text, toks[:, :-1]], 1), images)
    t0 = text.shape[1] + oc.p_latents
    for i in range(n):
        assert (full[:, t0 - 1 + i] - lg[:, i]).abs().max() < 2e-5
        assert torch.equal(full[:, t0 - 1 + i].argmax(-1), toks[:, i])
    # teacher forcing: the forced tokens are what gets embedded, the logits follow them
    forced = torch.randint(0, oc.vocab, (2, n), generator=torch.Generator().manual_seed(5))
    ftoks, flg = ref.generate(text, images, n, forced=forced)
    assert torch.equal(ftoks, forced)
    with torch.no_grad():
        full = ref(torch.cat([text, forced[:, :-1]], 1), images)
    assert (full[:, t0 - 1:] - flg).abs().max() < 2e-5
    # protocol details: later steps embed only the last token, at position len(prefix) + 1
    dec = ref.decoder
    st = {"is_first_step": False}
    prefix = torch.randint(0, oc.vocab, (1, 9))
    x, _ = dec.forward_embedding(prefix, None, st)
    assert x.shape == (1, 1, oc.dim)
    want = dec.embed_tokens(prefix[:, -1:]) + dec.embed_positions.weight[9 + 1][None, None]
    assert torch.allclose(x, want)


def test_emulation_changes_little_but_something(tiny):
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 1, 10)
    with torch.no_grad():
        a = model(text, images)
        model.set_emulation(True)
        b = model(text, images)
        model.set_emulation(False)
    d = (a - b).abs().max().item()
    assert 0 < d < 0.1 * a.abs().max().item()


def test_oracle_matches_golden(tiny):
    cfg, model = tiny
    g = torch.load(GOLDEN)
    assert g["cfg"] == cfg.__dict__
    for name, c in g["cases"].items():
        text, images = ko.make_inputs(cfg, c["B"], c["t_text"], seed=g["seed_inputs"])
        with torch.no_grad():
            y = model(text, images)
        assert (y[..., ::c["col_step"]] - c["logits"]).abs().max() < 2e-4, name


# --------------------------------------------------------------------------- multi-image splice (configs[4])
GOLDEN_MULTI = os.path.join(os.path.dirname(__file__), "golden", "tiny_golden_multi.pt")


@pytest.fixture(scope="module")
def tiny512():
    torch.set_num_threads(4)
    cfg = ko.OracleConfig.tiny(max_positions=512)
    return cfg, ko.build(cfg, seed=0)


def test_multi_image_default_is_the_reference_splice(tiny):
    """image_positions=None / [2] and a (B,1,3,H,W) image tensor are all the reference call (model.py:239-241)."""
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 2, 12)
    with torch.no_grad():
        a = model(text, images)
        b = model(text, images, image_positions=[2])
        c = model(text, images[:, None], image_positions=[2])
    assert torch.equal(a, b) and torch.allclose(a, c, atol=1e-6)


def test_multi_image_rows_and_media_positions(tiny512):
    """Image i occupies 64 rows in front of text token positions[i]; it is resampled with media_pos_emb[i]
    (flamingo indexes the table by media index); text rows keep their order."""
    cfg, model = tiny512
    pos = [2, 9, 9, 30]
    text, images = ko.make_inputs(cfg, 2, 30, seed=7, n_images=4)
    with torch.no_grad():
        st = model.stages(text, images, pos)
        assert st["logits"].shape == (2, 30 + 4 * 64, cfg.vocab)
        rows = [p + 64 * i for i, p in enumerate(pos)]
        for i in range(4):                                   # perturb one image: exactly its 64 rows of x0 move
            im2 = images.clone()
            im2[:, i] *= 0.5
            moved = (model.stages(text, im2, pos)["x0"] - st["x0"]).abs().amax(-1)[0] > 0
            want = torch.zeros_like(moved)
            want[rows[i]:rows[i] + 64] = True
            assert torch.equal(moved, want), i
        # text rows: embedding + position t+2 in spliced coordinates, and — torchscale's in-place `x += positions` on the
        # aliased `embed` (alias_embed_positions) — the position i+2 of the un-spliced text on top
        emb = model.embed(text)
        is_img = torch.zeros(st["x0"].shape[1], dtype=torch.bool)
        for r in rows:
            is_img[r:r + 64] = True
        T = is_img.numel()
        want_text = emb + model.embed_positions.weight[2:text.shape[1] + 2] + model.embed_positions.weight[2:T + 2][~is_img]
        assert torch.allclose(st["x0"][:, ~is_img], want_text, atol=1e-6)
        model.cfg.alias_embed_positions = False
        x0_plain = model.stages(text, images, pos)["x0"]
        model.cfg.alias_embed_positions = True
        assert torch.allclose(x0_plain[:, ~is_img], emb + model.embed_positions.weight[2:T + 2][~is_img], atol=1e-6)
        assert torch.equal(x0_plain[:, is_img], st["x0"][:, is_img])          # image rows: one positional embedding either way
        # the same picture at media index 0 and 1 gives different rows (media_pos_emb[0] vs [1]) ...
        same = images[:, :1].expand(-1, 2, -1, -1, -1).contiguous()
        r = model._image_rows(same)
        assert (r[:, 0] - r[:, 1]).abs().max() > 1e-3
        # ... and equal rows once the two table rows are made equal
        saved = model.perceive.media_pos_emb.data.clone()
        model.perceive.media_pos_emb.data[1] = saved[0]
        r = model._image_rows(same)
        model.perceive.media_pos_emb.data.copy_(saved)
        assert torch.allclose(r[:, 0], r[:, 1], atol=1e-5)
    with pytest.raises(ValueError):
        model(text, images, image_positions=[2, 9])
    with pytest.raises(ValueError):
        model(text, images, image_positions=[9, 2, 9, 30])


def test_oracle_matches_multi_image_golden(tiny512):
    cfg, model = tiny512
    g = torch.load(GOLDEN_MULTI)
    assert g["cfg"] == cfg.__dict__
    for name, c in g["cases"].items():
        text, images = ko.make_inputs(cfg, c["B"], c["t_text"], seed=g["seed_inputs"], n_images=len(c["positions"]))
        with torch.no_grad():
            y = model(text, images, image_positions=c["positions"])
        assert (y[..., ::c["col_step"]] - c["logits"]).abs().max() < 2e-4, name


# --------------------------------------------------------------------------- pins against independent implementations
def test_xpos_matches_flash_attn_xpos_rotary():
    """The xPos restatement (torchscale XPOS, SURVEY A.5 — torchscale itself is not installable here) against an
    independent implementation that IS installed: flash_attn.layers.rotary.RotaryEmbedding(scale_base=512,
    interleaved=True), the FlashAttention authors' xPos (same ζ = (2i + 0.4d)/(1.4d), base 10000, pair-interleaved
    rotation, q scaled up / k scaled down, positions centred on the sequence).  Even T: the rotated q and k agree
    elementwise.  Odd T: torchscale centres with floor(-T/2), flash_attn with T//2 — a constant shift of every position,
    which cancels in q·kᵀ (the relative-position property), so the score matrices agree."""
    import kosmos_oracle as ko
    rot = pytest.importorskip("flash_attn.layers.rotary")
    hd = 64
    x = ko.XPOS(hd, 512)
    for T in (128, 513):
        r = rot.RotaryEmbedding(hd, interleaved=True, scale_base=512)
        r._update_cos_sin_cache(T, device="cpu", dtype=torch.float32)
        g = torch.Generator().manual_seed(T)
        q, k = torch.randn(3, T, hd, generator=g), torch.randn(3, T, hd, generator=g)
        fq = rot.apply_rotary_emb_torch(q.unsqueeze(2), r._cos_cached, r._sin_cached, interleaved=True).squeeze(2)
        fk = rot.apply_rotary_emb_torch(k.unsqueeze(2), r._cos_k_cached, r._sin_k_cached, interleaved=True).squeeze(2)
        oq, ok = x(q, downscale=False), x(k, downscale=True)
        if T % 2 == 0:
            assert (oq - fq).abs().max().item() <= 2e-5 and (ok - fk).abs().max().item() <= 2e-5
        s_o, s_f = oq @ ok.transpose(1, 2), fq @ fk.transpose(1, 2)
        assert (s_o - s_f).abs().max().item() <= 2e-4 * s_f.abs().max().item()


def test_perceiver_attention_matches_hf_idefics_port():
    """flamingo_pytorch is not installable here; transformers' IdeficsPerceiverAttention is an independent port of the
    same PerceiverAttention (LayerNorm of media and latents, keys / values over [media ‖ latents], q·scale, amax-stabilised
    softmax, bias-free projections).  With the oracle's weights copied in (to_kv split into k / v) the outputs agree."""
    import kosmos_oracle as ko
    from transformers.models.idefics.perceiver import IdeficsPerceiverAttention
    cfg = ko.OracleConfig.tiny()
    torch.manual_seed(0)
    mine = ko._PerceiverAttention(cfg, ko._Emu())
    for ln in (mine.norm_media, mine.norm_latents):
        torch.nn.init.normal_(ln.weight, 1.0, 0.2)
        torch.nn.init.normal_(ln.bias, 0.0, 0.2)
    inner = cfg.p_heads * cfg.p_dim_head
    hf = IdeficsPerceiverAttention(cfg.vit_dim, cfg.p_heads, cfg.p_dim_head, qk_layer_norms=False)
    sd = mine.state_dict()
    hf.load_state_dict({
        "context_layer_norm.weight": sd["norm_media.weight"], "context_layer_norm.bias": sd["norm_media.bias"],
        "latents_layer_norm.weight": sd["norm_latents.weight"], "latents_layer_norm.bias": sd["norm_latents.bias"],
        "q_proj.weight": sd["to_q.weight"], "k_proj.weight": sd["to_kv.weight"][:inner], "v_proj.weight": sd["to_kv.weight"][inner:],
        "output_proj.weight": sd["to_out.weight"]})
    x = torch.randn(2, 1, cfg.vit_tokens, cfg.vit_dim)
    lat = torch.randn(2, 1, cfg.p_latents, cfg.vit_dim)
    with torch.no_grad():
        got = mine(x, lat)
        want = hf(x[:, 0], lat[:, 0])                       # the port has no media-time axis: (B, n, D)
    assert got.shape == (2, 1, cfg.p_latents, cfg.vit_dim) and want.shape == (2, cfg.p_latents, cfg.vit_dim)
    assert (got[:, 0] - want).abs().max().item() <= 2e-5


def test_perceiver_resampler_structure_matches_hf_idefics_port():
    """The whole resampler against transformers' IdeficsPerceiverResampler (same lineage: lucidrains' flamingo-pytorch):
    learned latents broadcast over the batch, depth x [latents += attn(media, latents); latents += ff(latents)], final
    LayerNorm, FF = LN -> Linear -> act -> Linear without biases.  Two things differ by construction and are neutralised:
    the port's FF activation is ReLU (flamingo: GELU) — swapped on the HF instance; the port has no media position
    embedding — the oracle's is zeroed (its broadcast rule is covered by test_multi_image_rows_and_media_positions)."""
    import kosmos_oracle as ko
    from transformers import IdeficsConfig
    from transformers.models.idefics.perceiver import IdeficsPerceiverResampler
    cfg = ko.OracleConfig.tiny()
    torch.manual_seed(0)
    mine = ko.PerceiverResampler(cfg, ko._Emu()).eval()
    with torch.no_grad():
        mine.media_pos_emb.zero_()
        for mod in mine.modules():
            if isinstance(mod, torch.nn.LayerNorm):
                torch.nn.init.normal_(mod.weight, 1.0, 0.2)
                torch.nn.init.normal_(mod.bias, 0.0, 0.2)
    hc = IdeficsConfig()
    hc.vision_config.embed_dim = cfg.vit_dim
    hf = IdeficsPerceiverResampler(hc, cfg.vit_dim, cfg.p_depth, cfg.p_heads, cfg.p_dim_head, cfg.p_latents).eval()
    inner = cfg.p_heads * cfg.p_dim_head
    sd = mine.state_dict()
    new = {"latents": sd["latents"], "layer_norm.weight": sd["norm.weight"], "layer_norm.bias": sd["norm.bias"]}
    for i in range(cfg.p_depth):
        a, f = f"layers.{i}.0.", f"layers.{i}.1."
        new.update({
            f"blocks.{i}.0.context_layer_norm.weight": sd[a + "norm_media.weight"], f"blocks.{i}.0.context_layer_norm.bias": sd[a + "norm_media.bias"],
            f"blocks.{i}.0.latents_layer_norm.weight": sd[a + "norm_latents.weight"], f"blocks.{i}.0.latents_layer_norm.bias": sd[a + "norm_latents.bias"],
            f"blocks.{i}.0.q_proj.weight": sd[a + "to_q.weight"], f"blocks.{i}.0.k_proj.weight": sd[a + "to_kv.weight"][:inner],
            f"blocks.{i}.0.v_proj.weight": sd[a + "to_kv.weight"][inner:], f"blocks.{i}.0.output_proj.weight": sd[a + "to_out.weight"],
            f"blocks.{i}.1.ln.weight": sd[f + "0.weight"], f"blocks.{i}.1.ln.bias": sd[f + "0.bias"],
            f"blocks.{i}.1.fc.weight": sd[f + "1.weight"], f"blocks.{i}.1.c_proj.weight": sd[f + "3.weight"]})
    hf.load_state_dict(new)
    for blk in hf.blocks:
        blk[1].act = torch.nn.GELU()
    x = torch.randn(2, cfg.vit_tokens, cfg.vit_dim)
    with torch.no_grad():
        got, want = mine(x), hf(x)
    assert got.shape == (2, 1, cfg.p_latents, cfg.vit_dim) and want.shape == (2, cfg.p_latents, cfg.vit_dim)
    assert (got[:, 0] - want).abs().max().item() <= 5e-5


def test_decoder_stack_matches_hf_kosmos2_text_model(tiny):
    """The glue of torchscale's ``Decoder`` (token embedding x 1.0, positions starting at padding_idx + 1 = 2, per-layer
    causal mask, layer loop, final ``layer_norm``, bias-free ``output_projection``) against HF's Kosmos-2 text model —
    an independent port of the same torchscale decoder.  xPos is switched off (Kosmos-2 has none) and the oracle's
    learned position table is loaded with the port's sinusoidal rows, so that both see the same inputs."""
    from transformers.models.kosmos2.configuration_kosmos2 import Kosmos2TextConfig
    from transformers.models.kosmos2.modeling_kosmos2 import Kosmos2TextForCausalLM
    cfg, _ = tiny
    torch.manual_seed(4)
    model = ko.KosmosOracle(cfg).eval()
    dec = model.decoder
    for layer in dec.layers:
        layer.self_attn.use_xpos = False
    hc = Kosmos2TextConfig(vocab_size=cfg.vocab, embed_dim=cfg.dim, attention_heads=cfg.heads, ffn_dim=cfg.ffn, layers=cfg.layers,
                           dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, activation_function="gelu", scale_embedding=False,
                           max_position_embeddings=cfg.max_positions, layerdrop=0.0, pad_token_id=1)
    hc._attn_implementation = "eager"
    hf = Kosmos2TextForCausalLM(hc).eval()
    sd = {}
    for k, v in dec.state_dict().items():
        if ".B." in k or "xpos" in k or k.startswith("embed_positions"):
            continue
        k = k.replace(".A.", ".")
        sd[("lm_head." + k[len("output_projection."):]) if k.startswith("output_projection.") else "model." + k] = v
    res = hf.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all("embed_positions" in k for k in res.missing_keys), (res.missing_keys, res.unexpected_keys)
    T = 23
    g = torch.Generator().manual_seed(7)
    tokens = torch.randint(2, cfg.vocab, (2, T), generator=g)                  # no padding (id 1): positions are 2 .. T+1
    pos_mod = hf.model.embed_positions
    ids = pos_mod.create_position_ids_from_input_ids(tokens, padding_idx=1) if hasattr(pos_mod, "create_position_ids_from_input_ids") \
        else None
    if ids is not None:
        assert torch.equal(ids, torch.arange(2, T + 2).expand(2, T))              # the "+2" of PositionalEmbedding.forward
    with torch.no_grad():
        table = pos_mod.weights if hasattr(pos_mod, "weights") else pos_mod.weight
        dec.embed_positions.weight[: cfg.max_positions].copy_(table[: cfg.max_positions].to(torch.float32))
        want = hf(input_ids=tokens).logits
        got, extra = dec(tokens)
    assert got.shape == want.shape == (2, T, cfg.vocab)
    assert (got - want).abs().max().item() <= 2e-4 * max(1.0, want.abs().max().item())
    assert len(extra["inner_states"]) == cfg.layers + 1


# --------------------------------------------------------------------------- round-2 additions
def test_forward_embedding_aliases_embed_like_torchscale(tiny):
    """torchscale: `x = embed = embed_scale * token_embedding; x += positions` — the second result IS the first, so the
    reference's model.py:238 splices embeddings that already carry positions 2..T_text+1 and text rows end up with two
    positional embeddings; image rows with one.  alias_embed_positions=False is the out-of-place reading."""
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 2, 12, seed=3)
    with torch.no_grad():
        x, embed = model.decoder.forward_embedding(text)
        assert x.data_ptr() == embed.data_ptr()
        assert torch.allclose(embed, model.embed(text) + model.embed_positions.weight[2:14], atol=1e-6)
        x0 = model.embed_inputs(text, images)
        T = 12 + cfg.p_latents
        pos = model.embed_positions.weight
        assert torch.allclose(x0[:, :2], model.embed(text[:, :2]) + pos[2:4] + pos[2:4], atol=1e-6)
        assert torch.allclose(x0[:, 66:], model.embed(text[:, 2:]) + pos[4:14] + pos[68:T + 2], atol=1e-6)
        model.cfg.alias_embed_positions = False
        try:
            x, embed = model.decoder.forward_embedding(text)
            assert x.data_ptr() != embed.data_ptr() and torch.equal(embed, model.embed(text))
            x0p = model.embed_inputs(text, images)
            assert torch.allclose(x0p[:, 66:], model.embed(text[:, 2:]) + pos[68:T + 2], atol=1e-6)
            assert torch.equal(x0p[:, 2:66], x0[:, 2:66])
        finally:
            model.cfg.alias_embed_positions = True


def test_incremental_decoding_equals_full_forward_with_aliased_positions(tiny):
    """KosmosOracle.generate continues the TEXT: step i equals row T0-1+i of one forward over the grown text, where the
    new text rows carry both positions (what a user of the reference, which has no generate, would compute)."""
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 2, 9, seed=4)
    forced = torch.randint(0, cfg.vocab, (2, 5), generator=torch.Generator().manual_seed(1))
    _, inc = model.generate(text, images, 5, forced=forced)
    with torch.no_grad():
        full = model(torch.cat([text, forced[:, :-1]], 1), images)
    t0 = 9 + cfg.p_latents
    assert (inc - full[:, t0 - 1:]).abs().max() < 2e-5


def test_loss_targets_reference_rule_is_the_pasted_loop():
    """rule="reference" == experimental/model/allModalities/notes.txt:566-574 applied literally to the reference layout."""
    t_text = 14
    text = torch.arange(100, 100 + t_text)[None].repeat(2, 1)
    tgt = ko.KosmosOracle.loss_targets(text, 64)
    T = t_text + 64
    rows = [0] + list(range(67, T))                       # outputs = cat([outputs[:, :1], outputs[:, 67:]])
    kept = rows[:-1]                                      # outputs[:, :-1]
    labels = torch.cat([text[:, 0:1], text[:, 3:]], 1)    # only_text_tokens (model.py:77): the text without the two markers
    assert torch.equal(tgt[:, kept], labels[:, 1:])       # labels[:, 1:]
    assert int((tgt >= 0).sum()) == 2 * len(kept)
    nt = ko.KosmosOracle.loss_targets(text, 64, rule="next_token")
    assert int(nt[0, 0]) == 101 and int(nt[0, 1]) == -100 and int(nt[0, 66]) == 103 and int(tgt[0, 66]) == -100
    padded = ko.KosmosOracle.loss_targets(text, 64, pad_token_id=105)
    assert int((padded >= 0).sum()) == int((tgt >= 0).sum()) - 2


def test_bf16_emulation_models_the_layernorm_fold(tiny):
    """emulate_bf16 with fold (what the inference kernels do: statistics of the bf16-rounded rows, gamma rounded into the
    weight) and without (LayerNorm materialised, what the training forward does) are two roundings of the same function:
    both within bf16 noise of fp32, and of each other, but not identical."""
    cfg, model = tiny
    text, images = ko.make_inputs(cfg, 2, 20, seed=2)
    with torch.no_grad():
        want = model(text, images)
        model.set_emulation(True, fold=True)
        a = model(text, images)
        model.set_emulation(True, fold=False)
        b = model(text, images)
        model.set_emulation(False)
    assert (a - want).abs().max() < 8e-2 and (b - want).abs().max() < 8e-2
    assert 0 < (a - b).abs().max() < 8e-2


def test_lr_schedules_equal_transformers():
    """train.py:206-251 builds transformers' get_cosine_schedule_with_warmup / get_linear_schedule_with_warmup; the
    multipliers of kosmosx.train must be theirs, step for step (optimizer step k uses the value after k-1 scheduler steps)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kosmos-x_b200"))
    from transformers import get_cosine_schedule_with_warmup, get_linear_schedule_with_warmup
    from kosmosx.train import cosine_with_warmup, linear_with_warmup
    for make_hf, make_mine in ((get_cosine_schedule_with_warmup, cosine_with_warmup), (get_linear_schedule_with_warmup, linear_with_warmup)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=1.0)
        hf = make_hf(opt, num_warmup_steps=7, num_training_steps=60)
        mine = make_mine(7, 60)
        for k in range(1, 70):
            assert abs(opt.param_groups[0]["lr"] - mine(k)) < 1e-12, (k, opt.param_groups[0]["lr"], mine(k))
            opt.step()
            hf.step()
