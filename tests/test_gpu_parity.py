"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances (stated per north_star: floating point, bf16 tensor-core operands, fp32 accumulate):
  * vs the oracle run with bf16 operand rounding at the same points (``emulate_bf16``): the only
    differences left are accumulation order, exp2/erf implementations and where a rounding lands
    on a tie — TOL_EMU below;
  * vs the fp32 oracle: bf16 operand rounding itself (spacing 7.8e-3 at 1.0) through every layer —
    TOL_F32 below.  north_star's 1e-3 is below one bf16 ulp of an O(1) logit, so it is only
    meaningful against the bf16-emulating oracle and is reported (printed) rather than asserted.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
# Every bound below is <= 1.5 x the largest value measured on B200 in round 2 (profiles/r2_parity_measured.md); the
# per-stage growth of the error is tabulated in profiles/r2_stage_errors_*.md.
TOL_EMU_TINY = 3.0e-2    # tiny config logits (std ~1): max-abs vs bf16-emulating oracle (measured 1.48e-2 .. 2.05e-2)
RMS_EMU_TINY = 5.0e-3    # ... and RMS (measured 3.0e-3 .. 3.4e-3)
TOL_F32_TINY = 3.2e-2    # tiny config logits: max-abs vs fp32 oracle (measured 1.73e-2 .. 2.12e-2)
TOL_EMU_FULL = 5.0e-2    # full-size (24+24 layers): max-abs vs bf16-emulating oracle (measured 3.12e-2 .. 3.40e-2)
TOL_F32_FULL = 6.0e-2    # full-size: max-abs vs fp32 oracle, logit std 1.0 (measured 3.99e-2)
RMS_F32_FULL = 1.2e-2    # full-size: RMS error vs fp32 oracle (measured 7.95e-3)
TOL_DEC_FULL = 4.0e-2    # full-size KV-cache decoding vs own full forward / vs the emulating oracle (measured 2.0e-2 .. 2.6e-2)


def _err(got, ref):
    d = (got.float().cpu() - ref.float().cpu())
    return d.abs().max().item(), d.pow(2).mean().sqrt().item()


@pytest.fixture(scope="module")
def tiny_pair(tiny_cfgs):
    import kosmos_oracle as ko
    from kosmosx import Kosmos
    oc, kc = tiny_cfgs
    ref = ko.build(oc, seed=0)
    mine = Kosmos(config=kc)
    mine.load_state_dict(ref.state_dict())
    return ref, mine.cuda(), oc


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(HERE, "golden", "tiny_golden.pt"), weights_only=False)


# --------------------------------------------------------------------------- per-kernel checks
def _kernel_cases():
    import kernel_check as kc
    return [c for c in kc.CASES if not c.startswith("bench")]


@pytest.mark.parametrize("case", ["gemm_basic", "gemm_shapes", "gemm_epilogue", "xpos", "gemm_qkv", "attn", "layernorm",
                                  "ln_fold", "embed", "perceiver_attn", "gemm_trans", "train_elementwise", "attn_bwd",
                                  "perceiver_bwd", "decode", "preprocess", "accurate", "attn_dropout"])
def test_kernel_against_torch_fp32(case):
    """Each kernel alone against a plain PyTorch fp32 restatement of the same op (tools/kernel_check.py)."""
    import kernel_check as kc
    assert set(_kernel_cases()) >= {case}
    assert kc.CASES[case](), f"kernel check {case} failed (see captured stdout)"
    torch.cuda.synchronize()


# --------------------------------------------------------------------------- tiny model vs oracle + golden
@pytest.mark.parametrize("name", ["b2_t10", "b1_t50", "b3_t130"])
def test_tiny_forward_matches_golden_and_oracle(tiny_pair, golden, name):
    import kosmos_oracle as ko
    ref, mine, oc = tiny_pair
    g = golden["cases"][name]
    text, images = ko.make_inputs(oc, g["B"], g["t_text"], seed=1)
    n0 = __import__("kosmosx").ops.launch_count()
    out = mine(text.cuda(), images.cuda())
    torch.cuda.synchronize()
    assert __import__("kosmosx").ops.launch_count() > n0, "no kernel of libkosmosx_sm100.so was launched"
    mine.check_tokens()
    assert out.shape == (g["B"], g["t_text"] + oc.p_latents, oc.vocab) and out.dtype == torch.float32
    assert torch.isfinite(out).all()
    sub = out[..., ::g["col_step"]]
    e_emu = _err(sub, g["logits_emu_bf16"])
    e_f32 = _err(sub, g["logits"])
    print(f"{name}: vs bf16-emulating oracle max={e_emu[0]:.3e} rms={e_emu[1]:.3e}; vs fp32 oracle max={e_f32[0]:.3e} "
          f"rms={e_f32[1]:.3e}")
    assert e_emu[0] <= TOL_EMU_TINY and e_emu[1] <= RMS_EMU_TINY
    assert e_f32[0] <= TOL_F32_TINY
    # and against the live oracle on every vocabulary column
    with torch.no_grad():
        ref.set_emulation(True)
        live = ref(text, images)
        ref.set_emulation(False)
    assert _err(out, live)[0] <= TOL_EMU_TINY


def test_error_is_below_the_reference_path_run_in_bf16(tiny_pair):
    """north_star quotes "logits max-abs-diff <= 1e-3 bf16".  One bf16 ulp of an O(1) logit is 7.8e-3, so no path with
    bf16 operands can be within 1e-3 of fp32 arithmetic; the meaningful bar is the error the REFERENCE's own eager path
    has when it is run in bf16 (``model.bfloat16()``, what train.py's mixed precision does).  Here the oracle (the
    restated reference path) runs in plain torch bf16 on the same GPU and both are measured against the fp32 oracle:
    the sm_100a path (fp32 residual stream, fp32 LayerNorm / softmax statistics) must not be worse."""
    import copy
    import kosmos_oracle as ko
    ref, mine, oc = tiny_pair
    text, images = ko.make_inputs(oc, 3, 40, seed=2)
    with torch.no_grad():
        want = ref(text, images)
        ref16 = copy.deepcopy(ref).to(device="cuda", dtype=torch.bfloat16)
        eager16 = ref16(text.cuda(), images.cuda().bfloat16()).float()
    got = mine(text.cuda(), images.cuda())
    e_mine, e_eager = _err(got, want), _err(eager16, want)
    print(f"vs fp32 oracle: sm_100a path max={e_mine[0]:.3e} rms={e_mine[1]:.3e}; reference path in torch bf16 "
          f"max={e_eager[0]:.3e} rms={e_eager[1]:.3e}")
    assert e_mine[1] <= e_eager[1] and e_mine[0] <= e_eager[0]


def test_tiny_stages_match_golden(tiny_pair, golden):
    """ViT output, spliced decoder input (image rows at 2..65, positions added) per stage."""
    import kosmos_oracle as ko
    ref, mine, oc = tiny_pair
    g = golden["cases"]["b2_t10"]
    text, images = ko.make_inputs(oc, g["B"], g["t_text"], seed=1)
    B, T = g["B"], g["t_text"] + oc.p_latents
    xv = mine._vit(images.cuda().float()).view(B, oc.vit_tokens, oc.vit_dim)
    ev = _err(xv[:, ::4], g["vit"])
    print(f"stages: ViT output vs fp32 golden max={ev[0]:.3e} rms={ev[1]:.3e}")
    assert ev[0] <= 1.2e-2                                          # measured 5.9e-3 (tools/stage_errors.py)
    with torch.no_grad():
        st = ref.stages(text, images)
    dp = mine.decoder._pack()
    x0 = torch.empty(B * T, oc.dim, device="cuda")
    mine._perceive_project(xv.view(-1, oc.vit_dim), B, x0, T, img_rows=(2,))
    from kosmosx import ops
    ops.embed_splice_pos(text.cuda(), dp["embed"], dp["pos"], x0, img_rows=(2,), n_img=oc.p_latents,
                         alias_positions=oc.alias_embed_positions)
    x0 = x0.view(B, T, oc.dim)
    assert _err(x0[:, :2], st["x0"][:, :2])[0] <= 1e-6            # text rows: exact gather + fp32 add
    assert _err(x0[:, 66:], st["x0"][:, 66:])[0] <= 1e-6
    ex = _err(x0[:, 2:66], st["x0"][:, 2:66])
    print(f"stages: image rows of x0 vs fp32 oracle max={ex[0]:.3e} rms={ex[1]:.3e}")
    assert ex[0] <= 1.6e-2                                          # image rows through ViT+perceiver (bf16 operands; measured 1.06e-2)
    assert _err(x0[:, ::2], g["x0"])[0] <= 1.6e-2


@pytest.fixture(scope="module")
def tiny512_pair():
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig
    oc = ko.OracleConfig.tiny(max_positions=512)
    ref = ko.build(oc, seed=0)
    mine = Kosmos(config=KosmosConfig(**{k: getattr(oc, k) for k in KosmosConfig.__dataclass_fields__}))
    mine.load_state_dict(ref.state_dict())
    return ref, mine.cuda(), oc


@pytest.mark.parametrize("name", ["m4", "m2_edges", "m3"])
def test_tiny_multi_image_splice(tiny512_pair, name):
    """BASELINE.json configs[4] shape in small: m images per sequence, each resampled with its own
    media_pos_emb row and spliced in front of text token positions[i] (also: adjacent images, an image at the
    very start / very end of the text).  Decoder input rows exact for text, bf16-tolerance for image rows;
    logits against the golden fixture and the live bf16-emulating oracle; CUDA-graph replay identical."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos, ops
    ref, mine, oc = tiny512_pair
    g = torch.load(os.path.join(HERE, "golden", "tiny_golden_multi.pt"), weights_only=False)["cases"][name]
    B, t_text, positions = g["B"], g["t_text"], g["positions"]
    m = len(positions)
    text, images = ko.make_inputs(oc, B, t_text, seed=7, n_images=m)
    with torch.no_grad():
        st = ref.stages(text, images, positions)
        ref.set_emulation(True)
        want = ref(text, images, image_positions=positions)
        ref.set_emulation(False)
    got = mine(text.cuda(), images.cuda(), image_positions=positions).clone()
    torch.cuda.synchronize()
    T = t_text + m * oc.p_latents
    assert got.shape == (B, T, oc.vocab)
    # the front end alone (the decoder updates the fp32 residual stream in place)
    xv = mine._vit(images.cuda().float().reshape(-1, 3, oc.image, oc.image), media=m)
    rows = tuple(p + i * oc.p_latents for i, p in enumerate(positions))
    x0 = torch.zeros(B * T, oc.dim, device="cuda")
    mine._perceive_project(xv, B, x0, T, rows)
    dp = mine.decoder._pack()
    ops.embed_splice_pos(text.cuda(), dp["embed"], dp["pos"], x0, img_rows=rows, n_img=oc.p_latents,
                         alias_positions=oc.alias_embed_positions)
    x0 = x0.view(B, T, oc.dim).cpu()
    is_img = torch.zeros(T, dtype=torch.bool)
    for r in rows:
        is_img[r:r + oc.p_latents] = True
    assert _err(x0[:, ~is_img], st["x0"][:, ~is_img])[0] <= 1e-6
    ei = _err(x0[:, is_img], st["x0"][:, is_img])
    print(f"multi-image {name}: image rows of x0 vs fp32 oracle max={ei[0]:.3e}")
    assert ei[0] <= 2.0e-2
    assert _err(x0[:, ::2], g["x0"])[0] <= 2.0e-2
    e = _err(got, want)
    eg = _err(got[..., ::g["col_step"]], g["logits_emu_bf16"])
    print(f"multi-image {name} m={m} at {positions}: vs bf16-emulating oracle max={e[0]:.3e} rms={e[1]:.3e}; golden max={eg[0]:.3e}")
    assert e[0] <= TOL_EMU_TINY and e[1] <= RMS_EMU_TINY and eg[0] <= TOL_EMU_TINY
    assert _err(got[..., ::g["col_step"]], g["logits"])[0] <= TOL_F32_TINY
    graphed = Kosmos(config=mine.cfg, cuda_graph=True)
    graphed.load_state_dict(ref.state_dict())
    graphed = graphed.cuda()
    g1 = graphed(text.cuda(), images.cuda(), image_positions=positions).clone()
    g2 = graphed(text.cuda(), images.cuda(), image_positions=positions).clone()
    assert torch.equal(g1, got) and torch.equal(g2, got)
    with pytest.raises(ValueError, match="image_positions"):
        mine(text.cuda(), images.cuda(), image_positions=positions[:-1])


def test_tiny_properties(tiny_pair):
    """Size-independent properties of the reference path: causality, batch independence, determinism."""
    import kosmos_oracle as ko
    _, mine, oc = tiny_pair
    text, images = ko.make_inputs(oc, 3, 70, seed=5)
    text, images = text.cuda(), images.cuda()
    a = mine(text, images).clone()
    b = mine(text, images).clone()
    assert torch.equal(a, b), "forward is not deterministic"
    t2 = text.clone()
    t2[:, 40:] = (t2[:, 40:] + 7) % oc.vocab
    c = mine(t2, images).clone()
    cut = 40 + oc.p_latents                                       # text token 40 sits at spliced row 104
    assert torch.equal(a[:, :cut], c[:, :cut]), "logits before a changed token moved: causality broken"
    assert not torch.equal(a[:, cut:], c[:, cut:])
    one = mine(text[1:2], images[1:2]).clone()
    assert torch.allclose(one[0], a[1], atol=1e-5, rtol=0), "sample depends on its batch neighbours"


def test_multiway_b_branch_is_inert_and_cuda_graph_matches(tiny_pair):
    import kosmos_oracle as ko
    from kosmosx import Kosmos
    ref, mine, oc = tiny_pair
    text, images = ko.make_inputs(oc, 2, 12, seed=3)
    text, images = text.cuda(), images.cuda()
    a = mine(text, images).clone()
    sd = {k: (torch.randn_like(v) if ".B." in k else v) for k, v in ref.state_dict().items()}
    other = Kosmos(config=mine.cfg, cuda_graph=True)
    other.load_state_dict(sd)
    other = other.cuda()
    b = other(text, images).clone()
    b2 = other(text, images).clone()                              # second call replays the captured graph
    assert torch.equal(a, b) and torch.equal(a, b2)


def test_reference_call_patterns(tiny_pair):
    """example.py:9 passes images.long(); forward_embedding / passed_x used as at model.py:238-250."""
    import kosmos_oracle as ko
    ref, mine, oc = tiny_pair
    text, images = ko.make_inputs(oc, 1, 20, seed=2)
    li = (images * 3).long()
    with torch.no_grad():
        ref.set_emulation(True)
        want = ref(text, li)
        ref.set_emulation(False)
    got = mine(text.cuda(), li.cuda())
    assert _err(got, want)[0] <= TOL_EMU_TINY
    # decoder surface: (x, embed) and passed_x
    x, emb = mine.decoder.forward_embedding(text.cuda())
    with torch.no_grad():
        rx, remb = ref.decoder.forward_embedding(text)
    assert _err(x, rx)[0] <= 1e-6 and _err(emb, remb)[0] <= 1e-6
    assert emb.data_ptr() == x.data_ptr() and remb.data_ptr() == rx.data_ptr()     # torchscale's in-place `x += positions`: embed IS x
    logits, extra = mine.decoder(text.cuda(), passed_x=x)
    with torch.no_grad():
        ref.set_emulation(True)
        rl = ref.decoder(rx, passed_x=rx)[0]
        ref.set_emulation(False)
    assert _err(logits, rl)[0] <= TOL_EMU_TINY
    assert set(extra) == {"inner_states", "l_aux", "attn"}


def test_reference_forward_spelled_out_through_the_decoder_surface(tiny_pair):
    """model.py:238-250 written with the public decoder calls — forward_embedding(text)[1], torch.cat with the image rows,
    forward_embedding(x, token_embedding=x)[0], decoder(x, passed_x=x) — equals Kosmos.forward: text rows carry the
    position of the un-spliced text AND the spliced one (aliased `embed`), image rows one."""
    import kosmos_oracle as ko
    ref, mine, oc = tiny_pair
    text, images = ko.make_inputs(oc, 2, 24, seed=9)
    want = mine(text.cuda(), images.cuda()).clone()
    st = mine.stages(text.cuda(), images.cuda())
    rows = st["x0"][:, 2:2 + oc.p_latents] - mine.embed_positions.weight[4:4 + oc.p_latents].float()     # image_proj output
    model_input = mine.decoder.forward_embedding(text.cuda())[1]
    model_input = torch.cat([model_input[:, 0:2], rows, model_input[:, 2:]], dim=1)
    model_input = mine.decoder.forward_embedding(model_input, token_embedding=model_input)[0]
    assert _err(model_input, st["x0"])[0] <= 1e-5
    got = mine.decoder(model_input, passed_x=model_input)[0]
    assert _err(got, want)[0] <= 2e-3
    with torch.no_grad():
        assert _err(st["x0"][:, 66:], ref.stages(text, images)["x0"][:, 66:])[0] <= 1e-6


def test_out_of_place_position_reading_matches_its_golden(tiny_cfgs, golden):
    """KosmosConfig.alias_embed_positions = False (one positional embedding per text row) against the fixture minted with
    OracleConfig.alias_embed_positions = False."""
    import dataclasses
    import kosmos_oracle as ko
    from kosmosx import Kosmos
    oc, kc = tiny_cfgs
    ref = ko.build(oc, seed=0)
    mine = Kosmos(config=dataclasses.replace(kc, alias_embed_positions=False))
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    g = golden["noalias_b1_t50"]
    text, images = ko.make_inputs(oc, g["B"], g["t_text"], seed=1)
    st = mine.stages(text.cuda(), images.cuda())
    assert _err(st["x0"][:, ::2][:, 33:], g["x0"][:, 33:])[0] <= 2e-3       # text rows 66..; fixture = every 2nd row, fp16
    e = _err(st["logits"][..., ::g["col_step"]], g["logits"])
    print(f"alias_embed_positions=False: vs fp32 oracle max={e[0]:.3e} rms={e[1]:.3e}")
    assert e[0] <= TOL_F32_TINY
    aliased = golden["cases"]["b1_t50"]["logits"]
    assert _err(g["logits"], aliased)[0] > 0.1, "the two readings of forward_embedding must differ measurably"


# --------------------------------------------------------------------------- verification precision (bf16x3)
TOL_STATED = 1e-3        # BASELINE.json north_star: "logits max-abs-diff <= 1e-3"


@pytest.mark.parametrize("name", ["b2_t10", "b1_t50", "b3_t130"])
def test_bf16x3_mode_meets_the_stated_tolerance_tiny(tiny_pair, golden, name):
    """Kosmos.forward(precision="bf16x3") — the same tcgen05 GEMM kernel on split operands, fp32 everywhere else — is within
    the tolerance BASELINE.json states (1e-3) of the fp32 oracle and of the fp32 golden logits."""
    import kosmos_oracle as ko
    ref, mine, oc = tiny_pair
    g = golden["cases"][name]
    text, images = ko.make_inputs(oc, g["B"], g["t_text"], seed=1)
    n0 = __import__("kosmosx").ops.launch_count()
    got = mine(text.cuda(), images.cuda(), precision="bf16x3")
    assert __import__("kosmosx").ops.launch_count() > n0
    with torch.no_grad():
        want = ref(text, images)
    e, eg = _err(got, want), _err(got[..., ::g["col_step"]], g["logits"])
    print(f"bf16x3 {name}: vs fp32 oracle max={e[0]:.3e} rms={e[1]:.3e}; golden max={eg[0]:.3e}")
    assert got.shape == want.shape and e[0] <= TOL_STATED and eg[0] <= TOL_STATED


def test_bf16x3_multi_image_and_language(tiny512_pair):
    import kosmos_oracle as ko
    from kosmosx import KosmosLanguage
    ref, mine, oc = tiny512_pair
    pos = [2, 9, 9, 30]
    text, images = ko.make_inputs(oc, 2, 30, seed=7, n_images=4)
    with torch.no_grad():
        want = ref(text, images, image_positions=pos)
    e = _err(mine(text.cuda(), images.cuda(), image_positions=pos, precision="bf16x3"), want)
    print(f"bf16x3 multi-image: vs fp32 oracle max={e[0]:.3e} rms={e[1]:.3e}")
    assert e[0] <= TOL_STATED
    oc2 = ko.OracleConfig.tiny(vocab=777)
    torch.manual_seed(0)
    lref = ko.KosmosLanguageOracle(oc2).eval()
    lm = KosmosLanguage(vocab_size=777, dim=oc2.dim, depth=oc2.layers, ffn_dim=oc2.ffn, decoder_heads=oc2.heads,
                        max_positions=oc2.max_positions, precision="bf16x3")
    lm.load_state_dict(lref.state_dict())
    lm = lm.cuda()
    x = torch.randint(0, 777, (2, 37), generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        lw = lref(x)
    el = _err(lm(x.cuda()), lw)
    print(f"bf16x3 KosmosLanguage: vs fp32 oracle max={el[0]:.3e}")
    assert el[0] <= TOL_STATED


def test_bf16_logits_and_owned_graph_outputs(tiny_pair):
    """logits_dtype=torch.bfloat16: the LM head stores bf16 rows (16-byte pitch, TMA-store epilogue) == the fp32 logits
    rounded once.  cuda_graph=True returns tensors the caller owns (a later call does not overwrite them) unless
    graph_alias_output=True."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos
    ref, mine, oc = tiny_pair
    text, images = ko.make_inputs(oc, 2, 30, seed=6)
    want = mine(text.cuda(), images.cuda()).clone()
    m16 = Kosmos(config=mine.cfg, logits_dtype=torch.bfloat16)
    m16.load_state_dict(ref.state_dict())
    m16 = m16.cuda()
    got = m16(text.cuda(), images.cuda())
    assert got.dtype == torch.bfloat16 and got.shape == want.shape
    assert torch.equal(got.float(), want.bfloat16().float())
    text2, images2 = ko.make_inputs(oc, 2, 30, seed=7)
    for alias in (False, True):
        gm = Kosmos(config=mine.cfg, cuda_graph=True, graph_alias_output=alias, logits_dtype=torch.bfloat16)
        gm.load_state_dict(ref.state_dict())
        gm = gm.cuda()
        a = gm(text.cuda(), images.cuda())
        a_copy = a.clone()
        b = gm(text2.cuda(), images2.cuda())
        c = gm(text2.cuda(), images2.cuda())
        torch.cuda.synchronize()
        assert torch.equal(a_copy, got) and torch.equal(b, c)
        assert torch.equal(a, a_copy) != alias, "owned outputs survive later calls; aliased ones are overwritten by the second-next"


def test_error_behaviour_on_gpu(tiny_pair):
    _, mine, oc = tiny_pair
    img = torch.zeros(1, 3, oc.image, oc.image, device="cuda")
    with pytest.raises(ValueError, match="exceeds the positional table"):
        mine(torch.zeros(1, oc.max_positions - oc.p_latents, dtype=torch.long, device="cuda"), img)
    with pytest.raises(ValueError, match="doesn't match model"):
        mine(torch.zeros(1, 8, dtype=torch.long, device="cuda"), torch.zeros(1, 3, 64, 64, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU path"):
        mine(torch.zeros(1, 8, dtype=torch.long), img.cpu())
    mine(torch.full((1, 8), oc.vocab + 5, dtype=torch.long, device="cuda"), img)     # flagged on device, no fault
    with pytest.raises(ValueError, match="out of range"):
        mine.check_tokens()
    # longest legal sequence: T = max_positions - 2
    out = mine(torch.zeros(1, oc.max_positions - 2 - oc.p_latents, dtype=torch.long, device="cuda"), img)
    assert out.shape[1] == oc.max_positions - 2 and torch.isfinite(out).all()


def test_language_model_matches_oracle():
    import kosmos_oracle as ko
    from kosmosx import KosmosLanguage
    oc = ko.OracleConfig.tiny(vocab=777)
    torch.manual_seed(0)
    ref = ko.KosmosLanguageOracle(oc, emulate_bf16=True).eval()
    mine = KosmosLanguage(vocab_size=777, dim=oc.dim, depth=oc.layers, ffn_dim=oc.ffn, decoder_heads=oc.heads,
                          max_positions=oc.max_positions)
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    x = torch.randint(0, 777, (2, 37), generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        want = ref(x)
    got = mine(x.cuda())
    assert got.shape == (2, 37, 777)
    assert _err(got, want)[0] <= TOL_EMU_TINY


# --------------------------------------------------------------------------- full size (BASELINE.json configs)

# --------------------------------------------------------------------------- incremental decoding
TOL_DEC_SELF = 1.9e-2    # KV-cache path vs the same model's full forward (bf16 operands both ways; measured 0.84e-2 .. 1.25e-2)


def test_tiny_incremental_decoding_vs_oracle(tiny_pair):
    """generate() through the KV-cache kernels against the oracle's restatement of torchscale's incremental protocol
    (teacher-forced so both see the same tokens), against the model's own full forward over prompt + continuation,
    graph replay against eager stepping, and the torchscale calling protocol through Decoder.forward."""
    import kosmos_oracle as ko
    ref, mine, oc = tiny_pair
    B, t_text, n = 2, 12, 7
    text, images = ko.make_inputs(oc, B, t_text, seed=1)
    forced = torch.randint(0, oc.vocab, (B, n), generator=torch.Generator().manual_seed(5))
    ref.set_emulation(True)
    _, want = ref.generate(text, images, n, forced=forced)
    ref.set_emulation(False)
    n0 = __import__("kosmosx").ops.launch_count()
    toks, got = mine.generate(text.cuda(), images.cuda(), n, forced_tokens=forced.cuda(), return_logits=True, one_kernel=True)
    torch.cuda.synchronize()
    assert __import__("kosmosx").ops.launch_count() > n0
    assert torch.equal(toks.cpu(), forced) and got.shape == (B, n, oc.vocab) and torch.isfinite(got).all()
    e = _err(got, want)
    print(f"incremental decoding vs bf16-emulating oracle: max={e[0]:.3e} rms={e[1]:.3e}")
    assert e[0] <= TOL_EMU_TINY and e[1] <= RMS_EMU_TINY
    # the model's own full forward over prompt + forced continuation: rows T0-1 .. T0+n-2
    t0 = t_text + oc.p_latents
    full = mine(torch.cat([text, forced[:, :-1]], 1).cuda(), images.cuda())
    es = _err(got, full[:, t0 - 1:])
    print(f"incremental decoding vs own full forward: max={es[0]:.3e} rms={es[1]:.3e}")
    assert es[0] <= TOL_DEC_SELF
    # the per-kernel path (kx_decode_linear / kx_decode_attn launches) against the one-kernel step (default for B <= 8)
    toks_pk, got_pk = mine.generate(text.cuda(), images.cuda(), n, forced_tokens=forced.cuda(), return_logits=True,
                                    one_kernel=False)
    ep = _err(got_pk, got)
    print(f"per-kernel decoding vs one-kernel step: max={ep[0]:.3e}")
    assert ep[0] <= TOL_DEC_SELF and _err(got_pk, want)[0] <= TOL_EMU_TINY and torch.equal(toks_pk.cpu(), forced)
    # greedy: one-kernel step, CUDA-graph replay of the per-kernel step and eager stepping agree; deterministic
    g0 = mine.generate(text.cuda(), images.cuda(), n, one_kernel=True)
    g1 = mine.generate(text.cuda(), images.cuda(), n, one_kernel=False)
    g2 = mine.generate(text.cuda(), images.cuda(), n, one_kernel=False, cuda_graph=False)
    g3 = mine.generate(text.cuda(), images.cuda(), n, forced_tokens=forced.cuda())
    assert torch.equal(g1, g2) and torch.equal(g3.cpu(), forced) and torch.equal(g0, mine.generate(text.cuda(), images.cuda(), n, one_kernel=True))
    assert (g0 == g1).float().mean() >= 0.8            # (different summation order: a near-tie may flip)
    assert int(mine._last_decode_state.err.item()) == 0
    assert torch.equal(g1[:, 0], mine(text.cuda(), images.cuda())[:, -1].argmax(-1))
    # torchscale's protocol: decoder(x, incremental_state={"is_first_step": True}, passed_x=x), then whole prefixes
    with torch.no_grad():
        x = ref.embed_inputs(text, images).cuda()
    st = {"is_first_step": True, "max_length": t0 + n}
    first, extra = mine.decoder(x, incremental_state=st, passed_x=x)
    assert first.shape == (B, t0, oc.vocab) and "kx_state" in st and st["kx_state"].length == t0
    assert _err(first[:, -1], got[:, 0])[0] <= TOL_DEC_SELF
    st["is_first_step"] = False
    # Decoder.forward's own later steps are torchscale's, literally: forward_embedding(prefix) embeds the last token with ONE
    # position, len(prefix)+1 (Kosmos.generate instead continues the text as Kosmos.forward would embed it) — so the
    # protocol is checked against the oracle's restatement of the same protocol
    ref.set_emulation(True)
    st_ref = {"is_first_step": True}
    with torch.no_grad():
        ref.decoder(x.cpu(), incremental_state=st_ref, passed_x=x.cpu())
    st_ref["is_first_step"] = False
    prefix = torch.zeros(B, t0, dtype=torch.int64, device="cuda")
    for i in range(n - 1):
        prefix = torch.cat([prefix, forced[:, i:i + 1].cuda()], 1)
        step_logits, _ = mine.decoder(prefix, incremental_state=st)
        assert step_logits.shape == (B, 1, oc.vocab)
        with torch.no_grad():
            want_step = ref.decoder(prefix.cpu(), incremental_state=st_ref)[0]
        assert _err(step_logits, want_step)[0] <= TOL_EMU_TINY, "Decoder.forward incremental protocol differs from the oracle's"
    ref.set_emulation(False)
    with pytest.raises(ValueError):
        mine.decoder(prefix, incremental_state=st)                    # prefix did not grow
    with pytest.raises(ValueError):
        mine.decoder(prefix, incremental_state={"is_first_step": False})
    with pytest.raises(ValueError):
        mine.generate(text.cuda(), images.cuda(), oc.max_positions)   # beyond the position table


def test_generate_edge_cases(tiny512_pair):
    """Decoding after a multi-image prompt (the cache continues the spliced sequence), a single sequence, a single new
    token, a cache that crosses attention-chunk boundaries (128 / 256 keys), and the per-kernel and one-kernel paths on
    each — all against the oracle's incremental restatement with forced tokens."""
    import kosmos_oracle as ko
    ref, mine, oc = tiny512_pair
    cases = [  # (B, t_text, image positions, new tokens)
        (2, 30, [2, 17], 5),          # two images: T0 = 158, crosses the 128-key chunk of kx_decode_attn
        (1, 60, None, 9),             # T0 = 124 -> 133: the newest row opens a new 128-key chunk mid-generation
        (3, 190, None, 6),            # T0 = 254 -> 260: crosses the 256-key chunk of the one-kernel step
        (2, 12, None, 1),             # a single new token: the prompt pass alone
    ]
    for B, t_text, positions, n in cases:
        m = 1 if positions is None else len(positions)
        text, images = ko.make_inputs(oc, B, t_text, seed=11, n_images=None if positions is None else m)
        forced = torch.randint(0, oc.vocab, (B, n), generator=torch.Generator().manual_seed(3))
        ref.set_emulation(True)
        _, want = ref.generate(text, images, n, image_positions=positions, forced=forced)
        ref.set_emulation(False)
        for one in (False, True):
            toks, got = mine.generate(text.cuda(), images.cuda(), n, image_positions=positions, forced_tokens=forced.cuda(),
                                      return_logits=True, one_kernel=one)
            e = _err(got, want)
            print(f"generate B={B} T0={t_text + 64 * m} n={n} one_kernel={one}: max={e[0]:.3e} rms={e[1]:.3e}")
            assert torch.equal(toks.cpu(), forced) and e[0] <= TOL_EMU_TINY and e[1] <= RMS_EMU_TINY
            free = mine.generate(text.cuda(), images.cuda(), n, image_positions=positions, one_kernel=one)
            assert free.shape == (B, n) and int(free.min()) >= 0 and int(free.max()) < oc.vocab
            assert int(mine._last_decode_state.err.item()) == 0


def test_language_model_generate_batches(tiny_cfgs):
    """KosmosLanguage.generate at batch sizes that use every batch-group width of kx_decode_linear (<=8, <=16, <=32)."""
    import kosmos_oracle as ko
    from kosmosx import KosmosLanguage
    oc, _ = tiny_cfgs
    lm_ref = ko.KosmosLanguageOracle(oc)
    lm = KosmosLanguage(vocab_size=oc.vocab, dim=oc.dim, depth=oc.layers, ffn_dim=oc.ffn, decoder_heads=oc.heads,
                        max_positions=oc.max_positions)
    lm.load_state_dict(lm_ref.state_dict())
    lm = lm.cuda()
    for B in (1, 11, 32):
        x = torch.randint(0, oc.vocab, (B, 9), generator=torch.Generator().manual_seed(B))
        n = 5
        forced = torch.randint(0, oc.vocab, (B, n), generator=torch.Generator().manual_seed(100 + B))
        toks, got = lm.generate(x.cuda(), n, forced_tokens=forced.cuda(), return_logits=True)
        full = lm(torch.cat([x, forced[:, :-1]], 1).cuda())
        e = _err(got, full[:, 8:])
        print(f"KosmosLanguage.generate B={B}: vs own full forward max={e[0]:.3e}")
        assert e[0] <= TOL_DEC_SELF
        with torch.no_grad():
            want = lm_ref(torch.cat([x, forced[:, :-1]], 1))[:, 8:]
        ef = _err(got, want)
        print(f"KosmosLanguage.generate B={B}: vs fp32 oracle max={ef[0]:.3e}")
        assert ef[0] <= TOL_F32_TINY
    with pytest.raises(ValueError):
        lm.generate(torch.zeros(33, 4, dtype=torch.int64, device="cuda"), 2)

# --------------------------------------------------------------------------- device-side preprocessing
def _u8_images(n, image, channels_last, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (n, 3, image, image), dtype=torch.uint8, generator=g)
    img[0, :, :16, :16] = torch.arange(256, dtype=torch.uint8).view(1, 16, 16)      # every (value, channel) pair
    return img.permute(0, 2, 3, 1).contiguous() if channels_last else img


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("image", [56, 224])
def test_clip_normalize_u8_is_bit_exact_vs_oracle(image, channels_last):
    """kx_clip_normalize_u8 against the oracle's restatement of CLIPImageProcessor's rescale + normalise (itself pinned
    bit for bit against the installed transformers in tests/test_preprocess.py): fp32, bit-exact."""
    import kosmos_oracle as ko
    from kosmosx import ops
    img = _u8_images(3, image, channels_last, seed=image)
    want = ko.clip_preprocess_u8(img, channels_last)
    got = ops.clip_normalize_u8(img.cuda(), image=image)
    assert got.dtype == torch.float32 and torch.equal(got.cpu(), want)
    # other statistics (a processor with its own image_mean / image_std)
    mean, std = (0.5, 0.25, 0.125), (0.5, 0.3, 0.7)
    assert torch.equal(ops.clip_normalize_u8(img.cuda(), image=image, mean=mean, std=std).cpu(),
                       ko.clip_preprocess_u8(img, channels_last, mean, std))
    with pytest.raises(ValueError):
        ops.clip_normalize_u8(torch.zeros(1, 3, image, image + 4, dtype=torch.uint8, device="cuda"), image=image)
    with pytest.raises(RuntimeError, match="std"):
        ops.clip_normalize_u8(img.cuda(), image=image, std=(1.0, 0.0, 1.0))


@pytest.mark.parametrize("channels_last", [False, True])
def test_fused_u8_patch_pack_equals_normalize_then_pack(tiny_pair, channels_last):
    """kx_im2col_patches_u8 == kx_im2col_patches(kx_clip_normalize_u8(.)) bit for bit, including the media-major slot
    order of the multi-image form and the CLS rows."""
    from kosmosx import ops
    _, mine, oc = tiny_pair
    vp = mine._pack_vision()
    n, media = 6, 2
    P, Tv, Dv = oc.vit_tokens - 1, oc.vit_tokens, oc.vit_dim
    img = _u8_images(n, oc.image, channels_last, seed=11).cuda()
    outs = []
    for fused in (False, True):
        patches = torch.full((n * P, vp["k_pad"]), 7.0, dtype=torch.bfloat16, device="cuda")
        x = torch.zeros(n, Tv, Dv, dtype=torch.float32, device="cuda")
        if fused:
            ops.im2col_patches_u8(img, patches, vp["cls"], vp["vpos"], x, image=oc.image, patch=oc.patch, media=media)
        else:
            ops.im2col_patches(ops.clip_normalize_u8(img, image=oc.image), patches, vp["cls"], vp["vpos"], x,
                               image=oc.image, patch=oc.patch, media=media)
        outs.append((patches, x))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert outs[1][0].float().abs().max().item() > 1.0          # not all zeros: values up to (1 - mean) / std ~ 2.1


def test_forward_on_raw_uint8_images_equals_forward_on_processor_output(tiny512_pair):
    """Kosmos.forward(normalize_images=True) on raw pixels == Kosmos.forward on KosmosTokenizer.tokenize_images' fp32
    pixel_values (device kernel) == the oracle fed the oracle's preprocessing, single- and multi-image, eager and
    graph replay; generate() takes the same flag.  Without the flag uint8 is cast to float like HF does."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosTokenizer
    ref, mine, oc = tiny512_pair

    class _Tok:
        pad_token_id = 1

        def convert_tokens_to_ids(self, t):
            return [oc.vocab - 2, oc.vocab - 1]

    tk = KosmosTokenizer(tokenizer=_Tok(), processor=object(), image_size=oc.image)
    text, _ = ko.make_inputs(oc, 2, 40, seed=4)
    graphed = Kosmos(config=mine.cfg, cuda_graph=True)
    graphed.load_state_dict(ref.state_dict())
    graphed = graphed.cuda()
    for m, cl, pos in ((1, True, None), (1, False, None), (3, True, [1, 7, 40])):
        raw = _u8_images(2 * m, oc.image, cl, seed=20 + m)
        pv = tk.tokenize_images(raw.cuda())                               # fp32 (2m,3,H,W), device kernel
        assert torch.equal(pv.cpu(), ko.clip_preprocess_u8(raw, cl))
        if m > 1:
            raw, pv = raw.view(2, m, *raw.shape[1:]), pv.view(2, m, *pv.shape[1:])
        want_mine = mine(text.cuda(), pv, image_positions=pos).clone()
        got = mine(text.cuda(), raw.cuda(), image_positions=pos, normalize_images=True).clone()
        assert torch.equal(got, want_mine)
        with torch.no_grad():
            ref.set_emulation(True)
            want = ref(text, pv.cpu(), image_positions=pos)
            ref.set_emulation(False)
        e = _err(got, want)
        print(f"raw uint8 forward m={m} channels_last={cl}: vs bf16-emulating oracle max={e[0]:.3e}")
        assert e[0] <= TOL_EMU_TINY
        g1 = graphed(text.cuda(), raw.cuda(), image_positions=pos, normalize_images=True).clone()
        g2 = graphed(text.cuda(), raw.cuda(), image_positions=pos, normalize_images=True).clone()   # replay, other slot
        assert torch.equal(g1, got) and torch.equal(g2, got)
        t1 = mine.generate(text.cuda(), raw.cuda(), 3, image_positions=pos, normalize_images=True)
        t2 = mine.generate(text.cuda(), pv, 3, image_positions=pos)
        assert torch.equal(t1, t2)
    raw = _u8_images(2, oc.image, False, seed=31)
    assert torch.equal(mine(text.cuda(), raw.cuda()), mine(text.cuda(), raw.float().cuda()))      # [HF]:208-209 cast
    with pytest.raises(TypeError, match="uint8"):
        mine(text.cuda(), raw.float().cuda(), normalize_images=True)
    with pytest.raises(ValueError, match="doesn't match model"):
        mine(text.cuda(), torch.zeros(2, oc.image, oc.image, 3, dtype=torch.uint8, device="cuda"))   # channels-last needs the flag


@pytest.fixture(scope="module")
def full_pair():
    """Reference-size model (24-layer ViT-L/14 + perceiver + 24-layer d=2048 decoder), weights from the
    oracle's seeded random init (multiway .B branches included, as the reference's state_dict)."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig
    torch.set_num_threads(os.cpu_count() or 8)
    oc = ko.OracleConfig(max_positions=2050)
    ref = ko.build(oc, seed=0)
    mine = Kosmos(config=KosmosConfig(max_positions=2050), device="cuda")
    mine.load_state_dict(ref.state_dict())
    return ref, mine, oc


def test_full_size_readme_example_vs_oracle(full_pair):
    """configs[0]: 1 x (3,224,224) image + 50 text tokens (README example), logits vs the CPU oracle."""
    import kosmos_oracle as ko
    ref, mine, oc = full_pair
    text, images = ko.make_inputs(oc, 1, 50, seed=1)
    with torch.no_grad():
        want32 = ref(text, images)
        ref.set_emulation(True)
        want16 = ref(text, images)
        ref.set_emulation(False)
    got = mine(text.cuda(), images.cuda())
    assert got.shape == (1, 114, 32002)
    e16, e32 = _err(got, want16), _err(got, want32)
    print(f"C1 full size: logits std={want32.std():.3f}; vs bf16-emulating oracle max={e16[0]:.3e} rms={e16[1]:.3e}; "
          f"vs fp32 oracle max={e32[0]:.3e} rms={e32[1]:.3e}")
    assert e16[0] <= TOL_EMU_FULL
    assert e32[0] <= TOL_F32_FULL and e32[1] <= RMS_F32_FULL
    agree = (got.cpu().argmax(-1) == want32.argmax(-1)).float().mean().item()
    print(f"C1 argmax agreement with fp32 oracle: {agree:.4f}")
    assert agree >= 0.97


def test_full_size_readme_example_bf16x3_meets_stated_tolerance(full_pair):
    """configs[0] at the reference size in the verification precision: logits within BASELINE.json's 1e-3 of the fp32
    CPU oracle (the bf16 throughput mode is ~4e-2 away: one bf16 ulp at 1.0 is 7.8e-3)."""
    import kosmos_oracle as ko
    ref, mine, oc = full_pair
    text, images = ko.make_inputs(oc, 1, 50, seed=1)
    with torch.no_grad():
        want32 = ref(text, images)
    got = mine(text.cuda(), images.cuda(), precision="bf16x3")
    e = _err(got, want32)
    agree = (got.cpu().argmax(-1) == want32.argmax(-1)).float().mean().item()
    print(f"C1 full size bf16x3: vs fp32 oracle max={e[0]:.3e} rms={e[1]:.3e}; argmax agreement {agree:.4f}")
    assert got.shape == (1, 114, 32002) and e[0] <= TOL_STATED
    assert agree >= 0.999
    mine._accurate().invalidate()          # 1.5x the fp32 weights in split form: give the memory back
    torch.cuda.empty_cache()


def test_full_size_seq2048_properties(full_pair):
    """configs[2] shape (B=8, T=2048): properties that do not need the oracle at this size, plus one
    sequence checked end to end against the CPU oracle run at B=1."""
    import kosmos_oracle as ko
    ref, mine, oc = full_pair
    text, images = ko.make_inputs(oc, 8, 1984, seed=1)
    tg, ig = text.cuda(), images.cuda()
    a = mine(tg, ig)
    assert a.shape == (8, 2048, 32002) and torch.isfinite(a).all()
    a_first = a[:, :1100].clone()
    a3 = a[3].clone()
    del a
    b = mine(tg, ig)
    assert torch.equal(b[3], a3), "forward is not deterministic at full size"
    del b
    t2 = tg.clone()
    t2[:, 1200:] = (t2[:, 1200:] + 11) % oc.vocab
    c = mine(t2, ig)
    assert torch.equal(c[:, :1100], a_first), "causality broken at T=2048"
    del c
    one = mine(tg[3:4], ig[3:4])
    assert torch.allclose(one[0], a3, atol=1e-4, rtol=0), "batch independence broken at B=8"
    with torch.no_grad():
        ref.set_emulation(True)
        want = ref(text[3:4], images[3:4])
        ref.set_emulation(False)
    e = _err(one, want)
    print(f"C3 sequence 3 (T=2048) vs bf16-emulating oracle: max={e[0]:.3e} rms={e[1]:.3e}")
    assert e[0] <= TOL_EMU_FULL


def test_full_size_generate_vs_forward_and_oracle(full_pair):
    """Reference-size decoding (24 layers, d=2048, 32 heads, vocab 32002): KV-cache steps against the model's own full
    forward over prompt + continuation, and against the CPU oracle's incremental path on one sequence."""
    import kosmos_oracle as ko
    ref, mine, oc = full_pair
    B, t_text, n = 8, 40, 6
    text, images = ko.make_inputs(oc, B, t_text, seed=3)
    forced = torch.randint(0, oc.vocab, (B, n), generator=torch.Generator().manual_seed(7))
    toks, got = mine.generate(text.cuda(), images.cuda(), n, forced_tokens=forced.cuda(), return_logits=True)
    t0 = t_text + oc.p_latents
    full = mine(torch.cat([text, forced[:, :-1]], 1).cuda(), images.cuda())
    e = _err(got, full[:, t0 - 1:])
    print(f"full-size incremental decoding vs own full forward: max={e[0]:.3e} rms={e[1]:.3e}")
    assert e[0] <= TOL_DEC_FULL
    agree = (got.argmax(-1) == full[:, t0 - 1:].argmax(-1)).float().mean().item()
    assert agree >= 0.9, agree
    with torch.no_grad():
        ref.set_emulation(True)
        _, want = ref.generate(text[:1], images[:1], n, forced=forced[:1])
        ref.set_emulation(False)
    e16 = _err(got[:1], want)
    print(f"full-size incremental decoding vs bf16-emulating oracle (sequence 0): max={e16[0]:.3e} rms={e16[1]:.3e}")
    assert e16[0] <= TOL_DEC_FULL
    _, got_pk = mine.generate(text.cuda(), images.cuda(), n, forced_tokens=forced.cuda(), return_logits=True, one_kernel=False)
    _, got_one = mine.generate(text.cuda(), images.cuda(), n, forced_tokens=forced.cuda(), return_logits=True, one_kernel=True)
    ep = _err(got_pk, got_one)
    print(f"full-size per-kernel decoding vs one-kernel step: max={ep[0]:.3e} rms={ep[1]:.3e}")
    assert ep[0] <= TOL_DEC_FULL and _err(got_one, full[:, t0 - 1:])[0] <= TOL_DEC_FULL
    g1 = mine.generate(text.cuda(), images.cuda(), n, one_kernel=False)
    g2 = mine.generate(text.cuda(), images.cuda(), n, one_kernel=False, cuda_graph=False)
    assert torch.equal(g1, g2)
    g0 = mine.generate(text.cuda(), images.cuda(), n, one_kernel=True)
    assert torch.equal(g0, mine.generate(text.cuda(), images.cuda(), n, one_kernel=True)), "one-kernel decoding is not deterministic"
    assert int(mine._last_decode_state.err.item()) == 0


def test_full_size_training_gradients_vs_oracle(full_pair):
    """configs[3] arithmetic at the reference size (24 layers, d=2048, 32 heads, vocab 32002) on the README-sized
    input (B=1, T=114) where CPU autograd through the oracle finishes in seconds: loss and the gradient of every
    trained parameter.  Tolerances as tests/test_gpu_train.py (bf16 operands, fp32 accumulate)."""
    import kosmos_oracle as ko
    from kosmosx import KosmosTrainer
    ref, mine, oc = full_pair
    text, images = ko.make_inputs(oc, 1, 50, seed=2)
    trainer = KosmosTrainer(mine, dropout=0.0, attention_dropout=0.0)
    loss = trainer.loss_and_grads(text.cuda(), images.cuda())
    torch.cuda.synchronize()
    names = {id(p): n for n, p in mine.named_parameters()}
    trained = {names[id(p)] for p in trainer.params}
    for n, p in ref.named_parameters():
        p.requires_grad_(n in trained)
    ref.set_emulation(True, fold=False)          # the training forward materialises its LayerNorms (no fold)
    want = ref.loss(text, images)
    want.backward()
    ref.set_emulation(False)
    print(f"full-size training: loss cuda {loss.item():.5f} oracle {want.item():.5f}")
    assert abs(loss.item() - want.item()) <= 2.5e-3           # measured 1.5e-3
    ref_named = dict(ref.named_parameters())
    worst = (0.0, "")
    for p in trainer.params:
        n = names[id(p)]
        g, g_ref = p.grad.detach().float().cpu(), ref_named[n].grad
        rel = ((g - g_ref).norm() / (g_ref.norm() + 1e-12)).item()
        worst = max(worst, (rel, n))
        assert rel <= 2.9e-2, f"{n}: relative gradient error {rel:.3e}"     # measured worst 1.9e-2
    print(f"full-size training: {len(trainer.params)} gradient tensors, worst relative error {worst[0]:.3e} ({worst[1]})")
    for p in ref.parameters():
        p.grad = None


# --------------------------------------------------------------------------- round 2: f4 resize + crop on device, f3 checkpoint on GPU
@pytest.mark.parametrize("hw,channels_last", [((300, 400), True), ((500, 333), False), ((100, 130), True), ((37, 53), False),
                                              ((224, 224), True), ((640, 257), True)])
def test_device_resize_center_crop_is_bit_exact_vs_oracle(hw, channels_last):
    """kx_resize_crop_u8 (PIL's fixed-point bicubic, crop fused) against the oracle's restatement of CLIPImageProcessor's
    resize + centre crop, itself pinned bit for bit against transformers / PIL in tests/test_preprocess.py."""
    import numpy as np
    import kosmos_oracle as ko
    from kosmosx.preprocess import resize_center_crop_u8
    h, w = hw
    imgs = torch.from_numpy(np.random.default_rng(h + w).integers(0, 256, (3, h, w, 3), dtype=np.uint8))
    want = torch.stack([ko.clip_resize_center_crop_u8(im, 224, 224) for im in imgs])
    src = imgs if channels_last else imgs.permute(0, 3, 1, 2).contiguous()
    got = resize_center_crop_u8(src.cuda(), 224, 224)
    assert got.shape == (3, 224, 224, 3) and got.dtype == torch.uint8
    assert torch.equal(got.cpu(), want)


def test_forward_on_raw_pictures_of_another_size(tiny512_pair):
    """Kosmos.forward(normalize_images=True) and KosmosTokenizer.tokenize_images on uint8 pictures that are NOT the model's
    size: resize (shortest edge) + centre crop + rescale + normalise all on the device == the oracle's host pipeline."""
    import numpy as np
    import kosmos_oracle as ko
    from kosmosx import KosmosTokenizer
    ref, mine, oc = tiny512_pair

    class _Tok:
        pad_token_id = 1

        def convert_tokens_to_ids(self, t):
            return [oc.vocab - 2, oc.vocab - 1]

    tk = KosmosTokenizer(tokenizer=_Tok(), processor=object(), image_size=oc.image)
    raw = torch.from_numpy(np.random.default_rng(3).integers(0, 256, (2, 90, 141, 3), dtype=np.uint8))
    cropped = torch.stack([ko.clip_resize_center_crop_u8(im, oc.image, oc.image) for im in raw])
    pv_want = ko.clip_preprocess_u8(cropped, channels_last=True)
    pv = tk.tokenize_images(raw.cuda())
    assert torch.equal(pv.cpu(), pv_want)
    text, _ = ko.make_inputs(oc, 2, 20, seed=4)
    a = mine(text.cuda(), raw.cuda(), normalize_images=True).clone()
    b = mine(text.cuda(), pv).clone()
    assert torch.equal(a, b)
    with torch.no_grad():
        ref.set_emulation(True)
        want = ref(text, pv_want)
        ref.set_emulation(False)
    assert _err(a, want)[0] <= TOL_EMU_TINY


def test_load_checkpoint_file_then_forward_matches_oracle(tiny_cfgs, tmp_path):
    """§8(f)3 on the GPU: a file in the reference's `final_model.pt` layout (train.py:688-695: the unwrapped state_dict,
    here with the `module.` prefixes a DDP-wrapped save carries and one name of each tied pair dropped) ->
    load_checkpoint -> forward == the oracle holding those weights; and save_checkpoint round-trips."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos
    oc, kc = tiny_cfgs
    ref = ko.build(oc, seed=123)
    sd = {"module." + k: v.clone() for k, v in ref.state_dict().items()
          if k not in ("decoder.embed_tokens.weight", "embed_positions.weight", "decoder.output_projection.weight")}
    path = tmp_path / "final_model.pt"
    torch.save(sd, path)
    mine = Kosmos(config=kc).cuda()
    res = mine.load_checkpoint(str(path))
    assert not res.missing_keys and not res.unexpected_keys
    text, images = ko.make_inputs(oc, 2, 30, seed=8)
    with torch.no_grad():
        want32 = ref(text, images)
        ref.set_emulation(True)
        want16 = ref(text, images)
    got = mine(text.cuda(), images.cuda())
    assert _err(got, want16)[0] <= TOL_EMU_TINY and _err(got, want32)[0] <= TOL_F32_TINY
    assert _err(mine(text.cuda(), images.cuda(), precision="bf16x3"), want32)[0] <= TOL_STATED
    out = tmp_path / "resaved.pt"
    mine.save_checkpoint(str(out))
    again = torch.load(out, map_location="cpu", weights_only=True)
    assert set(again) == set(ref.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(again[k].float().cpu(), v.float()), k
