"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol declared
in include/kosmosx_b200.h, entry points fail loudly without a GPU (no CPU fallback), and the
Python surface mirrors the reference's names / state_dict layout."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kosmosx_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from kosmosx import _abi
    names = _declared_symbols()
    assert len(names) >= 12
    raw = ctypes.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in kosmosx_b200.h but not exported"
    assert set(names) == set(_abi.SIGNATURES), "ctypes SIGNATURES and header disagree"
    assert _abi.lib.kx_abi_version() == int(re.search(r"#define KX_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_gemm_args_struct_matches_header():
    from kosmosx import _abi
    src = open(HEADER).read()
    body = re.search(r"typedef struct kx_gemm_args \{(.*?)\} kx_gemm_args;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    assert fields == [f[0] for f in _abi.GemmArgs._fields_]


def test_decode_linear_args_struct_matches_header():
    from kosmosx import _abi
    src = open(HEADER).read()
    body = re.search(r"typedef struct kx_decode_linear_args \{(.*?)\} kx_decode_linear_args;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    assert fields == [f[0] for f in _abi.DecodeLinearArgs._fields_]


def test_decode_entry_points_validate_arguments():
    """Incremental-decoding entry points: bad shapes are rejected before any launch; without a GPU a
    well-formed call reports KX_ERR_NO_DEVICE (no CPU path)."""
    from kosmosx import _abi
    buf = (ctypes.c_char * 65536)()
    addr = (ctypes.addressof(buf) + 255) & ~255
    g = _abi.DecodeLinearArgs()
    g.mode, g.out, g.ld_out, g.out_f32 = _abi.KX_DEC_PLAIN, addr, 64, 1
    assert _abi.lib.kx_decode_linear(addr, 48, 4, addr, 48, 64, 48, g, None) == -1            # K % 32 != 0
    assert "K % 32" in _abi.last_error()
    assert _abi.lib.kx_decode_linear(addr, 64, 33, addr, 64, 64, 64, g, None) == -1           # batch > KX_DECODE_MAX_BATCH
    g.mode = _abi.KX_DEC_QKV
    assert _abi.lib.kx_decode_linear(addr, 64, 4, addr, 64, 192, 64, g, None) == -1           # caches / tables missing
    assert "KX_DEC_QKV" in _abi.last_error()
    g.mode = 7
    assert _abi.lib.kx_decode_linear(addr, 64, 4, addr, 64, 64, 64, g, None) == -1
    assert _abi.lib.kx_kv_cache_store(addr, 192, 1, 9, 64, addr, addr, 8, None) == -1         # prompt longer than the cache
    assert _abi.lib.kx_decode_attn(addr, 64, addr, addr, 0, 1, 1, addr, 0.125, addr, addr, addr, 64, None) == -1
    assert _abi.lib.kx_decode_attn_scratch_bytes(2, 4, 300) == 2 * 4 * 3 * 66 * 4
    if not torch.cuda.is_available():
        g.mode = _abi.KX_DEC_PLAIN
        assert _abi.lib.kx_decode_linear(addr, 64, 4, addr, 64, 64, 64, g, None) == -2
        assert _abi.lib.kx_kv_cache_store(addr, 192, 1, 8, 64, addr, addr, 8, None) == -2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_entry_points_fail_loudly_without_gpu():
    from kosmosx import _abi
    assert _abi.lib.kx_device_check() == -2
    assert "no CPU fallback" in _abi.last_error()
    g = _abi.GemmArgs()
    g.M = g.N = g.K = 64
    buf = (ctypes.c_char * 65536)()
    addr = (ctypes.addressof(buf) + 255) & ~255
    g.out = addr
    g.ld_out = 64
    st = _abi.lib.kx_gemm_bf16(addr, 64, addr, 64, g, None)
    assert st == -2, _abi.last_error()


def test_argument_validation_happens_before_any_launch():
    from kosmosx import _abi
    g = _abi.GemmArgs()
    assert _abi.lib.kx_gemm_bf16(None, 0, None, 0, g, None) == -1
    assert "null" in _abi.last_error()
    buf = (ctypes.c_char * 4096)()
    addr = (ctypes.addressof(buf) + 255) & ~255
    g.M, g.N, g.K, g.out, g.ld_out = 8, 8, 8, addr, 8
    assert _abi.lib.kx_gemm_bf16(addr, 7, addr, 8, g, None) == -1          # row pitch not 16-byte aligned
    assert _abi.lib.kx_layernorm_fwd(addr, 0, 12, None, 0, 0, addr, addr, 1e-5, addr, 0, 16, 4, 12, 0, 0, 0, None) == -1
    assert "multiple of 8" in _abi.last_error()
    assert _abi.lib.kx_attn_fwd(addr, addr, addr, 8, addr, 8, 0, 1, 1, 1, 0.125, None, None) == -1
    assert _abi.lib.kx_rowstats_cast(addr, 12, addr, 16, addr, 4, 12, None) == -1      # n not a multiple of 8


def test_public_surface_matches_reference():
    import kosmosx
    import kosmosx.model as m
    assert set(kosmosx.__all__) >= {"KosmosTokenizer", "Kosmos", "KosmosLanguage"}     # reference __init__.py:4
    assert hasattr(m, "Decoder")                                                       # train.py:44 imports it
    import inspect
    sig = inspect.signature(m.Kosmos.forward)
    assert list(sig.parameters)[:3] == ["self", "text_tokens", "images"]
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY
               for n, p in inspect.signature(m.Kosmos.__init__).parameters.items() if n != "self")
    ks = inspect.signature(m.KosmosLanguage.__init__).parameters
    assert [n for n in ks][1:14] == ["vocab_size", "dim", "depth", "ffn_dim", "dropout", "multiway", "decoder_heads",
                                     "activation_fn", "subln", "alibi_pos_bias", "alibi_num_heads", "xpos_rel_pos",
                                     "max_rel_pos"]
    assert ks["vocab_size"].default == 64007 and ks["depth"].default == 24


def test_state_dict_layout_matches_oracle_and_appendix_b(tiny_cfgs):
    import kosmos_oracle as ko
    from kosmosx import Kosmos
    oc, kc = tiny_cfgs
    torch.manual_seed(0)
    mine, ref = Kosmos(config=kc), ko.KosmosOracle(oc)
    a, b = mine.state_dict(), ref.state_dict()
    assert set(a) == set(b)
    assert all(a[k].shape == b[k].shape for k in a)
    for k in ("clip_model.embeddings.class_embedding", "clip_model.pre_layrnorm.weight",
              "clip_model.encoder.layers.0.self_attn.q_proj.weight", "clip_model.post_layernorm.bias",
              "embed.weight", "decoder.embed_tokens.weight", "embed_positions.weight", "decoder.embed_positions.weight",
              "output_projection.weight", "decoder.output_projection.weight",
              "decoder.layers.1.self_attn.q_proj.A.weight", "decoder.layers.1.self_attn.q_proj.B.bias",
              "decoder.layers.0.self_attn.inner_attn_ln.A.weight", "decoder.layers.0.self_attn.xpos.scale",
              "decoder.layers.0.self_attn_layer_norm.B.weight", "decoder.layers.0.ffn.A.fc1.weight",
              "decoder.layers.0.ffn.B.ffn_layernorm.bias", "decoder.layers.0.final_layer_norm.A.bias",
              "decoder.layer_norm.weight", "perceive.latents", "perceive.media_pos_emb",
              "perceive.layers.0.0.norm_media.weight", "perceive.layers.1.0.to_kv.weight",
              "perceive.layers.0.1.0.weight", "perceive.layers.0.1.1.weight", "perceive.layers.0.1.3.weight",
              "perceive.norm.bias", "image_proj.weight"):
        assert k in a, k
    assert a["embed.weight"].data_ptr() == a["decoder.embed_tokens.weight"].data_ptr()
    mine.load_state_dict(b)                        # drop-in checkpoint load
    assert torch.equal(mine.state_dict()["decoder.layers.0.ffn.A.fc1.weight"], b["decoder.layers.0.ffn.A.fc1.weight"])


def test_hf_clip_checkpoint_loads_into_the_vision_tower(tiny_cfgs):
    """the reference takes its vision weights from ``CLIPModel.from_pretrained(...).vision_model``
    (model.py:154-156).  Offline there is no checkpoint, but the installed HF class defines the key layout: its
    state_dict must load into ``Kosmos().clip_model`` key for key (a non-persistent ``position_ids`` buffer aside)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from kosmosx import Kosmos
    oc, kc = tiny_cfgs
    hf = CLIPVisionModel(CLIPVisionConfig(hidden_size=oc.vit_dim, intermediate_size=oc.vit_mlp,
                                          num_hidden_layers=oc.vit_layers, num_attention_heads=oc.vit_heads,
                                          patch_size=oc.patch, image_size=oc.image, hidden_act="gelu")).vision_model
    mine = Kosmos(config=kc)
    sd = hf.state_dict()
    res = mine.clip_model.load_state_dict(sd, strict=False)
    assert not res.missing_keys, res.missing_keys
    assert set(res.unexpected_keys) <= {"embeddings.position_ids"}, res.unexpected_keys
    own = mine.clip_model.state_dict()
    assert all(torch.equal(own[k], v) for k, v in sd.items() if k in own)
    # and the whole-model layout: clip_model.* keys of Kosmos.state_dict() are exactly HF's
    full = {k[len("clip_model."):] for k in mine.state_dict() if k.startswith("clip_model.")}
    assert full == set(own)


def test_reference_error_behaviour(tiny_cfgs):
    from kosmosx import Kosmos, KosmosLanguage, KosmosTokenizer
    _, kc = tiny_cfgs
    model = Kosmos(config=kc)
    with pytest.raises(TypeError, match="must be instances of torch.Tensor"):       # model.py:222-227
        model([1, 2, 3], torch.zeros(1, 3, kc.image, kc.image))
    with pytest.raises(TypeError):
        model(torch.zeros(1, 5, dtype=torch.long), "img")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            model(torch.zeros(1, 5, dtype=torch.long), torch.zeros(1, 3, kc.image, kc.image))
    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    with pytest.raises(Exception):                  # hub loads of model.py:36-46 fail offline and re-raise (tests/test_preprocess.py)
        KosmosTokenizer()
    with pytest.raises(NotImplementedError):
        KosmosLanguage(vocab_size=100, dim=128, depth=1, ffn_dim=128, decoder_heads=2, activation_fn="swish")


def test_config_validation():
    from kosmosx import KosmosConfig
    with pytest.raises(ValueError):
        KosmosConfig(dim=2048, heads=16).validate()       # head_dim 128
    with pytest.raises(ValueError, match="multiple of patch"):
        KosmosConfig(image=225).validate()
    with pytest.raises(ValueError, match="multiple of patch"):
        KosmosConfig(image=42, patch=14).validate()       # 42 % 4 != 0: the patch pack reads 16-byte pixel vectors
    KosmosConfig().validate()


def test_full_size_parameter_counts():
    """SURVEY.md Appendix B counts, on the meta device (no memory)."""
    from kosmosx import Kosmos
    m = Kosmos(device="meta")
    n = lambda mod: sum(p.numel() for p in mod.parameters())
    assert n(m.clip_model) == 303_179_776
    assert n(m.perceive) == 21_314_560 + 0
    assert n(m.image_proj) == 2_097_152
    assert n(m.embed) == 65_540_096 and n(m.output_projection) == 65_540_096
    per_layer = sum(p.numel() for name, p in m.decoder.layers[0].named_parameters() if ".B." not in name)
    assert per_layer == 50_378_752


def test_reference_checkpoint_file_round_trip(tiny_cfgs, tmp_path):
    """the reference saves ``unwrapped_model.state_dict()`` as final_model.pt (train.py:688-695).  A file
    written from the oracle's state_dict (the reference's layout, Appendix B) loads through ``load_checkpoint``, also
    with wrapper prefixes, HF's legacy position_ids buffer, one name of a tied pair only, and another dtype."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos
    oc, kc = tiny_cfgs
    ref = ko.build(oc, seed=0)
    sd = ref.state_dict()
    path = tmp_path / "final_model.pt"
    torch.save(sd, path)
    mine = Kosmos(config=kc)
    res = mine.load_checkpoint(str(path))
    assert not res.missing_keys and not res.unexpected_keys
    own = mine.state_dict()
    assert all(torch.equal(own[k], v) for k, v in sd.items())
    # FSDP / DDP names, bf16 storage, position_ids, only one name of each tied pair
    messy = {"module._fsdp_wrapped_module." + k: v.to(torch.bfloat16) if v.is_floating_point() else v for k, v in sd.items()
             if not k.startswith(("decoder.embed_tokens.", "decoder.embed_positions.", "decoder.output_projection."))}
    messy["module._fsdp_wrapped_module.clip_model.embeddings.position_ids"] = torch.arange(oc.vit_tokens).unsqueeze(0)
    other = Kosmos(config=kc)
    res = other.load_checkpoint(messy)
    assert not res.missing_keys and not res.unexpected_keys
    got = other.state_dict()
    assert all(torch.equal(got[k], sd[k].to(torch.bfloat16).to(sd[k].dtype)) for k in sd if sd[k].is_floating_point())
    assert got["embed.weight"].data_ptr() == got["decoder.embed_tokens.weight"].data_ptr()
    # save_checkpoint writes the same layout back
    out = tmp_path / "again.pt"
    mine.save_checkpoint(str(out))
    again = torch.load(out, weights_only=True)
    assert set(again) == set(sd) and all(torch.equal(again[k], sd[k]) for k in sd)
    with pytest.raises(RuntimeError):                                   # strict by default: a missing tensor is an error
        Kosmos(config=kc).load_checkpoint({k: v for k, v in sd.items() if "image_proj" not in k})


def test_plain_torchscale_and_resized_position_checkpoints(tiny_cfgs):
    """A non-multiway torchscale decoder checkpoint (``q_proj.weight`` rather than ``q_proj.A.weight``) lands on the live
    ``.A`` branches; a positional table of another length needs ``resize_positions``."""
    import kosmos_oracle as ko
    from kosmosx import Kosmos, KosmosConfig
    oc, kc = tiny_cfgs
    sd = ko.build(oc, seed=0).state_dict()
    plain = {k.replace(".A.", "."): v for k, v in sd.items() if ".B." not in k}
    assert "decoder.layers.0.ffn.fc1.weight" in plain and "decoder.layers.1.self_attn.q_proj.bias" in plain
    mine = Kosmos(config=kc)
    res = mine.load_checkpoint(plain)
    assert not res.unexpected_keys and all(".B." in k for k in res.missing_keys)
    own = mine.state_dict()
    assert all(torch.equal(own[k], v) for k, v in sd.items() if ".B." not in k)
    with pytest.raises(RuntimeError, match="missing keys"):           # a plain checkpoint still has to be complete
        Kosmos(config=kc).load_checkpoint({k: v for k, v in plain.items() if "image_proj" not in k})
    longer = KosmosConfig(**{**{k: getattr(kc, k) for k in KosmosConfig.__dataclass_fields__}, "max_positions": kc.max_positions + 2})
    big = Kosmos(config=longer)
    with pytest.raises(RuntimeError, match="size mismatch"):
        big.load_checkpoint(sd)
    keep = big.state_dict()["embed_positions.weight"][-2:].clone()
    big.load_checkpoint(sd, resize_positions=True)
    pos = big.state_dict()["embed_positions.weight"]
    assert torch.equal(pos[:kc.max_positions], sd["embed_positions.weight"]) and torch.equal(pos[-2:], keep)


def test_autograd_bridge_mechanics_with_a_stub_trainer():
    """The autograd node behind ``Kosmos.forward`` in train mode (kosmosx/train.py: _KosmosAutograd), exercised on the CPU
    with a stand-in for the trainer (a one-matrix model whose hand-written backward fills a flat gradient buffer): the
    output requires grad, a PyTorch loss drives the hand-written backward, ``param.grad`` is re-attached to the flat buffer
    after ``zero_grad(set_to_none=True)``, and a backward whose activations were re-used fails loudly.  The real trainer
    is checked on the GPU (tests/test_gpu_train.py::test_pytorch_training_loop_through_autograd)."""
    from kosmosx.train import _KosmosAutograd

    class Stub:
        def __init__(self):
            self.w = torch.nn.Parameter(torch.randn(5, 3))
            self.G = torch.zeros(15)
            self.buf = torch.zeros(8, 5)
            self._fw_serial = 0
            self.params = [self.w]
            self.accumulate_seen = []

        def _g(self, p):
            return self.G.view(5, 3)

        def _forward(self, tokens, images, rows):
            self.x = images.reshape(8, 3)
            self.buf.copy_(self.x @ self.w.detach().t())
            return dict(logits=self.buf, B=2, T=4)

        def _backward(self, fw, tokens, rows, dlogits_in=None, accumulate=False):
            self.accumulate_seen.append(accumulate)
            g = dlogits_in.reshape(8, 5).t() @ self.x
            self.G.copy_((self.G.view(5, 3) + g if accumulate else g).reshape(-1))

        def attach_grads(self):
            for p in self.params:
                if p.grad is None or p.grad.data_ptr() != self._g(p).data_ptr():
                    p.grad = self._g(p)

    tr = Stub()
    images, tokens = torch.randn(2, 4, 3), torch.zeros(2, 4, dtype=torch.long)
    target = torch.randint(0, 5, (8,))
    opt = torch.optim.SGD([tr.w], lr=0.1)
    opt.zero_grad()                                              # .grad is None from here
    out = _KosmosAutograd.apply(tr.w, tr, tokens, images, (2,))
    assert out.requires_grad and out.shape == (2, 4, 5)
    torch.nn.functional.cross_entropy(out.reshape(-1, 5), target).backward()
    w2 = tr.w.detach().clone().requires_grad_(True)
    torch.nn.functional.cross_entropy(images.reshape(8, 3) @ w2.t(), target).backward()
    assert tr.w.grad.data_ptr() == tr.G.data_ptr() and torch.allclose(tr.w.grad, w2.grad, atol=1e-6)
    assert tr.accumulate_seen == [False]
    out = _KosmosAutograd.apply(tr.w, tr, tokens, images, (2,))     # no zero_grad in between: accumulate
    torch.nn.functional.cross_entropy(out.reshape(-1, 5), target).backward()
    assert tr.accumulate_seen == [False, True] and torch.allclose(tr.w.grad, 2 * w2.grad, atol=1e-6)
    before = tr.w.detach().clone()
    opt.step()
    assert not torch.equal(before, tr.w.detach())
    stale = _KosmosAutograd.apply(tr.w, tr, tokens, images, (2,)).sum()
    _KosmosAutograd.apply(tr.w, tr, tokens, images, (2,))
    with pytest.raises(RuntimeError, match="activations of this forward are gone"):
        stale.backward()
