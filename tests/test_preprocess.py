"""Host preprocessing (reference kosmosx/model.py:23-129), CPU side: the oracle's restatement of
CLIPImageProcessor's rescale + normalise pinned against the installed transformers, and ``KosmosTokenizer``'s
token / mask layout with injected (offline) tokenizer and processor objects.  The device kernels are checked
against the same oracle in tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

os.environ.setdefault("HF_HUB_OFFLINE", "1")


def _all_values(channels_last):
    """Every (uint8 value, channel) pair at least once, plus noise: (2, 3, 16, 16) or (2, 16, 16, 3)."""
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (2, 3, 16, 16), dtype=torch.uint8, generator=g)
    img[0] = torch.arange(256, dtype=torch.uint8).view(1, 16, 16)
    return img.permute(0, 2, 3, 1).contiguous() if channels_last else img


@pytest.mark.parametrize("channels_last", [False, True])
def test_oracle_preprocess_is_hf_rescale_then_normalize_bit_for_bit(channels_last):
    """transformers.image_transforms.rescale / normalize are the functions the reference's pinned slow CLIP
    processor calls; they are importable here, so this part of the oracle IS pinned."""
    import kosmos_oracle as ko
    from transformers import image_transforms as it
    from transformers.image_utils import ChannelDimension
    img = _all_values(channels_last)
    fmt = ChannelDimension.LAST if channels_last else ChannelDimension.FIRST
    want = []
    for a in img.numpy():
        r = it.rescale(a, 1 / 255, input_data_format=fmt)
        n = it.normalize(r, ko.OPENAI_CLIP_MEAN, ko.OPENAI_CLIP_STD, input_data_format=fmt)
        want.append(n.transpose(2, 0, 1) if channels_last else n)
    want = torch.from_numpy(np.stack(want))
    got = ko.clip_preprocess_u8(img, channels_last)
    assert got.dtype == torch.float32 and got.shape == (2, 3, 16, 16)
    assert torch.equal(got, want)


def test_oracle_preprocess_matches_the_installed_clip_image_processor():
    """End to end through ``CLIPImageProcessor`` (default = the OpenAI CLIP statistics of the laion checkpoint) at the
    model's size, where resize and centre crop are identities: <= 2 fp32 ulp (the installed version fuses the two
    steps), and identical once rounded to the bf16 the patch GEMM consumes."""
    import kosmos_oracle as ko
    from transformers import CLIPImageProcessor
    proc = CLIPImageProcessor()
    assert tuple(proc.image_mean) == ko.OPENAI_CLIP_MEAN and tuple(proc.image_std) == ko.OPENAI_CLIP_STD
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (2, 224, 224, 3), dtype=torch.uint8, generator=g)
    img[0, :16, :16] = torch.arange(256, dtype=torch.uint8).view(16, 16, 1)
    want = proc(images=list(img.numpy()), return_tensors="pt").pixel_values
    got = ko.clip_preprocess_u8(img, channels_last=True)
    assert (got - want).abs().max().item() <= 5e-7
    assert torch.equal(got.bfloat16(), want.bfloat16())


class _Enc:
    def __init__(self, ids):
        self.input_ids = ids


class _WordTokenizer:
    """Offline stand-in with the HF call convention: ``<s> words </s>``, right-padded with ``<pad>``."""

    def __init__(self):
        self.vocab = {"<pad>": 1, "<s>": 0, "</s>": 2, "<image>": 900, "</image>": 901}
        self.pad_token_id = 1

    def convert_tokens_to_ids(self, toks):
        return [self.vocab[t] for t in toks]

    def __call__(self, texts, return_tensors=None, padding=False, truncation=False):
        assert return_tensors == "pt" and padding and truncation        # the reference's arguments, model.py:69-71
        texts = [texts] if isinstance(texts, str) else texts
        rows = [[0] + [10 + (sum(map(ord, w)) % 800) for w in t.split()] + [2] for t in texts]
        n = max(map(len, rows))
        return _Enc(torch.tensor([r + [1] * (n - len(r)) for r in rows], dtype=torch.long))


def _tokenizer(**kw):
    from kosmosx import KosmosTokenizer
    from transformers import CLIPImageProcessor
    return KosmosTokenizer(tokenizer=_WordTokenizer(), processor=CLIPImageProcessor(), **kw)


def test_tokenizer_text_layout_matches_the_reference():
    import kosmos_oracle as ko
    tk = _tokenizer()
    assert (tk.im_idx, tk.im_end_idx) == (900, 901)
    texts = ["a photo of a cat", "two dogs"]
    both, only = tk.tokenize_texts(texts)
    ids = tk.tokenizer(texts, return_tensors="pt", padding=True, truncation=True).input_ids
    want_both, want_only = ko.tokenize_texts(ids, 900, 901)
    assert torch.equal(both, want_both) and torch.equal(only, want_only)
    assert both.shape == (2, 9) and both[:, 0].tolist() == [0, 0] and both[:, 1].tolist() == [900, 900] \
        and both[:, 2].tolist() == [901, 901]
    assert torch.equal(both[:, 3:], ids[:, 1:])


def test_tokenizer_sample_dict_matches_the_reference():
    import kosmos_oracle as ko
    tk = _tokenizer()
    g = torch.Generator().manual_seed(9)
    imgs = torch.randint(0, 256, (2, 224, 224, 3), dtype=torch.uint8, generator=g)
    out = tk.tokenize({"target_text": ["a photo of a cat", "two dogs"], "image": list(imgs.numpy())})
    assert set(out) == {"text_tokens", "images", "labels", "attention_mask"}
    assert out["text_tokens"].shape == (2, 9) and out["labels"].shape == (2, 7)
    assert out["images"].shape == (2, 3, 224, 224) and out["images"].dtype == torch.float32
    assert (out["images"] - ko.clip_preprocess_u8(imgs, channels_last=True)).abs().max().item() <= 5e-7
    # the reference's mask: 64 ones in FRONT, then tokens != pad (model.py:113-120)
    mask = out["attention_mask"]
    assert torch.equal(mask, ko.tokenize_attention_mask(out["text_tokens"], 1))
    assert mask.shape == (2, 64 + 9) and mask[:, :64].all() and mask[0].all() and mask[1, -3:].tolist() == [0, 0, 0]


def test_tokenizer_without_injection_fails_like_the_reference_offline(caplog):
    """No network: the hub loads of model.py:36-46 raise; the error is logged and re-raised (model.py:47-49)."""
    from kosmosx import KosmosTokenizer
    with caplog.at_level("ERROR", logger="kosmosx"):
        with pytest.raises(Exception):
            KosmosTokenizer()
    assert "Failed to initialize KosmosTokenizer" in caplog.text


def test_device_preprocessing_has_no_cpu_path():
    """uint8 pixels on the host go to the processor (reference behaviour); the device kernel is never emulated."""
    from kosmosx import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_parity.py")
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.clip_normalize_u8(torch.zeros(1, 3, 224, 224, dtype=torch.uint8), image=224)


def test_tokenizer_with_a_real_hf_fast_tokenizer_built_offline():
    """The same layout through a genuine ``PreTrainedTokenizerFast`` (word-level vocabulary built in memory, special tokens
    declared the way the reference declares them, model.py:39-46): the HF call convention the class relies on is the real one."""
    tokenizers = pytest.importorskip("tokenizers")
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import Whitespace
    from tokenizers.processors import TemplateProcessing
    from transformers import CLIPImageProcessor, PreTrainedTokenizerFast
    from kosmosx import KosmosTokenizer
    words = ["a", "photo", "of", "cat", "two", "dogs"]
    vocab = {"<s>": 0, "<pad>": 1, "<eos>": 2, "<unk>": 3, **{w: 4 + i for i, w in enumerate(words)}}
    tok = tokenizers.Tokenizer(WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = Whitespace()
    tok.post_processor = TemplateProcessing(single="<s> $A <eos>", special_tokens=[("<s>", 0), ("<eos>", 2)])
    hf = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<s>", eos_token="<eos>", pad_token="<pad>", unk_token="<unk>",
                                 additional_special_tokens=["<image>", "</image>"], model_max_length=8192)
    tk = KosmosTokenizer(tokenizer=hf, processor=CLIPImageProcessor())
    assert tk.im_idx == hf.convert_tokens_to_ids("<image>") and tk.im_end_idx == hf.convert_tokens_to_ids("</image>")
    assert tk.im_idx >= len(vocab) and tk.im_end_idx == tk.im_idx + 1          # appended after the base vocabulary
    both, only = tk.tokenize_texts(["a photo of a cat", "two dogs"])
    assert only.tolist() == [[0, 4, 5, 6, 4, 7, 2], [0, 8, 9, 2, 1, 1, 1]]
    assert both.tolist() == [[0, tk.im_idx, tk.im_end_idx, 4, 5, 6, 4, 7, 2], [0, tk.im_idx, tk.im_end_idx, 8, 9, 2, 1, 1, 1]]
    g = torch.Generator().manual_seed(2)
    imgs = torch.randint(0, 256, (2, 224, 224, 3), dtype=torch.uint8, generator=g)
    out = tk.tokenize({"target_text": ["a photo of a cat", "two dogs"], "image": list(imgs.numpy())})
    assert out["attention_mask"].shape == (2, 64 + 9) and out["attention_mask"][1, -3:].tolist() == [0, 0, 0]
    assert out["images"].shape == (2, 3, 224, 224)


# --------------------------------------------------------------------------- resize + centre crop (round 2)
_SIZES = [(300, 400), (224, 224), (500, 333), (100, 130), (640, 480), (37, 53), (224, 300), (1000, 257)]


@pytest.mark.parametrize("hw", _SIZES)
def test_oracle_resize_crop_is_hf_resize_then_center_crop_bit_for_bit(hw):
    """``clip_resize_center_crop_u8`` against the functions the reference's pinned (4.35, PIL-based) CLIPImageProcessor calls:
    get_resize_output_image_size(shortest_edge=224), image_transforms.resize(BICUBIC) -> PIL.Image.resize, center_crop.
    Up- and down-scaling, square, extreme aspect ratios: identical uint8 pixels."""
    import kosmos_oracle as ko
    from transformers import image_transforms as it
    from transformers.image_utils import ChannelDimension, PILImageResampling
    h, w = hw
    img = np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    size = it.get_resize_output_image_size(img, size=224, default_to_square=False, input_data_format=ChannelDimension.LAST)
    r = it.resize(img, size=size, resample=PILImageResampling.BICUBIC, input_data_format=ChannelDimension.LAST)
    want = it.center_crop(r, (224, 224), input_data_format=ChannelDimension.LAST)
    got = ko.clip_resize_center_crop_u8(img, 224, 224)
    assert want.dtype == np.uint8 and got.shape == (224, 224, 3)
    assert np.array_equal(got.numpy(), want)


@pytest.mark.parametrize("hw", _SIZES)
def test_product_resize_tables_reproduce_the_oracle(hw):
    """kosmosx/preprocess.py builds the fixed-point taps the device kernel consumes; applying them in plain integer numpy
    (what kx_resize_crop_u8 does on the GPU) must give the oracle's pixels — geometry, taps and crop offsets agree."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kosmos-x_b200"))
    import kosmos_oracle as ko
    from kosmosx import preprocess as pp
    h, w = hw
    img = np.random.default_rng(7 + h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    new_h, new_w = pp.resize_output_size(h, w, 224)
    top, left = (new_h - 224) // 2, (new_w - 224) // 2
    kx, bx = pp.bicubic_taps(w, new_w)
    ky, by = pp.bicubic_taps(h, new_h)
    kx, bx, ky, by = kx[left:left + 224], bx[left:left + 224], ky[top:top + 224], by[top:top + 224]
    src = img.astype(np.int64)
    tmp = np.empty((h, 224, 3), dtype=np.uint8)
    for xx in range(224):
        lo, n = bx[xx]
        acc = np.tensordot(src[:, lo:lo + n], kx[xx, :n].astype(np.int64), axes=(1, 0)) + (1 << 21)
        tmp[:, xx] = np.clip(acc >> 22, 0, 255)
    out = np.empty((224, 224, 3), dtype=np.uint8)
    t64 = tmp.astype(np.int64)
    for yy in range(224):
        lo, n = by[yy]
        acc = np.tensordot(ky[yy, :n].astype(np.int64), t64[lo:lo + n], axes=(0, 0)) + (1 << 21)
        out[yy] = np.clip(acc >> 22, 0, 255)
    assert np.array_equal(out, ko.clip_resize_center_crop_u8(img, 224, 224).numpy())
