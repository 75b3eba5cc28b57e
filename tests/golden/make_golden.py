"""Generates tests/golden/tiny_golden.pt from the oracle (seeded tiny configuration).

The reference holds no golden vectors for this path ("parity unpinned"), so the
fixtures are minted here from the restatement oracle after it has been pinned against the
installed HF CLIP tower and HF Kosmos-2 text block (tests/test_oracle.py).  They guard the oracle
against drift and give the GPU parity tests a fixed target that does not depend on the oracle
code being importable.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import kosmos_oracle as ko  # noqa: E402


def main():
    torch.set_num_threads(4)
    cfg = ko.OracleConfig.tiny()
    model = ko.build(cfg, seed=0)
    cases = {}
    for name, (B, t_text) in {"b2_t10": (2, 10), "b1_t50": (1, 50), "b3_t130": (3, 130)}.items():
        text, images = ko.make_inputs(cfg, B, t_text, seed=1)
        st = model.stages(text, images)
        model.set_emulation(True)
        emu = model(text, images)
        model.set_emulation(False)
        step = 1 if B * t_text <= 20 else 8              # keep the fixture small: every 8th vocab column
        cases[name] = dict(B=B, t_text=t_text, col_step=step, vit=st["vit"][:, ::4].half(),
                           perceive=st["perceive"].half(), x0=st["x0"][:, ::2].half(),
                           logits=st["logits"][..., ::step].clone(), logits_emu_bf16=emu[..., ::step].clone())
    # the out-of-place reading of torchscale's forward_embedding (one positional embedding per text row,
    # OracleConfig.alias_embed_positions = False): same weights, one case
    model.cfg.alias_embed_positions = False
    text, images = ko.make_inputs(cfg, 1, 50, seed=1)
    st = model.stages(text, images)
    noalias = dict(B=1, t_text=50, col_step=8, x0=st["x0"][:, ::2].half(), logits=st["logits"][..., ::8].clone())
    model.cfg.alias_embed_positions = True
    torch.save(dict(cfg=cfg.__dict__, seed_weights=0, seed_inputs=1, cases=cases, noalias_b1_t50=noalias),
               os.path.join(HERE, "tiny_golden.pt"))
    print("wrote", os.path.join(HERE, "tiny_golden.pt"), {k: tuple(v["logits"].shape) for k, v in cases.items()})


def multi_image():
    """BASELINE.json configs[4] in small: m images per sequence spliced in front of given text tokens
    (oracle extension of model.py:239-241) -> tiny_golden_multi.pt."""
    torch.set_num_threads(4)
    cfg = ko.OracleConfig.tiny(max_positions=512)         # 4 images x 64 rows + text needs more than tiny's 256 positions
    model = ko.build(cfg, seed=0)
    cases = {}
    for name, (B, t_text, pos) in {"m4": (2, 30, [2, 9, 9, 30]), "m2_edges": (3, 5, [0, 5]), "m3": (1, 60, [1, 20, 41])}.items():
        text, images = ko.make_inputs(cfg, B, t_text, seed=7, n_images=len(pos))
        st = model.stages(text, images, pos)
        model.set_emulation(True)
        emu = model(text, images, image_positions=pos)
        model.set_emulation(False)
        cases[name] = dict(B=B, t_text=t_text, positions=pos, col_step=8, x0=st["x0"][:, ::2].half(),
                           logits=st["logits"][..., ::8].clone(), logits_emu_bf16=emu[..., ::8].clone())
    torch.save(dict(cfg=cfg.__dict__, seed_weights=0, seed_inputs=7, cases=cases), os.path.join(HERE, "tiny_golden_multi.pt"))
    print("wrote tiny_golden_multi.pt", {k: tuple(v["logits"].shape) for k, v in cases.items()})


if __name__ == "__main__":
    if "multi" in sys.argv[1:]:
        multi_image()
    else:
        main()
