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


if __name__ == "__main__":
    main()
