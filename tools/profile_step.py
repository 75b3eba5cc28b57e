"""Run a few full-size Kosmos.forward steps (bench.py's workload, no timing legs) — the target of the
ncu captures under profiles/.  Usage: python tools/profile_step.py [--steps 2] [--batch 8] [--t-text 1984] [--train]
--train runs KosmosTrainer.step (configs[3]) instead of the forward."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kosmos-x_b200"))
from kosmosx import Kosmos, KosmosConfig, KosmosTrainer, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--t-text", type=int, default=1984)
ap.add_argument("--layers", type=int, default=24)
ap.add_argument("--vit-layers", type=int, default=24)
ap.add_argument("--train", action="store_true")
a = ap.parse_args()
torch.manual_seed(0)
model = Kosmos(config=KosmosConfig(max_positions=2050, layers=a.layers, vit_layers=a.vit_layers), device="cuda")
g = torch.Generator().manual_seed(1)
text = torch.randint(0, 32002, (a.batch, a.t_text), generator=g).cuda()
img = torch.randn(a.batch, 3, 224, 224, generator=g).cuda()
if a.train:
    trainer = KosmosTrainer(model, lr=1e-5, dropout=float(os.environ.get("KX_PROFILE_DROPOUT", "0.1")), attention_dropout=float(os.environ.get("KX_PROFILE_DROPOUT", "0.1")))
    model._pack_vision()
    torch.cuda.synchronize()
    print("staging launches:", ops.launch_count(), flush=True)
    for i in range(a.steps):
        n0 = ops.launch_count()
        loss = trainer.step(text, img)
        torch.cuda.synchronize()
        print(f"train step {i}: {ops.launch_count() - n0} launches, loss {float(loss):.4f}", flush=True)
    sys.exit(0)
model._pack_vision(); model.decoder._pack()
torch.cuda.synchronize()
print("staging launches:", ops.launch_count(), flush=True)
for i in range(a.steps):
    n0 = ops.launch_count()
    out = model(text, img)
    torch.cuda.synchronize()
    print(f"step {i}: {ops.launch_count() - n0} launches, logits {tuple(out.shape)}", flush=True)
    del out
