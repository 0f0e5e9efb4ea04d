"""Diagnostics: workspace-mode gradients vs the plain autograd path, per parameter (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ml_cvnets_b200 as m  # noqa: E402
from oracle import cvnets_oracle as O  # noqa: E402


def small(width=0.5, seed=11):
    model = m.MobileViTv2(m.default_opts(width_multiplier=width))
    model.load_state_dict(O.seeded_fill_(O.mobilevit_v2_shapes(width), seed), strict=True)
    return model.cuda().train()


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


B, res = 8, 64
x = O.seeded_input((B, 3, res, res), 5).cuda()
y = torch.arange(B, device="cuda") % 1000
ref = small()
logits = ref(x)
scale = 65536.0
loss_ref = m.cross_entropy(logits, y, label_smoothing=0.1)
(loss_ref * scale).backward()
ref2 = small()
(m.cross_entropy(ref2(x), y, label_smoothing=0.1) * scale).backward()
print("run-to-run (plain path) max rel:", max(rel(p.grad, q.grad) for p, q in zip(ref2.parameters(), ref.parameters())))
model = small()
ts = m.TrainStep(model, lr=0.0, weight_decay=0.0)
for it in range(3):
    loss = ts.step(x, y)
    torch.cuda.synchronize()
    bad = [(k, rel(p.grad, q.grad), float(q.grad.norm())) for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters())]
    nb = [t for t in bad if t[1] > 2e-3]
    print(f"step {it}: loss {float(loss):.6f} vs {float(loss_ref):.6f}; {len(nb)}/{len(bad)} parameters off")
    for t in nb[:12]:
        print("   ", t)
