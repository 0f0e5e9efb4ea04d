"""Diagnostic (GPU box): per-parameter gradient error of our path and of torch-autocast against the fp32 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import ml_cvnets_b200 as m
from oracle import cvnets_oracle as O

width, res, B, seed = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0, int(sys.argv[2]) if len(sys.argv) > 2 else 128, int(sys.argv[3]) if len(sys.argv) > 3 else 2, 21
model = m.MobileViTv2(m.default_opts(width_multiplier=width))
P = O.seeded_fill_(O.mobilevit_v2_shapes(width), seed)
model.load_state_dict(P, strict=True)
model = model.cuda().train()
x = O.seeded_input((B, 3, res, res), 321).cuda()
y = torch.arange(B).cuda() * 37 % 1000
logits = model(x)
F.cross_entropy(logits.float(), y, label_smoothing=0.1).backward()
P32 = O.clone_params(P, device="cuda")
l32 = O.mobilevit_v2_forward(P32, x, width_multiplier=width)
F.cross_entropy(l32, y, label_smoothing=0.1).backward()
Pa = O.clone_params(P, device="cuda")
with torch.autocast("cuda", dtype=torch.bfloat16):
    la = O.mobilevit_v2_forward(Pa, x, width_multiplier=width)
    lossa = F.cross_entropy(la, y, label_smoothing=0.1)
lossa.backward()
def rel(a, b): return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))
print(f"logits: ours {rel(logits, l32):.4g} autocast {rel(la, l32):.4g}")
named = dict(model.named_parameters())
rows = []
for k, p in named.items():
    g32 = P32[k].grad
    rows.append((k, float(g32.norm()), rel(p.grad, g32), rel(Pa[k].grad, g32)))
print(f"{'param':70s} {'|g32|':>10s} {'ours':>9s} {'autocast':>9s}")
for k, n, eo, ea in rows:
    flag = " <<<" if eo > max(3 * ea, 0.05) else ""
    print(f"{k:70s} {n:10.4g} {eo:9.4f} {ea:9.4f}{flag}")
