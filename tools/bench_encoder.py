"""Forward+backward time of ONE ViT-B/16 TransformerEncoder layer ([256, 197, 768], 12 heads, f = 3072, GELU) and of
the CLIP text geometry ([256, 77, 512], 8 heads, f = 2048, causal mask), CUDA-event timed, with the algorithmic FLOPs of SURVEY.md 8d."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml_cvnets_b200 as m

def run(name, N, S, C, F_, H, act, causal):
    opts = m.default_opts(**{"model.activation.name": act})
    enc = m.TransformerEncoder(opts, C, F_, num_heads=H).cuda().train()
    x = torch.randn(N, S, C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    gy = torch.randn(N, S, C, device="cuda", dtype=torch.bfloat16)
    mask = torch.full((S, S), float("-inf"), device="cuda").triu(1)[None].repeat(N, 1, 1).contiguous() if causal else None
    def step():
        y = enc(x, attn_mask=mask)
        y.backward(gy)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    M = N * S
    macs = M * (3 * C * C + C * C + 2 * C * F_) + N * H * 2 * S * S * (C // H)
    tf = 3 * 2 * macs / (ms * 1e-3) / 1e12
    print(f"{name}: {ms:.3f} ms / layer fwd+bwd (eager launches), {tf:.1f} TFLOP/s algorithmic, {N / (ms * 1e-3) / 12:.0f} img/s if 12 such layers were the whole model")

run("ViT-B/16 encoder layer [256,197,768] gelu", 256, 197, 768, 3072, 12, "gelu", False)
run("CLIP text encoder layer [256,77,512] gelu causal", 256, 77, 512, 2048, 8, "gelu", True)
run("MobileViT-v1-like [512,256,64] swish", 512, 256, 64, 128, 4, "swish", False)
