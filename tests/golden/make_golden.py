"""Generate the golden fixtures in this directory FROM THE REAL REFERENCE (apple/ml-cvnets @ /root/reference).

Runs only in the build container (the reference does not exist on the GPU box).  It imports the reference's own
``nn.Module``s, loads the deterministic parameters of ``oracle.cvnets_oracle.seeded_fill_`` into them (which also
asserts the ``state_dict`` key/shape contract of SURVEY.md App. B), runs forward+backward in fp32 on CPU and stores
the results.  ``tests/test_oracle_golden.py`` pins the oracle to these files; the ``-m gpu`` parity tests then compare
the CUDA path with the oracle and with these files.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
os.chdir(REF)  # registries glob relative to the library root (utils/import_utils.py:26-41)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from cvnets import get_model, modeling_arguments  # noqa: E402
from cvnets.layers import ConvLayer2d, LinearSelfAttention  # noqa: E402
from cvnets.modules import InvertedResidual, MobileViTBlockv2  # noqa: E402
from cvnets.modules.transformer import LinearAttnFFN  # noqa: E402
from oracle import cvnets_oracle as O  # noqa: E402


def make_opts(width=1.0):
    opts = modeling_arguments(argparse.ArgumentParser()).parse_args([])
    setattr(opts, "dataset.category", "classification")
    setattr(opts, "model.classification.name", "mobilevit_v2")
    setattr(opts, "model.classification.mitv2.width_multiplier", width)
    setattr(opts, "model.activation.name", "swish")
    setattr(opts, "model.classification.activation.name", "swish")
    return opts


def load_seeded(module, shapes_P, seed, prefix=""):
    """Check the state_dict contract and load seeded params."""
    sd = module.state_dict()
    want = {prefix + k if prefix else k: v for k, v in shapes_P.items()}
    assert set(sd.keys()) == set(want.keys()), (sorted(set(sd) ^ set(want)))
    for k in sd:
        assert tuple(sd[k].shape) == tuple(want[k].shape), (k, sd[k].shape, want[k].shape)
    O.seeded_fill_(shapes_P, seed)
    module.load_state_dict({k: v.clone() for k, v in want.items()}, strict=True)
    return shapes_P


def run_module(module, x, gy_seed):
    module.train()
    x = x.clone().requires_grad_(True)
    y = module(x)
    gy = O.seeded_input(tuple(y.shape), gy_seed)
    y.backward(gy)
    out = {"x": x.detach().clone(), "y": y.detach().clone(), "gy": gy, "gx": x.grad.clone()}
    out["grads"] = {k: p.grad.clone() for k, p in module.named_parameters()}
    out["buffers"] = {k: b.detach().clone() for k, b in module.named_buffers()}
    return out


def strip(prefix, d):
    return {k[len(prefix):]: v for k, v in d.items()}


def main():
    torch.manual_seed(0)
    opts = make_opts(1.0)
    fixtures = {}

    # ---- stem ConvLayer2d 3->16 k3 s2 + BN + SiLU ------------------------------------------------------
    P = {}
    O._conv_bn(P, "m", 3, 16, 3)
    m = ConvLayer2d(opts, 3, 16, 3, stride=2, use_norm=True, use_act=True)
    load_seeded(m, strip("m.", P), 11)
    fixtures["stem"] = dict(cfg=dict(cin=3, cout=16), seed=11, **run_module(m, O.seeded_input((2, 3, 16, 16), 101), 201))

    # ---- InvertedResidual, stride 1 with residual, stride 2 without --------------------------------------
    for name, (cin, cout, stride, seed) in {"ir_s1_res": (16, 16, 1, 12), "ir_s2": (16, 32, 2, 13)}.items():
        P = {}
        O.inverted_residual_shapes(P, "m", cin, cout, 2)
        m = InvertedResidual(opts, in_channels=cin, out_channels=cout, stride=stride, expand_ratio=2)
        load_seeded(m, strip("m.", P), seed)
        fixtures[name] = dict(cfg=dict(cin=cin, cout=cout, stride=stride, expand_ratio=2), seed=seed,
                              **run_module(m, O.seeded_input((2, cin, 8, 8), 100 + seed), 200 + seed))

    # ---- LinearSelfAttention and LinearAttnFFN on [B, d, P, N] ---------------------------------------------
    P = {}
    O._conv_bn(P, "m.qkv_proj", 16, 33, 1, norm=False, bias=True)
    O._conv_bn(P, "m.out_proj", 16, 16, 1, norm=False, bias=True)
    m = LinearSelfAttention(opts, embed_dim=16, attn_dropout=0.0, bias=True)
    load_seeded(m, strip("m.", P), 14)
    fixtures["lsa"] = dict(cfg=dict(d=16), seed=14, **run_module(m, O.seeded_input((2, 16, 4, 9), 114), 214))

    P = {}
    O.linear_attn_ffn_shapes(P, "m", 16, 32)
    m = LinearAttnFFN(opts, embed_dim=16, ffn_latent_dim=32, attn_dropout=0.0, dropout=0.0, ffn_dropout=0.0)
    load_seeded(m, strip("m.", P), 15)
    fixtures["laffn"] = dict(cfg=dict(d=16, ffn=32), seed=15, **run_module(m, O.seeded_input((2, 16, 4, 9), 115), 215))

    # ---- MobileViTBlockv2 ------------------------------------------------------------------------------------
    P = {}
    O.mobilevit_block_v2_shapes(P, "m", 32, 16, 2, 2.0)
    m = MobileViTBlockv2(opts, in_channels=32, attn_unit_dim=16, ffn_multiplier=2.0, n_attn_blocks=2, patch_h=2, patch_w=2)
    load_seeded(m, strip("m.", P), 16)
    fixtures["mvit_v2"] = dict(cfg=dict(c=32, d=16, n_attn_blocks=2), seed=16,
                               **run_module(m, O.seeded_input((2, 32, 8, 8), 116), 216))

    # unfold index map probe: patches[b,c,p,n] = x[b,c,(n//nw)*2 + p//2, (n%nw)*2 + p%2]
    xm = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).reshape(2, 3, 4, 6)
    patches, _ = m.unfolding_pytorch(xm)
    fixtures["unfold_probe"] = dict(x=xm, patches=patches)

    torch.save(fixtures, os.path.join(HERE, "modules_fp32.pt"))

    # ---- full MobileViTv2 models: contract + small end-to-end run ---------------------------------------------
    contract = {}
    model_fix = {}
    for width, res, seed in ((1.0, 128, 21), (0.5, 64, 22)):
        opts = make_opts(width)
        model = get_model(opts)
        sd = model.state_dict()
        contract[str(width)] = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
        P = O.mobilevit_v2_shapes(width)
        load_seeded(model, P, seed)
        model.train()
        x = O.seeded_input((2, 3, res, res), 300 + seed)
        labels = torch.tensor([3, 977])
        stage_out = {}
        hooks = []
        for name in ["conv_1", "layer_1", "layer_2", "layer_3", "layer_4", "layer_5"]:
            hooks.append(getattr(model, name).register_forward_hook(
                lambda mod, inp, out, name=name: stage_out.__setitem__(name, out.detach().clone())))
        logits = model(x)
        loss = F.cross_entropy(logits, labels, label_smoothing=0.1)
        loss.backward()
        for h in hooks:
            h.remove()
        grads = {k: p.grad for k, p in model.named_parameters()}
        keep_full = [k for k in grads if grads[k].numel() <= 4096]
        model_fix[str(width)] = dict(
            width=width, res=res, seed=seed, x_seed=300 + seed, labels=labels,
            logits=logits.detach().clone(), loss=loss.detach().clone(),
            stage_norms={k: float(v.norm()) for k, v in stage_out.items()},
            stage_sample={k: v.flatten()[:: max(1, v.numel() // 512)][:512].clone() for k, v in stage_out.items()},
            grad_norms={k: float(g.norm()) for k, g in grads.items()},
            grad_small={k: grads[k].clone() for k in keep_full},
            buffers_after={k: b.detach().clone() for k, b in model.named_buffers() if b.numel() <= 4096},
        )
    torch.save(model_fix, os.path.join(HERE, "mobilevit_v2_fp32.pt"))
    with open(os.path.join(HERE, "state_dict_contract.json"), "w") as f:
        json.dump(contract, f)
    for fn in ("modules_fp32.pt", "mobilevit_v2_fp32.pt", "state_dict_contract.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
