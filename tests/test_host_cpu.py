"""CPU-only checks (-m "not gpu"): the C-ABI library loads and exports every symbol include/cvnets_b200.h declares, the
host-side mirror keeps the reference's state_dict / signature contract, the product refuses to run without CUDA, the
reference-side registration works when the reference checkout is present, and the N>1 host logic works under gloo."""
import inspect
import json
import os
import re
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from ml_cvnets_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "cvnets_b200.h")).read()
    declared = set(re.findall(r"CVB_API\s+(?:const\s+char\*|int)\s+(cvb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cvb_abi_version() == _lib.ABI_VERSION
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cvb_[a-z0-9_]+)", out))
    assert declared <= exported


def test_struct_layouts_match_header():
    """ctypes mirrors of the argument structs must list the header's fields in order."""
    from ml_cvnets_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "cvnets_b200.h")).read()

    def fields(struct_name):
        body = dict((n, b) for b, n in re.findall(r"typedef struct \{([^}]*)\} (\w+);", hdr))[struct_name]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        return names

    for cname, cls in (("cvb_gemm_args", _lib.GemmArgs), ("cvb_wgrad_args", _lib.WgradArgs), ("cvb_dw_fwd_args", _lib.DwFwdArgs),
                       ("cvb_dw_bwd_args", _lib.DwBwdArgs), ("cvb_prep_desc", _lib.PrepDesc)):
        assert fields(cname) == [f[0] for f in cls._fields_], cname


def test_state_dict_contract_and_signatures(golden_dir):
    import ml_cvnets_b200 as m
    with open(os.path.join(golden_dir, "state_dict_contract.json")) as f:
        contract = json.load(f)
    for width, entries in contract.items():
        sd = m.MobileViTv2(m.default_opts(width_multiplier=float(width))).state_dict()
        assert list(sd.keys()) == [e[0] for e in entries]
        for k, shape, dtype in entries:
            assert list(sd[k].shape) == shape and str(sd[k].dtype) == "torch." + dtype, k
    # constructor signatures of the drop-ins
    assert list(inspect.signature(m.InvertedResidual.__init__).parameters)[1:8] == [
        "opts", "in_channels", "out_channels", "stride", "expand_ratio", "dilation", "skip_connection"]
    assert list(inspect.signature(m.MobileViTBlockv2.__init__).parameters)[1:14] == [
        "opts", "in_channels", "attn_unit_dim", "ffn_multiplier", "n_attn_blocks", "attn_dropout", "dropout", "ffn_dropout",
        "patch_h", "patch_w", "conv_ksize", "dilation", "attn_norm_layer"]
    assert list(inspect.signature(m.LinearSelfAttention.__init__).parameters)[1:5] == ["opts", "embed_dim", "attn_dropout", "bias"]
    model = m.MobileViTv2(m.default_opts())
    groups, mult = model.get_trainable_parameters(weight_decay=0.05, no_decay_bn_filter_bias=True)
    assert [len(g["params"]) for g in groups] == [65, 129] and mult == [1.0, 1.0]


def test_no_cpu_fallback():
    import ml_cvnets_b200 as m
    model = m.MobileViTv2(m.default_opts(width_multiplier=0.5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.randn(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.InvertedResidual(m.default_opts(), 16, 16, 1, 2)(torch.randn(1, 16, 8, 8))
    # the stand-alone layers have kernel paths of their own (round 2) -- and likewise no CPU path
    for layer, x in ((m.LinearSelfAttention(m.default_opts(), 16), torch.randn(1, 16, 4, 4)), (m.LayerNorm2D_NCHW(16), torch.randn(1, 16, 4, 4)),
                     (m.LayerNorm(16), torch.randn(1, 4, 16)), (m.LinearLayer(16, 16), torch.randn(1, 4, 16)), (m.GlobalPool(), torch.randn(1, 16, 4, 4)),
                     (m.ConvLayer2d(m.default_opts(), 16, 16, 1), torch.randn(1, 16, 4, 4))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            layer(x)
    with pytest.raises(RuntimeError, match="no CPU"):
        m.cross_entropy(torch.randn(2, 8), torch.tensor([1, 2]))


def test_product_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "ml-cvnets_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, fn)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{fn} mentions the oracle"


@pytest.mark.skipif(not os.path.isdir("/root/reference/cvnets"), reason="reference checkout not present (GPU box)")
def test_registration_with_reference_checkout():
    code = r"""
import sys, os, argparse
sys.path.insert(0, %r); sys.path.insert(0, "/root/reference"); os.chdir("/root/reference")
import ml_cvnets_b200.register as r
from cvnets import modeling_arguments, get_model
r.register_with_cvnets()
opts = modeling_arguments(argparse.ArgumentParser()).parse_args([])
for k, v in {"dataset.category": "classification", "model.classification.name": "mobilevit_v2_b200",
             "model.classification.mitv2.width_multiplier": 1.0, "model.activation.name": "swish"}.items():
    setattr(opts, k, v)
ours = get_model(opts)
setattr(opts, "model.classification.name", "mobilevit_v2")
ref = get_model(opts)
assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
ours.load_state_dict(ref.state_dict(), strict=True)
assert type(ours.layer_3[1]).__module__.startswith("ml_cvnets_b200")
g, _ = ours.get_trainable_parameters(weight_decay=0.05, no_decay_bn_filter_bias=True)
import ml_cvnets_b200 as m
assert type(ours).forward is m.MobileViTv2.forward and type(ours).extract_end_points_all is m.MobileViTv2.extract_end_points_all
assert len(ours._chain) == 10 and ours.fuse_boundaries and ours._chain[0] is ours.conv_1          # private state of the B200 model travels
seg = get_model(opts.__class__(**{**vars(opts), "model.classification.name": "mobilevit_v2_b200"}), category="classification", output_stride=8)
assert seg.layer_5[1].local_rep[0].block.conv.dilation == (4, 4)                                   # segmentation heads: output_stride is honoured
# the other registered assemblers: same keys / shapes as the reference models they replace
for ours_name, ref_name, extra in (("mobilevit_b200", "mobilevit", {"model.classification.mit.mode": "xx_small"}),
                                   ("vit_b200", "vit", {"model.classification.vit.mode": "tiny", "model.classification.vit.norm_layer": "layer_norm_fp32",
                                                        "model.activation.name": "gelu", "model.classification.activation.name": "gelu"})):
    for k, v in extra.items():
        setattr(opts, k, v)
    setattr(opts, "model.classification.name", ours_name)
    a = get_model(opts)
    setattr(opts, "model.classification.name", ref_name)
    b = get_model(opts)
    assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == {k: tuple(v.shape) for k, v in b.state_dict().items()}, ours_name
    a.load_state_dict(b.state_dict(), strict=True)
print("OK", sum(p.numel() for p in ours.parameters()), [len(x["params"]) for x in g])
""" % REPO
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 4901841" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/cvnets"), reason="reference checkout not present (GPU box)")
def test_transformer_dropins_match_reference_contract():
    """MultiHeadAttention / TransformerEncoder: same constructor parameters, forward parameters, state_dict keys and shapes as the
    reference classes, checked against the reference checkout itself; rebind_modules() swaps them in."""
    code = r"""
import sys, os, argparse, inspect
sys.path.insert(0, %r); sys.path.insert(0, "/root/reference"); os.chdir("/root/reference")
import ml_cvnets_b200 as ours
import ml_cvnets_b200.register as r
from cvnets import modeling_arguments
from cvnets.layers import MultiHeadAttention as RefMHA
from cvnets.modules import TransformerEncoder as RefEnc
def params(f): return [p for p in inspect.signature(f).parameters if p not in ("args", "kwargs")]
assert params(ours.MultiHeadAttention.__init__) == params(RefMHA.__init__)
assert params(ours.MultiHeadAttention.forward)[:5] == ["self", "x_q", "x_kv", "key_padding_mask", "attn_mask"]
assert params(ours.TransformerEncoder.__init__) == params(RefEnc.__init__)
assert params(ours.TransformerEncoder.forward) == params(RefEnc.forward)
opts = modeling_arguments(argparse.ArgumentParser()).parse_args([])
for act in ("swish", "gelu"):
    setattr(opts, "model.activation.name", act)
    a, b = ours.TransformerEncoder(opts, 64, 128, num_heads=4), RefEnc(opts, 64, 128, num_heads=4)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys()), (list(sa.keys()), list(sb.keys()))
    assert all(sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype for k in sa)
    a.load_state_dict(sb, strict=True)
    assert float(a.pre_norm_mha[0].eps) == float(b.pre_norm_mha[0].eps)
    assert repr(a).split("(")[0] == repr(b).split("(")[0]
m1, m2 = ours.MultiHeadAttention(64, 4), RefMHA(64, 4)
assert list(m1.state_dict().keys()) == list(m2.state_dict().keys())
r.rebind_modules()
import cvnets.modules as cm
assert cm.TransformerEncoder is ours.TransformerEncoder and cm.MobileViTBlockv2 is ours.MobileViTBlockv2
try:
    a(__import__("torch").zeros(2, 5, 64))
    raise SystemExit("CPU input must raise")
except RuntimeError:
    pass
print("OK")
""" % REPO
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/cvnets"), reason="reference checkout not present (GPU box)")
def test_se_block_and_dropout_children_match_reference_contract():
    """InvertedResidualSE / SqueezeExcitation and the dropout / stochastic-depth children of TransformerEncoder: constructor
    parameters, child tree, state_dict keys / shapes and repr head against the reference checkout; rebind_modules() swaps the block in."""
    code = r"""
import sys, os, argparse, inspect
sys.path.insert(0, %r); sys.path.insert(0, "/root/reference"); os.chdir("/root/reference")
import ml_cvnets_b200 as ours
import ml_cvnets_b200.register as r
from cvnets import modeling_arguments
from cvnets.modules import InvertedResidualSE as RefSE, SqueezeExcitation as RefSq, TransformerEncoder as RefEnc
def params(f): return [p for p in inspect.signature(f).parameters if p not in ("args", "kwargs")]
assert params(ours.InvertedResidualSE.__init__) == params(RefSE.__init__)
assert params(ours.SqueezeExcitation.__init__) == params(RefSq.__init__)
opts = modeling_arguments(argparse.ArgumentParser()).parse_args([])
for kw in (dict(expand_ratio=4, stride=1, use_se=True, act_fn_name="hard_swish"), dict(expand_ratio=3, stride=2, use_se=True, act_fn_name="relu"),
           dict(expand_ratio=1, stride=1, use_se=False, act_fn_name="relu"), dict(expand_ratio=2, stride=1, use_se=True, kernel_size=5)):
    a, b = ours.InvertedResidualSE(opts, 24, 24, **kw), RefSE(opts, 24, 24, **kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys()), (list(sa.keys()), list(sb.keys()))
    assert all(sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype for k in sa)
    a.load_state_dict(sb, strict=True)
    assert [n for n, _ in a.block.named_children()] == [n for n, _ in b.block.named_children()]
    assert list(a.block._modules) == list(b.block._modules)                       # incl. the shared activation registered twice
    assert repr(a) == repr(b), (repr(a), repr(b))
    assert a.use_res_connect == b.use_res_connect
e1, e2 = ours.TransformerEncoder(opts, 64, 128, num_heads=4, dropout=0.1, ffn_dropout=0.2), RefEnc(opts, 64, 128, num_heads=4, dropout=0.1, ffn_dropout=0.2)
assert [type(m).__name__ for m in e1.pre_norm_ffn] == [type(m).__name__ for m in e2.pre_norm_ffn]
assert (e1.pre_norm_mha[2].p, e1.pre_norm_ffn[3].p, e1.pre_norm_ffn[5].p) == (e2.pre_norm_mha[2].p, e2.pre_norm_ffn[3].p, e2.pre_norm_ffn[5].p)
s1, s2 = ours.TransformerEncoder(opts, 64, 128, num_heads=4, stochastic_dropout=0.2), RefEnc(opts, 64, 128, num_heads=4, stochastic_dropout=0.2)
assert type(s1.drop_path).__name__ == type(s2.drop_path).__name__ == "StochasticDepth" and s1.drop_path.p == s2.drop_path.p
assert list(s1.state_dict().keys()) == list(s2.state_dict().keys())
r.rebind_modules()
import cvnets.modules as cm
assert cm.InvertedResidualSE is ours.InvertedResidualSE and cm.SqueezeExcitation is ours.SqueezeExcitation
print("OK")
""" % REPO
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout


def _gloo_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    from ml_cvnets_b200 import dist as D
    import ml_cvnets_b200 as m
    r, w, lr = D.init("gloo")
    assert (r, w) == (rank, world)
    # max-over-ranks timing and the weak-scaling aggregate
    mx = D.max_over_ranks(10.0 + rank)
    thr = D.weak_scaling_throughput(128, w, mx)
    # gradient all-reduce semantics on the real parameter set (fp32 grads, SUM / world)
    torch.manual_seed(D.shard_seed(0, rank))
    model = m.MobileViTv2(m.default_opts(width_multiplier=0.5))
    ddp = D.wrap_ddp(model, lr, device_type="cpu")  # constructor broadcasts rank 0's parameters
    p0 = next(model.parameters()).detach().clone()
    grads = [torch.full_like(p, float(rank + 1)) for p in list(model.parameters())[:10]]
    D.allreduce_mean_(grads, w)
    ok_grad = all(torch.allclose(g, torch.full_like(g, (1 + world) / 2.0)) for g in grads)
    gathered = [torch.zeros_like(p0) for _ in range(w)]
    dist.all_gather(gathered, p0)
    ok_bcast = all(torch.equal(gathered[0], t) for t in gathered)
    q.put((rank, mx, thr, ok_grad, ok_bcast, D.shard_seed(0, rank)))
    dist.destroy_process_group()


def test_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29533 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mx, thr, ok_grad, ok_bcast, seed in res:
        assert mx == 11.0 and abs(thr - 2 * 128 / 11e-3) < 1e-6 and ok_grad and ok_bcast
    assert res[0][5] != res[1][5]
