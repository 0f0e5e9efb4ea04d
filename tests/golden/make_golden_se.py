"""InvertedResidualSE fixtures FROM THE REAL REFERENCE (cvnets/modules/mobilenetv2.py:16-138, cvnets/modules/squeeze_excitation.py); see
make_golden.py for the method.  Three MobileNetv3-style configurations: hard_swish + SE with a residual, relu + SE with stride 2, relu without SE.
fc1's activation is the model-wide ``model.activation.name`` (relu, the reference default and what the MobileNetv3 recipes use).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_se.py
"""
import copy
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import O, load_seeded, make_opts, run_module, strip, torch  # noqa: E402

from cvnets.modules import InvertedResidualSE  # noqa: E402

CASES = {
    "ir_se_hs_res": dict(cin=24, cout=24, expand_ratio=4, stride=1, use_se=True, act_fn_name="hard_swish", shape=(4, 24, 12, 10), seed=51),
    "ir_se_relu_s2": dict(cin=16, cout=40, expand_ratio=3, stride=2, use_se=True, act_fn_name="relu", shape=(4, 16, 12, 12), seed=52),
    "ir_nose_relu": dict(cin=16, cout=16, expand_ratio=2, stride=1, use_se=False, act_fn_name="relu", shape=(4, 16, 8, 8), seed=53),
}


def main():
    torch.manual_seed(0)
    opts = copy.deepcopy(make_opts(1.0))
    setattr(opts, "model.activation.name", "relu")
    fx = {}
    for name, c in CASES.items():
        P = {}
        O.inverted_residual_se_shapes(P, "m", c["cin"], c["cout"], c["expand_ratio"], use_se=c["use_se"])
        m = InvertedResidualSE(opts, c["cin"], c["cout"], c["expand_ratio"], stride=c["stride"], use_se=c["use_se"], act_fn_name=c["act_fn_name"])
        load_seeded(m, strip("m.", P), c["seed"])
        cfg = {k: v for k, v in c.items() if k not in ("shape", "seed")}
        fx[name] = dict(cfg=cfg, seed=c["seed"], **run_module(m, O.seeded_input(c["shape"], 100 + c["seed"]), 200 + c["seed"]))
        print(name, tuple(fx[name]["y"].shape), float(fx[name]["y"].abs().mean()))
    torch.save(fx, os.path.join(HERE, "inverted_residual_se_fp32.pt"))


if __name__ == "__main__":
    main()
