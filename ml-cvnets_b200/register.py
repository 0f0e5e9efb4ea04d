"""Injection into an importable apple/ml-cvnets checkout: the reference has no operator/FFI boundary, its
plug-in mechanism is ``MODEL_REGISTRY`` plus the import namespaces ``cvnets.modules`` / ``cvnets.layers``.

    import ml_cvnets_b200.register as r; r.register_with_cvnets()
    opts.model.classification.name = "mobilevit_v2_b200"   # or keep "mobilevit_v2" after rebind_modules()
    model = cvnets.get_model(opts)                          # engine/training_engine.py consumes it unchanged

Nothing here is needed (or importable) on the GPU box, where the reference does not exist; the product is standalone.
"""
from __future__ import annotations


def rebind_modules() -> None:
    """Injection point 1: replace the block classes in the namespaces the reference's model files import from
    (``from cvnets.modules import InvertedResidual, MobileViTBlockv2`` at mobilevit_v2.py:15-16).  Must run before the
    first ``MODEL_REGISTRY`` lookup (which lazily imports cvnets/models/**)."""
    import cvnets.modules as cm
    from . import modules as ours
    cm.InvertedResidual = ours.InvertedResidual
    cm.MobileViTBlockv2 = ours.MobileViTBlockv2
    cm.TransformerEncoder = ours.TransformerEncoder  # used by vit.py:29, mobilevit_block.py (v1), text_encoders/transformer.py:20


def register_with_cvnets(name: str = "mobilevit_v2_b200"):
    """Injection point 2: register the B200 assembler under a new model name (utils/registry.py:111-167 requires a
    ``BaseAnyNNModel`` subclass, so the class is derived from the reference's own base encoder)."""
    from cvnets.models import MODEL_REGISTRY
    from cvnets.models.classification.base_image_encoder import BaseImageEncoder
    from .models import MobileViTv2 as Ours

    if f"classification:{name}" in getattr(MODEL_REGISTRY, "registry", {}):
        return MODEL_REGISTRY.registry[f"classification:{name}"]

    class MobileViTv2B200(BaseImageEncoder):
        """Reference-side shell: BaseImageEncoder plumbing (get_trainable_parameters, freeze_norm_layers, info, ...) around
        the B200 modules; forward goes straight to the CUDA path."""

        def __init__(self, opts, *args, **kwargs) -> None:
            super().__init__(opts, *args, **kwargs)
            inner = Ours(opts)
            for attr in ("conv_1", "layer_1", "layer_2", "layer_3", "layer_4", "layer_5", "conv_1x1_exp", "classifier"):
                setattr(self, attr, getattr(inner, attr))
            self.model_conf_dict = inner.model_conf_dict
            self._head = None
            self.forward_classifier = lambda x, *a, **k: Ours.forward_classifier(self, x)

        def forward(self, x, *args, **kwargs):
            return Ours.forward_classifier(self, x)

    MODEL_REGISTRY.register(name=name, type="classification")(MobileViTv2B200)
    return MobileViTv2B200
