"""CLIP (mirror of cvnets/models/multi_modal_img_text/clip.py, cvnets/text_encoders/transformer.py:23-440, image_projection_layers/
simple_projection_head.py and loss_fn/multi_modal_img_text/contrastive_loss_clip.py) -- BASELINE.json configs[4].

    model = CLIP(default_clip_opts())                    # ViT-B/16 image tower + 12-layer text transformer, projection 512
    img, txt, logit_scale = model(images, text_tokens)   # L2-normalised features
    loss = clip_contrastive_loss(img, txt, logit_scale)  # all-gather over the process group when distributed

Same attribute names / ``state_dict`` keys as the reference (image_encoder.*, text_encoder.{embedding_layer, positional_embedding,
transformer.{i}, final_layer_norm, projection_layer}, logit_scale).  Host code only; every kernel is the library's.  Not implemented:
zero-shot evaluation paths, key_padding_mask with causal masking off, sinusoidal embeddings, dropout > 0 in training.
"""
from __future__ import annotations

import argparse
import math
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import functional as Fn
from .layers import Dropout, get_normalization_layer
from .models_vit import PositionalEmbedding, VisionTransformer, default_vit_opts
from .modules import TransformerEncoder, _require_cuda
from .ops import PreparedWeights as PW


def default_clip_opts(vit_mode: str = "base", projection_dim: int = 512, text_dim: int = 512, text_layers: int = 12, text_heads: int = 8,
                      vocab_size: int = 49408, context_length: int = 77, **extra) -> argparse.Namespace:
    """config/multi_modal_img_text/clip_vit.yaml model section."""
    opts = default_vit_opts(vit_mode)
    kv = {
        "model.multi_modal_image_text.name": "clip", "model.multi_modal_image_text.clip.projection_dim": projection_dim,
        "model.text.name": "transformer", "model.text.transformer.model_dim": text_dim, "model.text.transformer.n_transformer_layers": text_layers,
        "model.text.transformer.n_heads_per_layer": text_heads, "model.text.transformer.ffn_multiplier_per_layer": 4.0,
        "model.text.transformer.causal_masking": True, "model.text.transformer.norm_layer": "layer_norm_fp32",
        "model.text.transformer.dropout": 0.0, "model.text.transformer.attn_dropout": 0.0, "model.text.transformer.ffn_dropout": 0.0,
        "model.text.transformer.no_pos_embedding": False, "dataset.text_vocab_size": vocab_size, "dataset.text_context_length": context_length,
        "dataset.padding_index": None,
    }
    kv.update(extra)
    for k, v in kv.items():
        setattr(opts, k, v)
    return opts


class Embedding(nn.Embedding):
    """cvnets/layers/embedding.py."""

    def __init__(self, opts, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None, *args, **kwargs):
        super().__init__(num_embeddings=num_embeddings, embedding_dim=embedding_dim, padding_idx=padding_idx)


class _Projection(nn.Module):
    """Holder of the kernel-layout caches of an [in, out] projection parameter (the parameter itself stays on its owner)."""

    def __init__(self):
        super().__init__()
        self._cfg = None

    def apply_to(self, x: Tensor, P: nn.Parameter, owner: nn.Module) -> Tensor:
        if self._cfg is None:
            prep = PW()
            self._cfg = SimpleNamespace(prep=prep, i_p=prep.add(P, PW.KIND_ROWMAJOR), i_pt=prep.add(P, PW.KIND_TRANSPOSED))
        cfg = self._cfg
        cfg.ws, cfg.plist = getattr(owner, "_ws", None), [P]
        cfg.prep.prepare(force=owner.training)
        return Fn.ProjectionFn.apply(x, cfg, P)


class TextTransformer(nn.Module):
    def __init__(self, opts, projection_dim: int, *args, **kwargs) -> None:
        super().__init__()
        d = getattr(opts, "model.text.transformer.model_dim", 512)
        n_layers = getattr(opts, "model.text.transformer.n_transformer_layers", 6)
        heads = getattr(opts, "model.text.transformer.n_heads_per_layer", 8)
        mult = getattr(opts, "model.text.transformer.ffn_multiplier_per_layer", 4.0)
        norm_layer = getattr(opts, "model.text.transformer.norm_layer", "layer_norm")
        self.vocab_size = getattr(opts, "dataset.text_vocab_size")
        ctx_len = getattr(opts, "dataset.text_context_length")
        if getattr(opts, "dataset.padding_index", None) is not None:
            raise NotImplementedError("padding_idx is not implemented")
        self.projection_dim = projection_dim
        self.embedding_layer = Embedding(opts=opts, embedding_dim=d, padding_idx=None, num_embeddings=self.vocab_size)
        self.embed_scale = d ** -0.5
        no_pos = getattr(opts, "model.text.transformer.no_pos_embedding", False)
        self.positional_embedding = None if no_pos else PositionalEmbedding(opts=opts, num_embeddings=ctx_len, embedding_dim=d, is_learnable=True)
        self.embedding_dropout = Dropout(p=getattr(opts, "model.text.transformer.embed_dropout", 0.0))
        ffn_dims = [int(math.ceil(d * mult / 16.0) * 16.0)] * n_layers
        self.transformer = nn.ModuleList([
            TransformerEncoder(opts=opts, embed_dim=d, num_heads=heads, ffn_latent_dim=ffn_dims[i],
                               attn_dropout=getattr(opts, "model.text.transformer.attn_dropout", 0.0),
                               ffn_dropout=getattr(opts, "model.text.transformer.ffn_dropout", 0.0),
                               dropout=getattr(opts, "model.text.transformer.dropout", 0.0), transformer_norm_layer=norm_layer)
            for i in range(n_layers)])
        self.final_layer_norm = get_normalization_layer(opts, num_features=d, norm_type=norm_layer)
        self.projection_layer = nn.Parameter(torch.empty(d, projection_dim))
        self.model_dim = d
        self.causal_masking = getattr(opts, "model.text.transformer.causal_masking", False)
        self.reset_parameters_clip_style()
        self._emb = SimpleNamespace()
        self._proj = _Projection()
        self._mask = None

    def reset_parameters_clip_style(self):
        """transformer.py:181-211."""
        nn.init.normal_(self.embedding_layer.weight, mean=0.0, std=0.02)
        attn_std = self.model_dim ** -0.5
        proj_std = attn_std * ((2 * len(self.transformer)) ** -0.5)
        fc_std = (2 * self.model_dim) ** -0.5
        for block in self.transformer:
            nn.init.normal_(block.pre_norm_mha[1].qkv_proj.weight, mean=0.0, std=attn_std)
            nn.init.normal_(block.pre_norm_mha[1].out_proj.weight, mean=0.0, std=proj_std)
            nn.init.normal_(block.pre_norm_ffn[1].weight, mean=0.0, std=fc_std)
            nn.init.normal_(block.pre_norm_ffn[4].weight, mean=0.0, std=proj_std)
        nn.init.normal_(self.projection_layer, mean=0.0, std=attn_std)

    def build_attention_mask(self, context_length: int, batch_size: int, device) -> Tensor:
        """transformer.py:343-353: additive causal mask, -inf above the diagonal, expanded over the batch."""
        if self._mask is None or self._mask.shape[0] != batch_size or self._mask.shape[1] != context_length or self._mask.device != device:
            m = torch.full((context_length, context_length), float("-inf"), device=device).triu_(1)
            self._mask = m.unsqueeze(0).expand(batch_size, -1, -1).contiguous()
        return self._mask

    def forward(self, text_tokens: Tensor, key_padding_mask: Optional[Tensor] = None, *args, **kwargs) -> Tensor:
        _require_cuda(text_tokens, "TextTransformer")
        if text_tokens.dim() != 2:
            raise NotImplementedError("zero-shot text batches [B, classes, captions, L] are not implemented")
        if self.training and self.embedding_dropout.p > 0:
            raise NotImplementedError("embedding dropout > 0 in training mode is not implemented")
        pe = self.positional_embedding.pos_embed.pos_embed if self.positional_embedding is not None else None
        if pe is not None and pe.shape[2] != text_tokens.shape[1]:
            raise NotImplementedError("interpolated positional embeddings are not implemented (sequence length must equal the context length)")
        emb = self._emb
        emb.ws, emb.plist = getattr(self, "_ws", None), [self.embedding_layer.weight] + ([pe] if pe is not None else [])
        tokens = text_tokens.contiguous()
        x = Fn.EmbeddingFn.apply(tokens, emb, self.embedding_layer.weight, pe)
        attn_mask = None
        if self.causal_masking:
            attn_mask = self.build_attention_mask(tokens.shape[1], tokens.shape[0], tokens.device)
            key_padding_mask = None
        for layer in self.transformer:
            x = layer(x, key_padding_mask=key_padding_mask, attn_mask=attn_mask)
        x = Fn.EotGatherFn.apply(x, tokens)       # LayerNorm is per token: normalising only the gathered token equals norm-then-gather
        x = self.final_layer_norm(x)
        x = self._proj.apply_to(x, self.projection_layer, self)
        return Fn.L2NormFn.apply(x)


class SimpleImageProjectionHead(nn.Module):
    """image_projection_layers/simple_projection_head.py:20-80 (``simple_projection_nc2nc``): x @ proj, then F.normalize."""

    def __init__(self, opts, in_dim: int, out_dim: int, *args, **kwargs) -> None:
        super().__init__()
        self.proj = nn.Parameter((in_dim ** -0.5) * torch.randn(size=(in_dim, out_dim)))
        self.in_dim, self.out_dim, self.feature_normalization = in_dim, out_dim, True
        self._proj = _Projection()

    def forward(self, x: Tensor, *args, **kwargs) -> Tensor:
        return Fn.L2NormFn.apply(self._proj.apply_to(x, self.proj, self))


class CLIP(nn.Module):
    def __init__(self, opts, *args, **kwargs) -> None:
        super().__init__()
        proj = getattr(opts, "model.multi_modal_image_text.clip.projection_dim", 256)
        self.image_encoder = VisionTransformer(opts)
        self.image_encoder.classifier = SimpleImageProjectionHead(opts, self.image_encoder.embed_dim, proj)
        self.text_encoder = TextTransformer(opts, projection_dim=proj)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1.0 / 0.07))

    def forward(self, images: Tensor, text_tokens: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        return self.image_encoder(images), self.text_encoder(text_tokens), self.logit_scale


def clip_contrastive_loss(image_features: Tensor, text_features: Tensor, logit_scale: Tensor, process_group=None, _cfg=None) -> Tensor:
    """ContrastiveLossClip (loss_fn/multi_modal_img_text/contrastive_loss_clip.py:56-97): features of all ranks are gathered when a process
    group is initialised (gather_all_features), the labels are the global diagonal."""
    import torch.distributed as dist
    cfg = _cfg if _cfg is not None else SimpleNamespace(scale=None, ws=None)
    if not hasattr(cfg, "world"):
        on = dist.is_available() and dist.is_initialized()
        cfg.world, cfg.rank, cfg.group = (dist.get_world_size(process_group), dist.get_rank(process_group), process_group) if on else (1, 0, None)
    return Fn.ClipLossFn.apply(image_features, text_features, logit_scale, cfg)
