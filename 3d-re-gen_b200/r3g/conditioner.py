"""Image conditioner, interface of Hunyuan3D-2/hy3dgen/shapegen/models/conditioner.py:57-131,239-257:
SingleImageEncoder(main_image_encoder=DinoImageEncoder) -> {'main': [B, 1370, 1536]}; the unconditional
embedding is zeros.  DINOv2-giant runs once per object (~3 TFLOP) through HF transformers' Dinov2Model on the
GPU (library code: SURVEY.md section 8a row a2 keeps it host-side PyTorch; it is a 'next' row of 8f)."""
import torch
import torch.nn.functional as F

DINOV2_GIANT = dict(hidden_size=1536, num_hidden_layers=40, num_attention_heads=24, mlp_ratio=4, patch_size=14,
                    image_size=518, use_swiglu_ffn=True, layerscale_value=1.0, qkv_bias=True,
                    hidden_act="gelu", layer_norm_eps=1e-6)


class DinoImageEncoder:
    mean = [0.485, 0.456, 0.406]
    std = [0.229, 0.224, 0.225]

    def __init__(self, version=None, config=None, use_cls_token=True, image_size=224, device="cuda",
                 dtype=torch.float16, **kwargs):
        from transformers import Dinov2Config, Dinov2Model
        if config is None and version is not None:
            self.model = Dinov2Model.from_pretrained(version)
        else:
            cfg = Dinov2Config(**(config or DINOV2_GIANT))
            with torch.device(device):
                self.model = Dinov2Model(cfg)
        self.model = self.model.to(device=device, dtype=dtype).eval().requires_grad_(False)
        self.use_cls_token = use_cls_token
        self.image_size = image_size
        self.num_patches = (image_size // 14) ** 2 + (1 if use_cls_token else 0)

    def _transform(self, image):
        """Resize(518, bilinear, antialias) + CenterCrop(518) + Normalize (conditioner.py:78-88)."""
        s = self.image_size
        h, w = image.shape[-2:]
        if h <= w:   # torchvision's Resize(int): short edge -> s, long edge -> int(s * long / short) (truncated)
            nh, nw = s, int(s * w / h)
        else:
            nh, nw = int(s * h / w), s
        image = F.interpolate(image, size=(nh, nw), mode="bilinear", antialias=True, align_corners=False)
        top, left = int(round((nh - s) / 2.0)), int(round((nw - s) / 2.0))   # torchvision's center_crop offsets
        image = image[..., top:top + s, left:left + s]
        mean = torch.tensor(self.mean, device=image.device, dtype=image.dtype)[None, :, None, None]
        std = torch.tensor(self.std, device=image.device, dtype=image.dtype)[None, :, None, None]
        return (image - mean) / std

    @torch.no_grad()
    def __call__(self, image, mask=None, value_range=(-1, 1), **kwargs):
        if value_range is not None:
            lo, hi = value_range
            image = (image - lo) / (hi - lo)
        p = next(self.model.parameters())
        image = image.to(p.device, dtype=p.dtype)
        hidden = self.model(self._transform(image)).last_hidden_state
        return hidden if self.use_cls_token else hidden[:, 1:, :]

    def unconditional_embedding(self, batch_size, **kwargs):
        p = next(self.model.parameters())
        return torch.zeros(batch_size, self.num_patches, self.model.config.hidden_size, device=p.device, dtype=p.dtype)


class SingleImageEncoder:
    def __init__(self, main_image_encoder=None, **kwargs):
        self.main_image_encoder = main_image_encoder if main_image_encoder is not None else DinoImageEncoder(**kwargs)

    def __call__(self, image, mask=None, **kwargs):
        return {"main": self.main_image_encoder(image, mask=mask, **kwargs)}

    forward = __call__

    def unconditional_embedding(self, batch_size, **kwargs):
        return {"main": self.main_image_encoder.unconditional_embedding(batch_size, **kwargs)}
