"""SD-1.5-shaped UNet2DConditionModel stand-in (TEST / BENCH SCAFFOLDING, not product).

diffusers is not installed in the build image and there is no network, so the
benchmark needs a UNet body with the shapes, FLOPs and module surface of
Stable Diffusion 1.5 (block_out 320/640/1280/1280, 2 layers
per block, 8 heads, cross-attention dim 768, GroupNorm-32, GEGLU feed-forward)
to host the FRESCO hooks.  Random-init weights; plain torch modules (cuDNN /
cuBLAS).  It exposes exactly the surface the FRESCO hook code touches:

* ``unet.attn_processors`` / ``unet.set_attn_processor`` with diffusers' key
  names (``up_blocks.2.attentions.0.transformer_blocks.0.attn1.processor`` ...),
* ``Attention`` modules with ``to_q/to_k/to_v/to_out/heads/...`` whose forward
  calls ``self.processor(self, hidden_states, encoder_hidden_states=...)``,
* ``unet.up_blocks[i](hidden_states=..., temb=..., res_hidden_states_tuple=...)``,
* ``forward(sample, timestep, encoder_hidden_states, ..., return_dict)`` with
  ControlNet residual arguments.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_dim: Optional[int], heads: int):
        super().__init__()
        self.heads = heads
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.scale = (query_dim // heads) ** -0.5
        kv_dim = cross_dim if cross_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, query_dim, bias=False)
        self.to_k = nn.Linear(kv_dim, query_dim, bias=False)
        self.to_v = nn.Linear(kv_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim), nn.Dropout(0.0)])
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor

    def prepare_attention_mask(self, attention_mask, target_length, batch_size):
        return attention_mask

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask)


class PlainProcessor:
    """default processor: torch SDPA (what diffusers' AttnProcessor2_0 does)"""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        b = hidden_states.shape[0]
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        h = attn.heads
        q = attn.to_q(hidden_states)
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        d = q.shape[-1] // h
        o = F.scaled_dot_product_attention(q.view(b, -1, h, d).transpose(1, 2), k.view(b, -1, h, d).transpose(1, 2),
                                           v.view(b, -1, h, d).transpose(1, 2), attn_mask=attention_mask)
        o = o.transpose(1, 2).reshape(b, -1, h * d)
        return attn.to_out[1](attn.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, cross_dim):
        super().__init__()
        self.norm = nn.GroupNorm(32, channels, eps=1e-6)
        self.proj_in = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, cross_dim)])
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def forward(self, x, encoder_hidden_states):
        b, c, h, w = x.shape
        res = x
        y = self.proj_in(self.norm(x)).permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            y = blk(y, encoder_hidden_states)
        y = y.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(y) + res


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return h + (x if self.conv_shortcut is None else self.conv_shortcut(x))


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb_dim, heads, cross_dim, has_attn, add_down):
        super().__init__()
        self.has_cross_attention = has_attn
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim) for i in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in range(2)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, hidden_states, temb, encoder_hidden_states=None):
        outs = ()
        for i, r in enumerate(self.resnets):
            hidden_states = r(hidden_states, temb)
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            hidden_states = self.downsamplers[0](hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class MidBlock(nn.Module):
    def __init__(self, c, temb_dim, heads, cross_dim):
        super().__init__()
        self.has_cross_attention = True
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_dim), ResnetBlock2D(c, c, temb_dim)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, cross_dim)])

    def forward(self, x, temb, encoder_hidden_states=None):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, encoder_hidden_states)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev_out, temb_dim, heads, cross_dim, has_attn, add_up):
        super().__init__()
        self.has_cross_attention = has_attn
        res = []
        for i in range(3):
            skip = cin if i == 2 else cout
            res.append(ResnetBlock2D((prev_out if i == 0 else cout) + skip, cout, temb_dim))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in range(3)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, hidden_states, temb=None, res_hidden_states_tuple=(), encoder_hidden_states=None, **kwargs):
        for i, r in enumerate(self.resnets):
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = r(torch.cat([hidden_states, skip], dim=1), temb)
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states)
        if self.upsamplers is not None:
            hidden_states = self.upsamplers[0](hidden_states)
        return hidden_states


class SD15UNet(nn.Module):
    """UNet2DConditionModel with Stable-Diffusion-1.5 hyper-parameters."""

    def __init__(self, block_out=(320, 640, 1280, 1280), heads=8, cross_dim=768, in_channels=4):
        super().__init__()
        self.config = SimpleNamespace(in_channels=in_channels, block_out_channels=block_out,
                                      cross_attention_dim=cross_dim, attention_head_dim=heads,
                                      center_input_sample=False)
        temb = block_out[0] * 4
        self.time_embed_dim0 = block_out[0]
        self.time_embedding = nn.Sequential(nn.Linear(block_out[0], temb), nn.SiLU(), nn.Linear(temb, temb))
        self.conv_in = nn.Conv2d(in_channels, block_out[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cin = block_out[0]
        for i, cout in enumerate(block_out):
            last = i == len(block_out) - 1
            self.down_blocks.append(DownBlock(cin, cout, temb, heads, cross_dim, has_attn=not last, add_down=not last))
            cin = cout
        self.mid_block = MidBlock(block_out[-1], temb, heads, cross_dim)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(block_out))
        prev = rev[0]
        for i, cout in enumerate(rev):
            cin_skip = rev[min(i + 1, len(rev) - 1)]
            last = i == len(rev) - 1
            self.up_blocks.append(UpBlock(cin_skip, cout, prev, temb, heads, cross_dim, has_attn=i > 0, add_up=not last))
            prev = cout
        self.num_upsamplers = len(block_out) - 1
        self.conv_norm_out = nn.GroupNorm(32, block_out[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out[0], in_channels, 3, padding=1)
        self.set_attn_processor(PlainProcessor())

    # ---- diffusers-compatible processor plumbing
    @property
    def attn_processors(self) -> Dict[str, object]:
        out = {}
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                out[name + ".processor"] = m.processor
        return out

    def set_attn_processor(self, processor):
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                m.set_processor(processor[name + ".processor"] if isinstance(processor, dict) else processor)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def _time_proj(self, timesteps: torch.Tensor) -> torch.Tensor:
        half = self.time_embed_dim0 // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
        emb = timesteps[:, None].float() * torch.exp(exponent)[None]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)          # flip_sin_to_cos=True

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None, encoder_attention_mask=None,
                return_dict: bool = True):
        if not torch.is_tensor(timestep):
            # one device tensor per distinct python timestep, made once: a host-to-device copy per forward would be a
            # sync (and is not allowed while a CUDA graph is being captured)
            key = (int(timestep), str(sample.device))
            cache = self.__dict__.setdefault("_timestep_cache", {})
            if key not in cache:
                cache[key] = torch.tensor([int(timestep)], dtype=torch.int64, device=sample.device)
            timestep = cache[key]
        elif timestep.ndim == 0:
            timestep = timestep[None].to(sample.device)
        timestep = timestep.expand(sample.shape[0])
        emb = self.time_embedding(self._time_proj(timestep).to(sample.dtype))
        sample = self.conv_in(sample)
        res = (sample,)
        for blk in self.down_blocks:
            sample, outs = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
            res += outs
        if down_block_additional_residuals is not None:
            res = tuple(r + a for r, a in zip(res, down_block_additional_residuals))
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        if mid_block_additional_residual is not None:
            sample = sample + mid_block_additional_residual
        for blk in self.up_blocks:
            n = len(blk.resnets)
            skips, res = res[-n:], res[:-n]
            sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=skips,
                         encoder_hidden_states=encoder_hidden_states)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        if not return_dict:
            return (sample,)
        return SimpleNamespace(sample=sample)


class FakePipe:
    """The two attributes of a diffusers pipeline that the FRESCO hook functions touch."""

    def __init__(self, unet):
        self.unet = unet
