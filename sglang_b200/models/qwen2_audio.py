"""Qwen2-Audio: Whisper-style audio encoder (+ 2x average pooling) -> linear projector -> Qwen2 decoder; audio clips are spliced into the
prompt as runs of soft tokens, exactly like image features.

Parity: reference ``python/sglang/srt/models/qwen2_audio.py`` (``Qwen2AudioForConditionalGeneration``) and
``multimodal/processors/qwen_audio.py`` (each ``<|AUDIO|>`` placeholder expands to the encoder's output length for that clip).

The log-mel input is always padded to the encoder's fixed window (2 x ``max_source_positions`` frames); the clip's real length only
enters as a key-padding mask of the encoder self-attention and as the number of output rows kept: ``L1 = (frames - 1) // 2 + 1`` after the
stride-2 convolution, ``L2 = (L1 - 2) // 2 + 1`` after pooling."""
from __future__ import annotations

from typing import Any, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..configs.model_config import ModelConfig
from .llama import _DT, LlamaForCausalLM
from .whisper import WhisperEncoder


def audio_output_lengths(frames: int) -> Tuple[int, int]:
    l1 = (int(frames) - 1) // 2 + 1
    return l1, (l1 - 2) // 2 + 1


class Qwen2AudioEncoder(WhisperEncoder):
    def forward(self, feats: torch.Tensor, frames: Optional[List[int]] = None) -> List[torch.Tensor]:       # [B, mel, 2 * max_pos] -> per clip [L2, H]
        x = F.gelu(self.conv2(F.gelu(self.conv1(feats)))).transpose(1, 2)
        x = x + self.embed_positions.weight[: x.shape[1]]
        B, S, _ = x.shape
        frames = frames or [feats.shape[-1]] * B
        lens = [audio_output_lengths(f) for f in frames]
        valid = torch.arange(S, device=x.device)[None, :] < torch.tensor([l1 for l1, _ in lens], device=x.device)[:, None]       # [B, S] keys
        mask = valid[:, None, None, :]
        for l in self.layers:
            x = l(x, mask)
        x = F.avg_pool1d(x.transpose(1, 2), 2, 2).transpose(1, 2)
        x = self.layer_norm(x)
        return [x[i, : l2] for i, (_, l2) in enumerate(lens)]


class Qwen2AudioForConditionalGeneration(LlamaForCausalLM):
    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        super().__init__(cfg, quant_name, device)
        a, dt = cfg.audio, _DT[cfg.dtype]
        self.audio_tower = Qwen2AudioEncoder(a, dt)
        self.projector = nn.Linear(a["d_model"], cfg.hidden_size, dtype=dt)
        self.vision_tower = self.audio_tower          # the shared splice path only asks "is there a local encoder"

    @torch.no_grad()
    def encode_audio(self, feats: torch.Tensor, frames: Any) -> torch.Tensor:
        outs = self.audio_tower(feats, [int(f) for f in np.asarray(frames).reshape(-1)])
        return torch.cat([self.projector(o) for o in outs], 0)

    def load_weights(self, weights: Iterable[Tuple[str, torch.Tensor]]) -> List[str]:
        own = dict(self.named_parameters())
        rest, loaded = [], []
        for name, w in weights:
            n = name[len("model."):] if name.startswith(("model.audio_tower.", "model.multi_modal_projector.", "model.language_model.")) else name
            if n.startswith("language_model."):
                n = n[len("language_model."):]
                rest.append((n if n.startswith(("model.", "lm_head.")) else "model." + n, w))
            elif n.startswith(("audio_tower.", "multi_modal_projector.", "projector.", "model.projector.")):
                n = n[len("model."):] if n.startswith("model.projector.") else n
                n = n.replace(".self_attn.", ".").replace("audio_tower.ln_post.", "audio_tower.layer_norm.").replace("multi_modal_projector.linear.", "projector.")
                if n in own:
                    own[n].data.copy_(w.reshape(own[n].shape))
                    loaded.append(name)
            else:
                rest.append((name, w))
        return loaded + super().load_weights(rest)


class VoxtralForConditionalGeneration(Qwen2AudioForConditionalGeneration):
    """Voxtral (reference ``python/sglang/srt/models/voxtral.py``): the plain Whisper encoder over fixed windows (no length mask), groups of
    ``intermediate_size / hidden_size`` consecutive frames concatenated into one soft token, bias-free two-layer projector, Llama / Mistral
    decoder.  Audio longer than one window is cut into windows; every window contributes the same number of tokens."""

    def __init__(self, cfg: ModelConfig, quant_name: Optional[str] = None, device="cpu"):
        LlamaForCausalLM.__init__(self, cfg, quant_name, device)
        a, dt = cfg.audio, _DT[cfg.dtype]
        self.audio_tower = WhisperEncoder(a, dt)
        self.group = int(a["group"])
        self.projector = nn.Sequential(nn.Linear(a["d_model"] * self.group, cfg.hidden_size, bias=False, dtype=dt), nn.GELU(),
                                       nn.Linear(cfg.hidden_size, cfg.hidden_size, bias=False, dtype=dt))
        self.vision_tower = self.audio_tower

    @torch.no_grad()
    def encode_audio(self, feats: torch.Tensor, frames: Any) -> torch.Tensor:
        h = self.audio_tower(feats)                                   # [windows, S, H]
        return self.projector(h.reshape(-1, h.shape[-1] * self.group))

    def load_weights(self, weights):
        def rename():
            for name, w in weights:
                yield name.replace("multi_modal_projector.linear_1.", "projector.0.").replace("multi_modal_projector.linear_2.", "projector.2."), w
        return super().load_weights(rename())
