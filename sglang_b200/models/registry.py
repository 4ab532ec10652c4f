"""Architecture registry (reference ``python/sglang/srt/models/registry.py:20`` EntryClass scan)."""
from __future__ import annotations

from typing import Callable, Dict

_REGISTRY: Dict[str, Callable] = {}


def register_model(arch: str, cls: Callable) -> None:
    _REGISTRY[arch] = cls


def get_model_class(arch: str):
    _ensure()
    if arch not in _REGISTRY:
        raise ValueError(f"unsupported architecture {arch!r}; supported: {sorted(_REGISTRY)}")
    return _REGISTRY[arch]


def supported_architectures():
    _ensure()
    return sorted(_REGISTRY)


_BUILTINS_DONE = False


def _ensure() -> None:
    global _BUILTINS_DONE
    if _BUILTINS_DONE:        # (a flag, not "registry non-empty": plugins may register architectures before the built-ins are imported)
        return
    plugin_entries = dict(_REGISTRY)
    _register_builtins()
    from ..configs.model_config import _HF_ARCH_MAP

    for alias, base in _HF_ARCH_MAP.items():          # checkpoints whose architecture is a configuration of a registered family
        if base in _REGISTRY:
            _REGISTRY.setdefault(alias, _REGISTRY[base])
    _REGISTRY.update(plugin_entries)      # an out-of-tree class registered under a built-in name wins
    _BUILTINS_DONE = True


def _register_builtins() -> None:
    from . import llama, llava, mixtral

    for name in ("LlamaForCausalLM", "MistralForCausalLM", "Qwen2ForCausalLM", "Qwen3ForCausalLM", "GemmaForCausalLM"):
        register_model(name, getattr(llama, name))
    register_model("MixtralForCausalLM", mixtral.MixtralForCausalLM)
    register_model("MiniMaxM2ForCausalLM", mixtral.MiniMaxM2ForCausalLM)
    from . import phimoe as _pm

    register_model("PhimoeForCausalLM", _pm.PhimoeForCausalLM)
    register_model("PhiMoEForCausalLM", _pm.PhimoeForCausalLM)          # (the architecture string of the released checkpoints)
    register_model("LlavaForConditionalGeneration", llava.LlavaForConditionalGeneration)
    register_model("LlavaLlamaForCausalLM", llava.LlavaForConditionalGeneration)
    from . import families as _fam

    for name in ("Phi3ForCausalLM", "InternLM2ForCausalLM", "Qwen2MoeForCausalLM", "Qwen3MoeForCausalLM", "Gemma2ForCausalLM",
                 "GraniteForCausalLM", "Olmo2ForCausalLM", "Exaone4ForCausalLM", "VaultGemmaForCausalLM"):
        register_model(name, getattr(_fam, name))
    from . import classic as _cl

    for name in ("GPT2LMHeadModel", "OPTForCausalLM", "GPTNeoXForCausalLM", "Starcoder2ForCausalLM", "PhiForCausalLM", "GPTJForCausalLM",
                 "GPTBigCodeForCausalLM", "CodeGenForCausalLM", "GPTNeoForCausalLM", "PersimmonForCausalLM"):
        register_model(name, getattr(_cl, name))
    from . import llama4 as _l4

    register_model("Llama4ForCausalLM", _l4.Llama4ForCausalLM)
    register_model("Llama4ForConditionalGeneration", _l4.Llama4ForCausalLM)       # text tower of the multimodal checkpoint
    from . import qwen3_next as _qn

    register_model("Qwen3NextForCausalLM", _qn.Qwen3NextForCausalLM)
    register_model("Qwen3_5ForCausalLM", _qn.Qwen3_5ForCausalLM)
    register_model("Qwen3_5MoeForCausalLM", _qn.Qwen3_5MoeForCausalLM)
    register_model("Qwen3_5ForConditionalGeneration", _qn.Qwen3_5ForCausalLM)            # text tower of the multimodal checkpoints
    register_model("Qwen3_5MoeForConditionalGeneration", _qn.Qwen3_5MoeForCausalLM)
    from . import internvl as _iv

    register_model("InternVLForConditionalGeneration", _iv.InternVLForConditionalGeneration)
    from . import whisper as _wh

    register_model("WhisperForConditionalGeneration", _wh.WhisperForConditionalGeneration)
    from . import families2 as _f2

    for name in ("GlmForCausalLM", "Glm4ForCausalLM", "Gemma3ForCausalLM", "CohereForCausalLM", "OlmoeForCausalLM", "NemotronForCausalLM",
                 "Glm4MoeForCausalLM", "Dots1ForCausalLM", "Ernie4_5_MoeForCausalLM", "OlmoForCausalLM", "StableLmForCausalLM", "ApertusForCausalLM"):
        register_model(name, getattr(_f2, name))
    from . import gemma3_mm as _g3

    register_model("FlexOlmoForCausalLM", _f2._make_flex_olmo())
    register_model("HunYuanDenseV1ForCausalLM", _f2.HunYuanDenseV1ForCausalLM)
    register_model("HunYuanMoEV1ForCausalLM", _f2._make_hunyuan_moe())
    register_model("Gemma3ForConditionalGeneration", _g3.Gemma3ForConditionalGeneration)
    from . import qwen2_vl as _qv

    register_model("Qwen2VLForConditionalGeneration", _qv.Qwen2VLForConditionalGeneration)
    register_model("Qwen2_5_VLForConditionalGeneration", _qv.Qwen2_5_VLForConditionalGeneration)
    register_model("Qwen3VLForConditionalGeneration", _qv.Qwen3VLForConditionalGeneration)
    register_model("Qwen3VLMoeForConditionalGeneration", _qv._make_qwen3_vl_moe())
    from . import glm4v as _g4v

    register_model("Glm4vForConditionalGeneration", _g4v.Glm4vForConditionalGeneration)
    register_model("Glm4vMoeForConditionalGeneration", _g4v.Glm4vMoeForConditionalGeneration)
    from . import qwen2_audio as _qa

    register_model("Qwen2AudioForConditionalGeneration", _qa.Qwen2AudioForConditionalGeneration)
    register_model("VoxtralForConditionalGeneration", _qa.VoxtralForConditionalGeneration)
    from . import pixtral as _px

    register_model("Mistral3ForConditionalGeneration", _px.Mistral3ForConditionalGeneration)
    register_model("PixtralForConditionalGeneration", _px.PixtralForConditionalGeneration)
    register_model("LightOnOcrForConditionalGeneration", _px.LightOnOCRForConditionalGeneration)
    register_model("LightOnOCRForConditionalGeneration", _px.LightOnOCRForConditionalGeneration)
    from . import gpt_oss as _go

    register_model("GptOssForCausalLM", _go.GptOssForCausalLM)
    from . import mamba as _mb

    register_model("NemotronHForCausalLM", _mb.NemotronHForCausalLM)
    register_model("Mamba2ForCausalLM", _mb.Mamba2ForCausalLM)
    register_model("BambaForCausalLM", _mb.BambaForCausalLM)
    for name in ("MambaForCausalLM", "FalconMambaForCausalLM", "JambaForCausalLM", "GraniteMoeHybridForCausalLM", "FalconH1ForCausalLM"):
        register_model(name, getattr(_mb, name))
    register_model("GraniteMoeForCausalLM", _mb.GraniteMoeHybridForCausalLM)          # Granite 3.x MoE: the same block, attention-only pattern
    register_model("GraniteMoeSharedForCausalLM", _mb.GraniteMoeHybridForCausalLM)
    from . import deepseek as _ds

    register_model("DeepseekV2ForCausalLM", _ds.DeepseekV2ForCausalLM)
    register_model("DeepseekV3ForCausalLM", _ds.DeepseekV3ForCausalLM)
    from . import embedding as _emb

    register_model("LlamaEmbeddingModel", _emb.LlamaEmbeddingModel)
    register_model("LlamaForSequenceClassification", _emb.LlamaForSequenceClassification)
    for fam in ("Qwen2", "Qwen3", "Gemma2", "Phi3", "Gemma"):
        if fam + "ForCausalLM" in _REGISTRY:
            register_model(fam + "ForSequenceClassification", _emb.make_classifier(_REGISTRY[fam + "ForCausalLM"]))
    register_model("MistralForSequenceClassification", _emb.LlamaForSequenceClassification)
    register_model("Qwen2ForRewardModel", _emb.make_reward_model(_REGISTRY["Qwen2ForCausalLM"], "Qwen2ForRewardModel"))
    register_model("Qwen3ForRewardModel", _emb.make_reward_model(_REGISTRY["Qwen3ForCausalLM"], "Qwen3ForRewardModel"))
    register_model("InternLM2ForRewardModel", _emb.make_v_head_reward(_REGISTRY["InternLM2ForCausalLM"], "InternLM2ForRewardModel"))
    register_model("LlamaForSequenceClassificationWithNormal_Weights", _emb.LlamaForSequenceClassificationWithNormal_Weights)
    register_model("Qwen3Model", type("Qwen3EmbeddingModel", (_emb.LlamaEmbeddingModel, _REGISTRY["Qwen3ForCausalLM"]), {}))
    register_model("MistralModel", _emb.LlamaEmbeddingModel)
    register_model("Qwen2Model", type("Qwen2EmbeddingModel", (_emb.LlamaEmbeddingModel,), {}))
    from . import transformers_fallback as _tff

    register_model("TransformersForCausalLM", _tff.TransformersForCausalLM)
    register_model("TransformersMoEForCausalLM", _tff.TransformersForCausalLM)
    from . import bert as _bert

    for name in ("BertModel", "XLMRobertaModel", "BertForSequenceClassification", "XLMRobertaForSequenceClassification"):
        register_model(name, getattr(_bert, name))
    register_model("Contriever", _bert.Contriever)
    register_model("RobertaModel", _bert.XLMRobertaModel)
    register_model("RobertaForSequenceClassification", _bert.XLMRobertaForSequenceClassification)
