"""Chat templates for the frontend language (reference ``python/sglang/lang/chat_template.py``: registry keyed by name,
``get_prefix_and_suffix`` per role, model-path matching)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass
class ChatTemplate:
    name: str
    default_system_prompt: Optional[str]
    role_prefix_and_suffix: Dict[str, Tuple[str, str]]
    stop_str: List[str] = field(default_factory=list)
    image_token: str = "<image>"

    def get_prefix_and_suffix(self, role: str, hist_messages: List[Dict]) -> Tuple[str, str]:
        return self.role_prefix_and_suffix.get(role, ("", ""))

    def get_prompt(self, messages: List[Dict]) -> str:
        out = ""
        for i, m in enumerate(messages):
            if i == 0 and m["role"] != "system" and self.default_system_prompt:
                p, s = self.get_prefix_and_suffix("system", [])
                out += f"{p}{self.default_system_prompt}{s}"
            p, s = self.get_prefix_and_suffix(m["role"], messages[:i])
            out += f"{p}{m['content']}{s}"
        return out


_REGISTRY: Dict[str, ChatTemplate] = {}


def register_chat_template(t: ChatTemplate) -> None:
    _REGISTRY[t.name] = t


def get_chat_template(name: str) -> ChatTemplate:
    return _REGISTRY[name]


def get_chat_template_by_model_path(path: str) -> ChatTemplate:
    p = (path or "").lower()
    if "llama-3" in p or "llama3" in p:
        return _REGISTRY["llama-3-instruct"]
    if "llama-2" in p:
        return _REGISTRY["llama-2-chat"]
    if "qwen" in p or "chatml" in p:
        return _REGISTRY["chatml"]
    if "llava" in p or "vicuna" in p:
        return _REGISTRY["vicuna_v1.1"]
    if "mistral" in p or "mixtral" in p:
        return _REGISTRY["mistral"]
    if "gemma" in p:
        return _REGISTRY["gemma-it"]
    return _REGISTRY["default"]


register_chat_template(ChatTemplate("default", None, {"system": ("SYSTEM:", "\n"), "user": ("USER:", "\n"), "assistant": ("ASSISTANT:", "\n")}))
register_chat_template(ChatTemplate("chatml", None, {"system": ("<|im_start|>system\n", "<|im_end|>\n"), "user": ("<|im_start|>user\n", "<|im_end|>\n"),
                                                     "assistant": ("<|im_start|>assistant\n", "<|im_end|>\n")}, ["<|im_end|>"]))
register_chat_template(ChatTemplate("llama-3-instruct", None, {
    "system": ("<|start_header_id|>system<|end_header_id|>\n\n", "<|eot_id|>"), "user": ("<|start_header_id|>user<|end_header_id|>\n\n", "<|eot_id|>"),
    "assistant": ("<|start_header_id|>assistant<|end_header_id|>\n\n", "<|eot_id|>")}, ["<|eot_id|>"]))
register_chat_template(ChatTemplate("llama-2-chat", None, {"system": ("<<SYS>>\n", "\n<</SYS>>\n\n"), "user": ("[INST] ", " [/INST]"),
                                                           "assistant": ("", " </s><s>")}))
register_chat_template(ChatTemplate("vicuna_v1.1", "A chat between a curious user and an artificial intelligence assistant. The assistant gives "
                                    "helpful, detailed, and polite answers to the user's questions.",
                                    {"system": ("", " "), "user": ("USER:", " "), "assistant": ("ASSISTANT:", "</s>")}, image_token=" <image>\n"))
register_chat_template(ChatTemplate("mistral", None, {"system": ("[SYSTEM_PROMPT] ", " [/SYSTEM_PROMPT]"), "user": ("[INST] ", " [/INST]"),
                                                      "assistant": ("", " </s><s>")}, ["</s>"]))
register_chat_template(ChatTemplate("gemma-it", None, {"system": ("", ""), "user": ("<start_of_turn>user\n", "<end_of_turn>\n"),
                                                       "assistant": ("<start_of_turn>model\n", "<end_of_turn>\n")}))
