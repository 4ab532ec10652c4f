"""OpenAI-compatible request/response schema (pydantic).  Parity: reference
``python/sglang/srt/entrypoints/openai/protocol.py`` (the fields clients actually send; SGLang extensions
``regex``, ``ebnf``, ``json_schema`` via response_format, ``top_k``, ``min_p``, ``ignore_eos``, ``lora_path``...)."""
from __future__ import annotations

import time
import uuid
from typing import Any, Dict, List, Literal, Optional, Union

from pydantic import BaseModel, Field


class ModelCard(BaseModel):
    id: str
    object: str = "model"
    created: int = Field(default_factory=lambda: int(time.time()))
    owned_by: str = "sglang_b200"
    root: Optional[str] = None
    max_model_len: Optional[int] = None


class ModelList(BaseModel):
    object: str = "list"
    data: List[ModelCard] = []


class UsageInfo(BaseModel):
    prompt_tokens: int = 0
    total_tokens: int = 0
    completion_tokens: Optional[int] = 0
    prompt_tokens_details: Optional[Dict[str, int]] = None


class StreamOptions(BaseModel):
    include_usage: Optional[bool] = False


class JsonSchemaResponseFormat(BaseModel):
    name: str = "schema"
    description: Optional[str] = None
    schema_: Optional[Dict[str, Any]] = Field(default=None, alias="schema")
    strict: Optional[bool] = False


class ResponseFormat(BaseModel):
    type: Literal["text", "json_object", "json_schema", "regex", "ebnf"] = "text"
    json_schema: Optional[JsonSchemaResponseFormat] = None
    pattern: Optional[str] = None
    grammar: Optional[str] = None


class _SamplingMixin(BaseModel):
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    min_p: float = 0.0
    n: int = 1
    stop: Optional[Union[str, List[str]]] = None
    stop_token_ids: Optional[List[int]] = None
    presence_penalty: float = 0.0
    frequency_penalty: float = 0.0
    repetition_penalty: float = 1.0
    seed: Optional[int] = None
    logit_bias: Optional[Dict[str, float]] = None
    ignore_eos: bool = False
    skip_special_tokens: bool = True
    no_stop_trim: bool = False
    min_tokens: int = 0
    regex: Optional[str] = None
    ebnf: Optional[str] = None
    json_schema: Optional[str] = None
    lora_path: Optional[str] = None
    session_params: Optional[Dict] = None
    priority: Optional[int] = None
    rid: Optional[str] = None

    def base_sampling(self, max_tokens: Optional[int]) -> Dict[str, Any]:
        d: Dict[str, Any] = {"temperature": self.temperature, "top_p": self.top_p, "top_k": self.top_k, "min_p": self.min_p,
                             "stop": self.stop, "stop_token_ids": self.stop_token_ids, "presence_penalty": self.presence_penalty,
                             "frequency_penalty": self.frequency_penalty, "repetition_penalty": self.repetition_penalty,
                             "ignore_eos": self.ignore_eos, "skip_special_tokens": self.skip_special_tokens,
                             "no_stop_trim": self.no_stop_trim, "min_new_tokens": self.min_tokens, "n": self.n,
                             "logit_bias": self.logit_bias, "sampling_seed": self.seed}
        if max_tokens is not None:
            d["max_new_tokens"] = max_tokens
        for k in ("regex", "ebnf", "json_schema"):
            if getattr(self, k) is not None:
                d[k] = getattr(self, k)
        return d


class CompletionRequest(_SamplingMixin):
    model: str = "default"
    prompt: Union[str, List[str], List[int], List[List[int]]]
    max_tokens: Optional[int] = 16
    echo: bool = False
    logprobs: Optional[int] = None
    stream: bool = False
    stream_options: Optional[StreamOptions] = None
    suffix: Optional[str] = None
    user: Optional[str] = None
    response_format: Optional[ResponseFormat] = None


class LogProbs(BaseModel):
    text_offset: List[int] = []
    token_logprobs: List[Optional[float]] = []
    tokens: List[str] = []
    top_logprobs: List[Optional[Dict[str, float]]] = []


class CompletionChoice(BaseModel):
    index: int
    text: str
    logprobs: Optional[LogProbs] = None
    finish_reason: Optional[str] = None
    matched_stop: Union[None, int, str] = None


class CompletionResponse(BaseModel):
    id: str
    object: str = "text_completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    choices: List[CompletionChoice]
    usage: UsageInfo


class FunctionDef(BaseModel):
    name: str
    description: Optional[str] = None
    parameters: Optional[Dict[str, Any]] = None
    strict: Optional[bool] = False


class Tool(BaseModel):
    type: str = "function"
    function: FunctionDef


class ToolChoiceFunc(BaseModel):
    name: str


class ToolChoice(BaseModel):
    type: str = "function"
    function: ToolChoiceFunc


class FunctionCall(BaseModel):
    name: Optional[str] = None
    arguments: Optional[str] = None


class ToolCall(BaseModel):
    id: Optional[str] = None
    index: Optional[int] = None
    type: str = "function"
    function: FunctionCall


class ChatMessage(BaseModel):
    role: str
    content: Optional[Union[str, List[Dict[str, Any]]]] = None
    name: Optional[str] = None
    tool_calls: Optional[List[ToolCall]] = None
    tool_call_id: Optional[str] = None
    reasoning_content: Optional[str] = None


class ChatCompletionRequest(_SamplingMixin):
    model: str = "default"
    messages: List[ChatMessage]
    max_tokens: Optional[int] = None
    max_completion_tokens: Optional[int] = None
    logprobs: bool = False
    top_logprobs: Optional[int] = None
    stream: bool = False
    stream_options: Optional[StreamOptions] = None
    response_format: Optional[ResponseFormat] = None
    tools: Optional[List[Tool]] = None
    tool_choice: Union[str, ToolChoice] = "auto"
    chat_template_kwargs: Optional[Dict[str, Any]] = None
    separate_reasoning: bool = True
    stream_reasoning: bool = True
    continue_final_message: bool = False
    user: Optional[str] = None


class ChatCompletionChoice(BaseModel):
    index: int
    message: ChatMessage
    logprobs: Optional[Any] = None
    finish_reason: Optional[str] = None
    matched_stop: Union[None, int, str] = None


class ChatCompletionResponse(BaseModel):
    id: str
    object: str = "chat.completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    choices: List[ChatCompletionChoice]
    usage: UsageInfo


class DeltaMessage(BaseModel):
    role: Optional[str] = None
    content: Optional[str] = None
    reasoning_content: Optional[str] = None
    tool_calls: Optional[List[ToolCall]] = None


class ChatCompletionStreamChoice(BaseModel):
    index: int
    delta: DeltaMessage
    logprobs: Optional[Any] = None
    finish_reason: Optional[str] = None
    matched_stop: Union[None, int, str] = None


class ChatCompletionStreamResponse(BaseModel):
    id: str
    object: str = "chat.completion.chunk"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: str
    choices: List[ChatCompletionStreamChoice]
    usage: Optional[UsageInfo] = None


class EmbeddingRequest(BaseModel):
    model: str = "default"
    input: Union[str, List[str], List[int], List[List[int]]]
    encoding_format: str = "float"
    dimensions: Optional[int] = None
    user: Optional[str] = None


class EmbeddingObject(BaseModel):
    embedding: List[float]
    index: int
    object: str = "embedding"


class EmbeddingResponse(BaseModel):
    data: List[EmbeddingObject]
    model: str
    object: str = "list"
    usage: UsageInfo


class TokenizeRequest(BaseModel):
    model: str = "default"
    prompt: Union[str, List[str]]
    add_special_tokens: bool = True


class DetokenizeRequest(BaseModel):
    model: str = "default"
    tokens: Union[List[int], List[List[int]]]
    skip_special_tokens: bool = True


class ScoringRequest(BaseModel):
    model: str = "default"
    query: str
    items: List[str]
    label_token_ids: List[int]
    apply_softmax: bool = True


class RerankRequest(BaseModel):
    model: str = "default"
    query: str
    documents: List[str]
    top_n: Optional[int] = None


class ErrorResponse(BaseModel):
    object: str = "error"
    message: str
    type: str = "BadRequestError"
    param: Optional[str] = None
    code: int = 400


def new_id(prefix: str) -> str:
    return f"{prefix}-{uuid.uuid4().hex}"
