"""IPC message schema between front end, scheduler and detokenizer.

Parity: reference ``python/sglang/srt/managers/io_struct.py`` (``GenerateReqInput:162``,
``TokenizedGenerateReqInput:944``, ``BatchTokenIDOutput:1397``, ``BatchStrOutput:1497`` and the
control messages).  Messages are plain dataclasses serialised as tagged msgpack maps
(``encode``/``decode``) so the same structs work in-process, over ZMQ and over the C++ ring.
"""
from __future__ import annotations

import dataclasses

import numpy as np
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Union

import msgpack

_TYPES: Dict[str, type] = {}


def message(cls):
    cls = dataclass(cls)
    _TYPES[cls.__name__] = cls
    return cls


def _pack(o: Any) -> Any:
    if dataclasses.is_dataclass(o) and not isinstance(o, type):
        d = {f.name: _pack(getattr(o, f.name)) for f in dataclasses.fields(o)}
        d["__t"] = type(o).__name__
        return d
    if isinstance(o, (list, tuple)):
        return [_pack(x) for x in o]
    if isinstance(o, dict):
        return {k: _pack(v) for k, v in o.items()}
    if hasattr(o, "tolist") and not isinstance(o, (str, bytes)):
        return o.tolist()
    return o


def _unpack(o: Any) -> Any:
    if isinstance(o, dict):
        t = o.pop("__t", None)
        d = {k: _unpack(v) for k, v in o.items()}
        if t is not None:
            return _TYPES[t](**d)
        return d
    if isinstance(o, list):
        return [_unpack(x) for x in o]
    return o


def encode(msg: Any) -> bytes:
    return msgpack.packb(_pack(msg), use_bin_type=True)


def decode(buf: bytes) -> Any:
    return _unpack(msgpack.unpackb(buf, raw=False, strict_map_key=False))


# --------------------------------------------------------------------------- requests
@message
class GenerateReqInput:
    text: Optional[Union[str, List[str]]] = None
    input_ids: Optional[Union[List[int], List[List[int]]]] = None
    input_embeds: Optional[Any] = None
    image_data: Optional[Any] = None
    video_data: Optional[Any] = None
    audio_data: Optional[Any] = None
    sampling_params: Optional[Union[Dict, List[Dict]]] = None
    rid: Optional[Union[str, List[str]]] = None
    return_logprob: Union[bool, List[bool]] = False
    logprob_start_len: Union[int, List[int]] = -1
    top_logprobs_num: Union[int, List[int]] = 0
    token_ids_logprob: Optional[Any] = None
    return_text_in_logprobs: bool = False
    stream: bool = False
    return_hidden_states: bool = False
    lora_path: Optional[Union[str, List[Optional[str]]]] = None
    session_params: Optional[Dict] = None
    custom_logit_processor: Optional[str] = None
    bootstrap_host: Optional[str] = None
    bootstrap_port: Optional[int] = None
    bootstrap_room: Optional[int] = None
    data_parallel_rank: Optional[int] = None
    routing_key: Optional[str] = None
    priority: Optional[int] = None
    extra_key: Optional[str] = None

    def is_batch(self) -> bool:
        if self.text is not None:
            return isinstance(self.text, list)
        if self.input_ids is not None:
            return len(self.input_ids) > 0 and isinstance(self.input_ids[0], list)
        raise ValueError("one of text / input_ids / input_embeds must be given")

    def normalize(self) -> List["GenerateReqInput"]:
        """Expand a batch request (and ``n>1``) into single requests
        (reference ``normalize_batch_and_arguments``)."""
        given = sum(x is not None for x in (self.text, self.input_ids, self.input_embeds))
        if given != 1:
            raise ValueError("exactly one of text / input_ids / input_embeds must be provided")
        if not self.is_batch():
            n = 1
            texts, ids = [self.text], [self.input_ids]
        else:
            src = self.text if self.text is not None else self.input_ids
            n = len(src)
            texts = self.text if self.text is not None else [None] * n
            ids = self.input_ids if self.input_ids is not None else [None] * n

        def pick(v, i):
            return v[i] if isinstance(v, list) else v

        out = []
        for i in range(n):
            sp = self.sampling_params
            sp_i = dict(sp[i] if isinstance(sp, list) else (sp or {}))
            par = int(sp_i.get("n", 1) or 1)
            rid = pick(self.rid, i) if self.rid is not None else None
            img = self.image_data
            if self.is_batch() and isinstance(img, list) and len(img) == n:
                img = img[i]
            aud = self.audio_data
            if self.is_batch() and isinstance(aud, list) and len(aud) == n and not np.isscalar(aud[0]):
                aud = aud[i]
            for j in range(par):
                sp_ij = dict(sp_i)
                sp_ij["n"] = 1
                r = dataclasses.replace(
                    self, text=texts[i], input_ids=ids[i], sampling_params=sp_ij, image_data=img, audio_data=aud,
                    rid=(rid if (rid and par == 1) else (f"{rid}-{j}" if rid else None)),     # None: the engine assigns (SPMD-consistent) ids
                    return_logprob=pick(self.return_logprob, i), logprob_start_len=pick(self.logprob_start_len, i),
                    top_logprobs_num=pick(self.top_logprobs_num, i), lora_path=pick(self.lora_path, i),
                )
                out.append(r)
        return out


@message
class EmbeddingReqInput:
    text: Optional[Union[str, List[str]]] = None
    input_ids: Optional[Any] = None
    rid: Optional[Union[str, List[str]]] = None
    image_data: Optional[Any] = None
    dimensions: Optional[int] = None


@message
class TokenizedGenerateReqInput:
    rid: str
    input_text: Optional[str]
    input_ids: List[int]
    sampling_params: Dict[str, Any]
    return_logprob: bool = False
    logprob_start_len: int = -1
    top_logprobs_num: int = 0
    token_ids_logprob: Optional[List[int]] = None
    stream: bool = False
    return_hidden_states: bool = False
    lora_id: Optional[str] = None
    session_params: Optional[Dict] = None
    mm_inputs: Optional[Dict[str, Any]] = None
    custom_logit_processor: Optional[str] = None
    bootstrap_host: Optional[str] = None
    bootstrap_port: Optional[int] = None
    bootstrap_room: Optional[int] = None
    priority: Optional[int] = None
    extra_key: Optional[str] = None
    is_embedding: bool = False
    created_time: float = 0.0
    trace_context: Optional[Dict[str, str]] = None


@message
class BatchTokenIDOutput:
    rids: List[str]
    finished_reasons: List[Optional[Dict[str, Any]]]
    output_ids: List[List[int]]          # NEW token ids since the last message (incremental)
    read_offsets: List[int]
    prompt_tokens: List[int]
    completion_tokens: List[int]
    cached_tokens: List[int]
    skip_special_tokens: List[bool]
    no_stop_trim: List[bool]
    stop_strs: List[List[str]]
    output_logprobs: Optional[List[Any]] = None
    input_logprobs: Optional[List[Any]] = None
    top_logprobs: Optional[List[Any]] = None
    spec_verify_ct: Optional[List[int]] = None
    spec_accepted_tokens: Optional[List[int]] = None
    embeddings: Optional[List[Any]] = None
    hidden_states: Optional[List[Any]] = None
    ttfts: Optional[List[Optional[float]]] = None      # scheduler-measured time to first token (s)
    queue_times: Optional[List[Optional[float]]] = None


@message
class BatchStrOutput:
    rids: List[str]
    finished_reasons: List[Optional[Dict[str, Any]]]
    output_strs: List[str]               # incremental text
    output_ids: List[List[int]]
    prompt_tokens: List[int]
    completion_tokens: List[int]
    cached_tokens: List[int]
    output_logprobs: Optional[List[Any]] = None
    input_logprobs: Optional[List[Any]] = None
    top_logprobs: Optional[List[Any]] = None
    spec_verify_ct: Optional[List[int]] = None
    spec_accepted_tokens: Optional[List[int]] = None
    embeddings: Optional[List[Any]] = None
    hidden_states: Optional[List[Any]] = None
    ttfts: Optional[List[Optional[float]]] = None
    queue_times: Optional[List[Optional[float]]] = None


# --------------------------------------------------------------------------- control plane
@message
class AbortReq:
    rid: str = ""
    abort_all: bool = False


@message
class FlushCacheReq:
    pass


@message
class FlushCacheResp:
    success: bool = True


@message
class GetInternalStateReq:
    pass


@message
class GetInternalStateResp:
    state: Dict[str, Any] = field(default_factory=dict)


@message
class ProfileReq:
    action: str = "start"  # start|stop
    output_dir: Optional[str] = None
    num_steps: Optional[int] = None
    activities: Optional[List[str]] = None
    profile_by_stage: bool = False


@message
class ProfileResp:
    success: bool = True
    message: str = ""


@message
class UpdateWeightsFromDiskReq:
    model_path: str = ""
    load_format: Optional[str] = None


@message
class UpdateWeightsFromTensorReq:
    named_tensors: Any = None  # list[(name, nested list)] or serialized handle
    flush_cache: bool = True


@message
class InitWeightsUpdateGroupReq:
    """join a dedicated process group with an external trainer (rank 0) — reference engine.init_weights_update_group"""
    master_address: str = "127.0.0.1"
    master_port: int = 0
    rank_offset: int = 1
    world_size: int = 2
    group_name: str = "weight_update_group"
    backend: str = "nccl"


@message
class DestroyWeightsUpdateGroupReq:
    group_name: str = "weight_update_group"


@message
class UpdateWeightsFromDistributedReq:
    names: Any = None
    dtypes: Any = None
    shapes: Any = None
    group_name: str = "weight_update_group"
    flush_cache: bool = True


@message
class UpdateWeightsResp:
    success: bool = True
    message: str = ""


@message
class GetWeightsByNameReq:
    name: str = ""
    truncate_size: int = 100


@message
class RpcReq:
    method: str = ""
    kwargs: Optional[Dict[str, Any]] = None


@message
class RpcResp:
    success: bool = True
    result: Any = None
    message: str = ""


@message
class SendWeightsToRemoteInstanceReq:
    """seed side of the remote-instance loader: join a 2-rank group per TP rank (the new instance hosts the store at
    ``master_address:master_port + tp_rank``) and broadcast every local parameter shard"""
    master_address: str = "127.0.0.1"
    master_port: int = 0
    group_name: str = "remote_instance_load"
    backend: str = "nccl"


@message
class WeightsCheckerReq:
    """checksum every parameter shard (reference /weights_checker, utils/weight_checker.py): compare instances / ranks / before-after a swap"""
    action: str = "checksum"       # checksum | snapshot | compare (against the stored snapshot)


@message
class SlowDownReq:
    """fault injection (reference /slow_down): sleep this long before every forward; None / 0 clears"""
    forward_sleep_time: Optional[float] = None


@message
class GetWeightsByNameResp:
    parameter: Any = None


@message
class ReleaseMemoryOccupationReq:
    tags: Optional[List[str]] = None


@message
class ResumeMemoryOccupationReq:
    tags: Optional[List[str]] = None


@message
class GenericResp:
    success: bool = True
    message: str = ""
    payload: Any = None


@message
class PauseGenerationReq:
    mode: str = "abort"  # abort|retract|in_place


@message
class ContinueGenerationReq:
    pass


@message
class OpenSessionReq:
    capacity_of_str_len: int = 0
    session_id: Optional[str] = None


@message
class CloseSessionReq:
    session_id: str = ""


@message
class OpenSessionResp:
    session_id: Optional[str] = None
    success: bool = True


@message
class LoadLoRAAdapterReq:
    lora_name: str = ""
    lora_path: str = ""


@message
class UnloadLoRAAdapterReq:
    lora_name: str = ""


@message
class SetTraceLevelReq:
    level: int = 1


@message
class HealthCheckReq:
    pass


@message
class HealthCheckResp:
    ok: bool = True
    forward_ct: int = 0


@message
class GetLoadReq:
    pass


@message
class GetKvEventsReq:
    """drain the radix cache's KV events (BlockStored / BlockRemoved / AllBlocksCleared with chained page hashes)"""
    pass


@message
class GetKvEventsResp:
    events: Any = None
    page_size: int = 1


@message
class GetLoadResp:
    num_running_reqs: int = 0
    num_waiting_reqs: int = 0
    num_tokens: int = 0
    token_usage: float = 0.0


@message
class ExpertDistributionReq:
    action: str = "start"  # start|stop|dump


@message
class ShutdownReq:
    pass
