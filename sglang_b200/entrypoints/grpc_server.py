"""gRPC front end (``sglang_b200.runtime.v1.SglbService``): typed native RPCs + JSON pass-through OpenAI RPCs + admin RPCs.

Parity: reference ``proto/sglang/runtime/v1/*.proto`` (``SglangService``: TextGenerate/Generate streams, Embed, Tokenize,
HealthCheck, GetModelInfo/GetServerInfo/GetLoad, Abort, FlushCache, Pause/Continue, ChatComplete/Complete JSON streams,
StartProfile/StopProfile, UpdateWeightsFromDisk), ``rust/sglang-grpc`` (tonic server + Python bridge) and
``python/sglang/srt/entrypoints/grpc_server.py:263``.

There is no ``protoc`` in this image, so the message schema is declared ONCE below (``SCHEMA``) and used twice: to build real
protobuf descriptors at import time (``descriptor_pb2`` -> message classes, wire-compatible with any protoc-generated client) and to
emit the ``.proto`` text (``python -m sglang_b200.entrypoints.grpc_server --emit-proto`` -> ``proto/sglb/runtime/v1/sglb.proto``).
The server uses grpc's generic handlers, one thread pool; generation RPCs bridge into the engine's async generator.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import logging
import threading
from concurrent import futures
from typing import Any, Dict, Iterator, List, Optional, Tuple

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

logger = logging.getLogger(__name__)
PACKAGE = "sglang_b200.runtime.v1"
SERVICE = "SglbService"

# field spec: (name, type, number[, "repeated"|"optional"])
SCHEMA: Dict[str, List[Tuple]] = {
    "DisaggregatedParams": [("bootstrap_host", "string", 1), ("bootstrap_port", "int32", 2), ("bootstrap_room", "int64", 3)],
    "SamplingParams": [("temperature", "float", 1, "optional"), ("top_p", "float", 2, "optional"), ("top_k", "int32", 3, "optional"),
                       ("min_p", "float", 4, "optional"), ("frequency_penalty", "float", 5, "optional"),
                       ("presence_penalty", "float", 6, "optional"), ("repetition_penalty", "float", 7, "optional"),
                       ("max_new_tokens", "int32", 8, "optional"), ("min_new_tokens", "int32", 9, "optional"),
                       ("stop", "string", 10, "repeated"), ("stop_token_ids", "int32", 11, "repeated"), ("ignore_eos", "bool", 12),
                       ("skip_special_tokens", "bool", 13, "optional"), ("regex", "string", 14, "optional"),
                       ("json_schema", "string", 15, "optional"), ("ebnf", "string", 16, "optional"), ("n", "int32", 17, "optional"),
                       ("sampling_seed", "int64", 18, "optional")],
    "TextGenerateRequest": [("text", "string", 1), ("sampling_params", "SamplingParams", 2), ("stream", "bool", 3), ("rid", "string", 4),
                            ("return_logprob", "bool", 5), ("top_logprobs_num", "int32", 6), ("lora_path", "string", 7),
                            ("disaggregated_params", "DisaggregatedParams", 8), ("priority", "int32", 9, "optional")],
    "GenerateRequest": [("input_ids", "int32", 1, "repeated"), ("sampling_params", "SamplingParams", 2), ("stream", "bool", 3),
                        ("rid", "string", 4), ("return_logprob", "bool", 5), ("top_logprobs_num", "int32", 6), ("lora_path", "string", 7),
                        ("disaggregated_params", "DisaggregatedParams", 8), ("priority", "int32", 9, "optional")],
    "GenerateMeta": [("rid", "string", 1), ("prompt_tokens", "int32", 2), ("completion_tokens", "int32", 3), ("cached_tokens", "int32", 4),
                     ("finish_reason", "string", 5), ("matched_stop", "string", 6), ("e2e_latency", "double", 7),
                     ("output_token_logprobs", "float", 8, "repeated")],
    "TextGenerateResponse": [("text", "string", 1), ("output_ids", "int32", 2, "repeated"), ("finished", "bool", 3), ("meta", "GenerateMeta", 4)],
    "GenerateResponse": [("output_ids", "int32", 1, "repeated"), ("finished", "bool", 2), ("meta", "GenerateMeta", 3)],
    "TextEmbedRequest": [("text", "string", 1, "repeated")],
    "EmbedRequest": [("input_ids", "int32", 1, "repeated")],
    "Embedding": [("values", "float", 1, "repeated"), ("prompt_tokens", "int32", 2)],
    "EmbedResponse": [("embeddings", "Embedding", 1, "repeated")],
    "TokenizeRequest": [("text", "string", 1), ("add_special_tokens", "bool", 2)],
    "TokenizeResponse": [("tokens", "int32", 1, "repeated"), ("count", "int32", 2)],
    "DetokenizeRequest": [("tokens", "int32", 1, "repeated"), ("skip_special_tokens", "bool", 2)],
    "DetokenizeResponse": [("text", "string", 1)],
    "Empty": [],
    "HealthCheckResponse": [("healthy", "bool", 1), ("message", "string", 2)],
    "JsonResponse": [("json", "string", 1)],
    "ListModelsResponse": [("models", "string", 1, "repeated")],
    "GetLoadResponse": [("num_running_reqs", "int32", 1), ("num_waiting_reqs", "int32", 2), ("num_used_tokens", "int64", 3),
                        ("token_usage", "float", 4)],
    "AbortRequest": [("rid", "string", 1), ("abort_all", "bool", 2)],
    "StatusResponse": [("success", "bool", 1), ("message", "string", 2)],
    "OpenAIRequest": [("json", "string", 1)],
    "OpenAIStreamChunk": [("json", "string", 1), ("done", "bool", 2)],
    "StartProfileRequest": [("output_dir", "string", 1), ("num_steps", "int32", 2, "optional")],
    "UpdateWeightsRequest": [("model_path", "string", 1), ("load_format", "string", 2)],
}
# rpc name -> (request, response, server_streaming)
RPCS: Dict[str, Tuple[str, str, bool]] = {
    "TextGenerate": ("TextGenerateRequest", "TextGenerateResponse", True), "Generate": ("GenerateRequest", "GenerateResponse", True),
    "TextEmbed": ("TextEmbedRequest", "EmbedResponse", False), "Embed": ("EmbedRequest", "EmbedResponse", False),
    "Tokenize": ("TokenizeRequest", "TokenizeResponse", False), "Detokenize": ("DetokenizeRequest", "DetokenizeResponse", False),
    "HealthCheck": ("Empty", "HealthCheckResponse", False), "GetModelInfo": ("Empty", "JsonResponse", False),
    "GetServerInfo": ("Empty", "JsonResponse", False), "ListModels": ("Empty", "ListModelsResponse", False),
    "GetLoad": ("Empty", "GetLoadResponse", False), "Abort": ("AbortRequest", "StatusResponse", False),
    "FlushCache": ("Empty", "StatusResponse", False), "PauseGeneration": ("Empty", "StatusResponse", False),
    "ContinueGeneration": ("Empty", "StatusResponse", False), "ChatComplete": ("OpenAIRequest", "OpenAIStreamChunk", True),
    "Complete": ("OpenAIRequest", "OpenAIStreamChunk", True), "OpenAIEmbed": ("OpenAIRequest", "JsonResponse", False),
    "Score": ("OpenAIRequest", "JsonResponse", False), "Rerank": ("OpenAIRequest", "JsonResponse", False),
    "StartProfile": ("StartProfileRequest", "StatusResponse", False), "StopProfile": ("Empty", "StatusResponse", False),
    "UpdateWeightsFromDisk": ("UpdateWeightsRequest", "StatusResponse", False),
}
_SCALARS = {"string": 9, "int32": 5, "int64": 3, "float": 2, "double": 1, "bool": 8}


def _build_messages() -> Dict[str, Any]:
    fd = descriptor_pb2.FileDescriptorProto(name="sglb/runtime/v1/sglb.proto", package=PACKAGE, syntax="proto3")
    for mname, fields in SCHEMA.items():
        m = fd.message_type.add(name=mname)
        for f in fields:
            name, typ, num = f[0], f[1], f[2]
            mod = f[3] if len(f) > 3 else None
            fld = m.field.add(name=name, number=num)
            fld.label = 3 if mod == "repeated" else 1
            if typ in _SCALARS:
                fld.type = _SCALARS[typ]
            else:
                fld.type = 11
                fld.type_name = f".{PACKAGE}.{typ}"
            if mod == "optional":
                fld.proto3_optional = True
                m.oneof_decl.add(name="_" + name)
                fld.oneof_index = len(m.oneof_decl) - 1
    svc = fd.service.add(name=SERVICE)
    for rname, (req, resp, stream) in RPCS.items():
        svc.method.add(name=rname, input_type=f".{PACKAGE}.{req}", output_type=f".{PACKAGE}.{resp}", server_streaming=stream)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{PACKAGE}.{n}")) for n in SCHEMA}


MSG = _build_messages()


def emit_proto() -> str:
    out = ['syntax = "proto3";', f"package {PACKAGE};", "", f"service {SERVICE} {{"]
    for r, (a, b, s) in RPCS.items():
        out.append(f"  rpc {r}({a}) returns ({'stream ' if s else ''}{b});")
    out.append("}")
    for m, fields in SCHEMA.items():
        out += ["", f"message {m} {{"]
        for f in fields:
            mod = (f[3] + " ") if len(f) > 3 else ""
            out.append(f"  {mod}{f[1]} {f[0]} = {f[2]};")
        out.append("}")
    return "\n".join(out) + "\n"


def _sampling_dict(sp: Any) -> Dict[str, Any]:
    d: Dict[str, Any] = {}
    for name, _, _, *mod in SCHEMA["SamplingParams"]:
        m = mod[0] if mod else None
        if m == "optional":
            if sp.HasField(name):
                d[name] = getattr(sp, name)
        elif m == "repeated":
            v = list(getattr(sp, name))
            if v:
                d[name] = v
        elif getattr(sp, name):
            d[name] = getattr(sp, name)
    return d


class _Servicer:
    def __init__(self, engine: Any, serving: Any = None):
        self.engine = engine
        if serving is None:
            from .openai.serving import OpenAIServing

            serving = OpenAIServing(engine)
        self.serving = serving
        self.loop = asyncio.new_event_loop()
        threading.Thread(target=self.loop.run_forever, daemon=True, name="sglb-grpc-loop").start()

    # ---- helpers
    def _run(self, coro):
        return asyncio.run_coroutine_threadsafe(coro, self.loop).result()

    def _aiter(self, agen) -> Iterator[Any]:
        """drain an async generator from a sync gRPC handler thread"""
        while True:
            try:
                yield asyncio.run_coroutine_threadsafe(agen.__anext__(), self.loop).result()
            except StopAsyncIteration:
                return

    @staticmethod
    def _meta(m: Dict[str, Any]) -> Any:
        fr = m.get("finish_reason") or {}
        meta = MSG["GenerateMeta"](rid=str(m.get("id", "")), prompt_tokens=int(m.get("prompt_tokens", 0)),
                                   completion_tokens=int(m.get("completion_tokens", 0)), cached_tokens=int(m.get("cached_tokens", 0)),
                                   finish_reason=str(fr.get("type", "")) if fr else "", matched_stop=str(fr.get("matched", "")) if fr and fr.get("matched") is not None else "",
                                   e2e_latency=float(m.get("e2e_latency", 0.0)))
        for t in m.get("output_token_logprobs") or []:
            meta.output_token_logprobs.append(float(t[0]))
        return meta

    def _gen_kwargs(self, r: Any) -> Dict[str, Any]:
        kw: Dict[str, Any] = dict(sampling_params=_sampling_dict(r.sampling_params), return_logprob=r.return_logprob,
                                  top_logprobs_num=r.top_logprobs_num)
        if r.rid:
            kw["rid"] = r.rid
        if r.lora_path:
            kw["lora_path"] = r.lora_path
        if r.HasField("priority"):
            kw["priority"] = r.priority
        if r.HasField("disaggregated_params") and r.disaggregated_params.bootstrap_host:
            d = r.disaggregated_params
            kw.update(bootstrap_host=d.bootstrap_host, bootstrap_port=d.bootstrap_port, bootstrap_room=d.bootstrap_room)
        return kw

    def _generate(self, prompt: Any, input_ids: Any, r: Any, text_out: bool) -> Iterator[Any]:
        kw = self._gen_kwargs(r)
        Resp = MSG["TextGenerateResponse"] if text_out else MSG["GenerateResponse"]

        def mk(o: Dict[str, Any], ids: List[int], text: str):
            fin = o["meta_info"].get("finish_reason") is not None
            return (Resp(text=text, output_ids=ids, finished=fin, meta=self._meta(o["meta_info"])) if text_out
                    else Resp(output_ids=ids, finished=fin, meta=self._meta(o["meta_info"])))

        if r.stream:
            async def start():
                return await self.engine.async_generate(prompt, input_ids=input_ids, stream=True, **kw)
            agen = self._run(start())
            sent_t, sent_i = 0, 0
            for o in self._aiter(agen):
                ids, text = o["output_ids"][sent_i:], (o.get("text") or "")[sent_t:]
                sent_i, sent_t = len(o["output_ids"]), len(o.get("text") or "")
                yield mk(o, ids, text)
        else:
            o = self._run(self.engine.async_generate(prompt, input_ids=input_ids, **kw))
            yield mk(o, o["output_ids"], o.get("text") or "")

    # ---- native RPCs
    def TextGenerate(self, r, ctx):
        yield from self._generate(r.text, None, r, True)

    def Generate(self, r, ctx):
        yield from self._generate(None, list(r.input_ids), r, False)

    def TextEmbed(self, r, ctx):
        outs = self.engine.encode(list(r.text))
        return MSG["EmbedResponse"](embeddings=[MSG["Embedding"](values=o["embedding"], prompt_tokens=o["meta_info"].get("prompt_tokens", 0)) for o in outs])

    def Embed(self, r, ctx):
        outs = self.engine.encode([self.engine.tokenizer.decode(list(r.input_ids))])
        return MSG["EmbedResponse"](embeddings=[MSG["Embedding"](values=o["embedding"], prompt_tokens=o["meta_info"].get("prompt_tokens", 0)) for o in outs])

    def Tokenize(self, r, ctx):
        ids = self.engine.tokenizer.encode(r.text)
        return MSG["TokenizeResponse"](tokens=ids, count=len(ids))

    def Detokenize(self, r, ctx):
        return MSG["DetokenizeResponse"](text=self.engine.tokenizer.decode(list(r.tokens)))

    def HealthCheck(self, r, ctx):
        return MSG["HealthCheckResponse"](healthy=True, message="ok")

    def GetModelInfo(self, r, ctx):
        a = self.engine.server_args
        return MSG["JsonResponse"](json=json.dumps({"model_path": a.model_path, "tokenizer_path": a.tokenizer_path,
                                                     "is_generation": not a.is_embedding, "architecture": self.engine.model_config.architecture,
                                                     "weight_version": self.engine.weight_version}))

    def GetServerInfo(self, r, ctx):
        return MSG["JsonResponse"](json=json.dumps(self.engine.get_server_info(), default=str))

    def ListModels(self, r, ctx):
        return MSG["ListModelsResponse"](models=[self.engine.server_args.served_model_name])

    def GetLoad(self, r, ctx):
        ld = self.engine.get_load()
        return MSG["GetLoadResponse"](num_running_reqs=int(ld.get("num_running_reqs", 0)), num_waiting_reqs=int(ld.get("num_waiting_reqs", 0)),
                                      num_used_tokens=int(ld.get("num_used_tokens", 0)), token_usage=float(ld.get("token_usage", 0.0)))

    def Abort(self, r, ctx):
        self.engine.abort_request(r.rid, abort_all=r.abort_all)
        return MSG["StatusResponse"](success=True)

    def FlushCache(self, r, ctx):
        ok = self.engine.flush_cache()
        return MSG["StatusResponse"](success=bool(ok) if ok is not None else True)

    def PauseGeneration(self, r, ctx):
        self.engine.pause_generation()
        return MSG["StatusResponse"](success=True)

    def ContinueGeneration(self, r, ctx):
        self.engine.continue_generation()
        return MSG["StatusResponse"](success=True)

    # ---- OpenAI JSON pass-through
    def _openai_stream(self, r, chat: bool):
        from .openai import protocol as P

        body = json.loads(r.json)
        req = (P.ChatCompletionRequest if chat else P.CompletionRequest)(**body)
        out = self._run(self.serving.chat_completions(req) if chat else self.serving.completions(req))
        if req.stream:
            for line in self._aiter(out):
                if line.startswith("data: "):
                    payload = line[6:].strip()
                    if payload == "[DONE]":
                        yield MSG["OpenAIStreamChunk"](json="", done=True)
                    else:
                        yield MSG["OpenAIStreamChunk"](json=payload, done=False)
        else:
            yield MSG["OpenAIStreamChunk"](json=out.model_dump_json(), done=True)

    def ChatComplete(self, r, ctx):
        yield from self._openai_stream(r, True)

    def Complete(self, r, ctx):
        yield from self._openai_stream(r, False)

    def OpenAIEmbed(self, r, ctx):
        from .openai import protocol as P

        return MSG["JsonResponse"](json=self._run(self.serving.embeddings(P.EmbeddingRequest(**json.loads(r.json)))).model_dump_json())

    def Score(self, r, ctx):
        b = json.loads(r.json)
        s = self.engine.score(b["query"], b["items"], b.get("label_token_ids"), b.get("apply_softmax", False))
        return MSG["JsonResponse"](json=json.dumps({"scores": s, "object": "scoring"}))

    def Rerank(self, r, ctx):
        b = json.loads(r.json)
        return MSG["JsonResponse"](json=json.dumps(self.engine.rerank(b["query"], b["documents"])))

    # ---- admin
    def StartProfile(self, r, ctx):
        self.engine.start_profile(r.output_dir or None, r.num_steps if r.HasField("num_steps") else None)
        return MSG["StatusResponse"](success=True)

    def StopProfile(self, r, ctx):
        self.engine.stop_profile()
        return MSG["StatusResponse"](success=True)

    def UpdateWeightsFromDisk(self, r, ctx):
        res = self.engine.update_weights_from_disk(r.model_path, r.load_format or None)
        return MSG["StatusResponse"](success=bool(res.success), message=str(getattr(res, "message", "")))


def make_server(engine: Any, host: str = "127.0.0.1", port: int = 0, max_workers: int = 32) -> Tuple[grpc.Server, int]:
    sv = _Servicer(engine)
    handlers = {}
    for name, (req, resp, stream) in RPCS.items():
        fn = getattr(sv, name)

        def wrap(f):
            def call(request, context):
                try:
                    return f(request, context)
                except Exception as e:  # surface engine errors as gRPC status instead of killing the stream silently
                    context.abort(grpc.StatusCode.INVALID_ARGUMENT if isinstance(e, (ValueError, KeyError)) else grpc.StatusCode.INTERNAL,
                                  f"{type(e).__name__}: {e}")
            return call

        def wrap_stream(f):
            def call(request, context):
                try:
                    yield from f(request, context)
                except Exception as e:
                    context.abort(grpc.StatusCode.INVALID_ARGUMENT if isinstance(e, (ValueError, KeyError)) else grpc.StatusCode.INTERNAL,
                                  f"{type(e).__name__}: {e}")
            return call

        mk = grpc.unary_stream_rpc_method_handler if stream else grpc.unary_unary_rpc_method_handler
        handlers[name] = mk(wrap_stream(fn) if stream else wrap(fn), request_deserializer=MSG[req].FromString,
                            response_serializer=MSG[resp].SerializeToString)
    sa = getattr(engine, "server_args", None)
    api_key, admin_key = getattr(sa, "api_key", None), getattr(sa, "admin_api_key", None)
    interceptors = [_AuthInterceptor(api_key, admin_key)] if (api_key or admin_key) else []
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers), interceptors=interceptors)
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(f"{PACKAGE}.{SERVICE}", handlers),))
    cert, pkey = getattr(sa, "ssl_certfile", None), getattr(sa, "ssl_keyfile", None)
    if cert and pkey:
        with open(pkey, "rb") as fk, open(cert, "rb") as fc:
            creds = grpc.ssl_server_credentials(((fk.read(), fc.read()),))
        bound = server.add_secure_port(f"{host}:{port}", creds)
    else:
        bound = server.add_insecure_port(f"{host}:{port}")
    return server, bound


class _AuthInterceptor(grpc.ServerInterceptor):
    """same policy table as the HTTP middleware (utils/auth.py): bearer token in the ``authorization`` metadata entry"""

    def __init__(self, api_key: Optional[str], admin_key: Optional[str]):
        self.api_key, self.admin_key = api_key, admin_key

    def intercept_service(self, continuation, handler_call_details):
        from ..utils import auth as _auth

        md = dict(handler_call_details.invocation_metadata or ())
        ok, msg = _auth.decide_grpc(handler_call_details.method, md.get("authorization"), self.api_key, self.admin_key)
        if ok:
            return continuation(handler_call_details)
        handler = continuation(handler_call_details)

        def deny(request, context):
            context.abort(grpc.StatusCode.UNAUTHENTICATED, msg or "invalid API key")

        if handler is not None and handler.response_streaming:
            def deny_stream(request, context):
                context.abort(grpc.StatusCode.UNAUTHENTICATED, msg or "invalid API key")
                yield  # pragma: no cover
            return grpc.unary_stream_rpc_method_handler(deny_stream, request_deserializer=handler.request_deserializer,
                                                        response_serializer=handler.response_serializer)
        return grpc.unary_unary_rpc_method_handler(deny, request_deserializer=getattr(handler, "request_deserializer", None),
                                                   response_serializer=getattr(handler, "response_serializer", None))


class SglbGrpcClient:
    """minimal typed client (tests, router, bench): ``client.call("HealthCheck", MSG["Empty"]())``"""

    def __init__(self, target: str, api_key: Optional[str] = None):
        self.channel = grpc.insecure_channel(target)
        self.metadata = (("authorization", f"Bearer {api_key}"),) if api_key else None

    def call(self, rpc: str, request: Any):
        req, resp, stream = RPCS[rpc]
        path = f"/{PACKAGE}.{SERVICE}/{rpc}"
        mk = self.channel.unary_stream if stream else self.channel.unary_unary
        return mk(path, request_serializer=MSG[req].SerializeToString, response_deserializer=MSG[resp].FromString)(request,
                                                                                                                  metadata=self.metadata)

    def close(self) -> None:
        self.channel.close()


def main(argv: Optional[List[str]] = None) -> None:
    ap = argparse.ArgumentParser("sglang_b200 gRPC server")
    ap.add_argument("--emit-proto", action="store_true")
    ap.add_argument("--grpc-port", type=int, default=30001)
    ns, rest = ap.parse_known_args(argv)
    if ns.emit_proto:
        print(emit_proto(), end="")
        return
    from ..server_args import prepare_server_args
    from .engine import Engine

    args = prepare_server_args(rest)
    eng = Engine(server_args=args)
    eng.start_background_loop()
    server, port = make_server(eng, args.host, ns.grpc_port)
    server.start()
    logger.info("gRPC server on %s:%d", args.host, port)
    server.wait_for_termination()


if __name__ == "__main__":
    main()
