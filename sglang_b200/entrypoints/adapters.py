"""Protocol adapters mounted on the HTTP app: Anthropic Messages, Ollama, SageMaker, Vertex AI, OpenAI Responses / classify.

Parity: reference ``python/sglang/srt/entrypoints/anthropic/serving.py`` (Messages API incl. SSE event stream),
``entrypoints/ollama/serving.py`` (``/api/generate``, ``/api/chat``, ``/api/tags``, ``/api/show``),
``http_server.py:1973-2060`` (SageMaker ``/ping`` + ``/invocations``, Vertex ``/vertex_generate``),
``openai/serving_responses.py`` (Responses API), ``serving_classify.py``.

All of them are thin shims: they translate the foreign request into the engine's chat / generate path (chat templating,
sampling params, tool calls and streaming stay in ``OpenAIServing``) and translate the result back.
"""
from __future__ import annotations

import json
import re
import time
import uuid
from typing import Any, AsyncIterator, Dict, List, Optional

from fastapi import Request
from fastapi.responses import JSONResponse, PlainTextResponse, StreamingResponse

from .openai import protocol as P


def _now_iso() -> str:
    return time.strftime("%Y-%m-%dT%H:%M:%S.000000Z", time.gmtime())


# ---------------------------------------------------------------- Anthropic Messages
def _anthropic_to_chat(body: Dict[str, Any]) -> P.ChatCompletionRequest:
    msgs: List[Dict[str, Any]] = []
    system = body.get("system")
    if system:
        if isinstance(system, list):
            system = "".join(b.get("text", "") for b in system if b.get("type") == "text")
        msgs.append({"role": "system", "content": system})
    for m in body.get("messages", []):
        c = m.get("content")
        if isinstance(c, list):
            parts, tool_results = [], []
            for b in c:
                t = b.get("type")
                if t == "text":
                    parts.append({"type": "text", "text": b.get("text", "")})
                elif t == "image":
                    src = b.get("source", {})
                    url = src.get("url") or f"data:{src.get('media_type', 'image/png')};base64,{src.get('data', '')}"
                    parts.append({"type": "image_url", "image_url": {"url": url}})
                elif t == "tool_result":
                    rc = b.get("content")
                    if isinstance(rc, list):
                        rc = "".join(x.get("text", "") for x in rc if x.get("type") == "text")
                    tool_results.append({"role": "tool", "tool_call_id": b.get("tool_use_id"), "content": rc or ""})
                elif t == "tool_use":
                    msgs.append({"role": "assistant", "content": None,
                                 "tool_calls": [{"id": b.get("id"), "type": "function",
                                                 "function": {"name": b.get("name"), "arguments": json.dumps(b.get("input", {}))}}]})
            msgs.extend(tool_results)
            if parts:
                only_text = all(p["type"] == "text" for p in parts)
                msgs.append({"role": m["role"], "content": "".join(p["text"] for p in parts) if only_text else parts})
        else:
            msgs.append({"role": m["role"], "content": c})
    tools = None
    if body.get("tools"):
        tools = [{"type": "function", "function": {"name": t["name"], "description": t.get("description", ""),
                                                   "parameters": t.get("input_schema", {})}} for t in body["tools"]]
    tc = body.get("tool_choice")
    tool_choice: Any = "auto"
    if isinstance(tc, dict):
        tool_choice = {"any": "required", "auto": "auto", "none": "none"}.get(tc.get("type"), "auto")
        if tc.get("type") == "tool":
            tool_choice = {"type": "function", "function": {"name": tc.get("name")}}
    kw: Dict[str, Any] = dict(model=body.get("model", "default"), messages=msgs, max_tokens=body.get("max_tokens", 1024),
                              stream=bool(body.get("stream")), tools=tools, tool_choice=tool_choice if tools else "none")
    for src, dst in (("temperature", "temperature"), ("top_p", "top_p"), ("top_k", "top_k"), ("stop_sequences", "stop")):
        if body.get(src) is not None:
            kw[dst] = body[src]
    return P.ChatCompletionRequest(**kw)


_STOP_MAP = {"stop": "end_turn", "length": "max_tokens", "tool_calls": "tool_use", "abort": "end_turn", None: "end_turn"}


def _chat_to_anthropic(resp: P.ChatCompletionResponse, model: str) -> Dict[str, Any]:
    ch = resp.choices[0]
    content: List[Dict[str, Any]] = []
    if ch.message.content:
        content.append({"type": "text", "text": ch.message.content})
    for tcall in ch.message.tool_calls or []:
        try:
            args = json.loads(tcall.function.arguments or "{}")
        except json.JSONDecodeError:
            args = {"_raw": tcall.function.arguments}
        content.append({"type": "tool_use", "id": tcall.id, "name": tcall.function.name, "input": args})
    stop = _STOP_MAP.get(ch.finish_reason, "end_turn")
    if ch.finish_reason == "stop" and ch.matched_stop is not None and isinstance(ch.matched_stop, str):
        stop = "stop_sequence"
    return {"id": "msg_" + uuid.uuid4().hex[:24], "type": "message", "role": "assistant", "model": model, "content": content,
            "stop_reason": stop, "stop_sequence": ch.matched_stop if stop == "stop_sequence" else None,
            "usage": {"input_tokens": resp.usage.prompt_tokens, "output_tokens": resp.usage.completion_tokens}}


async def _anthropic_stream(chat_stream: AsyncIterator[str], model: str) -> AsyncIterator[str]:
    """OpenAI chat SSE chunks -> Anthropic event stream (message_start, content_block_*, message_delta, message_stop)"""
    def ev(name: str, data: Dict[str, Any]) -> str:
        return f"event: {name}\ndata: {json.dumps(data)}\n\n"

    mid = "msg_" + uuid.uuid4().hex[:24]
    yield ev("message_start", {"type": "message_start", "message": {"id": mid, "type": "message", "role": "assistant", "model": model,
                                                                     "content": [], "stop_reason": None, "stop_sequence": None,
                                                                     "usage": {"input_tokens": 0, "output_tokens": 0}}})
    block_open, idx, n_out, stop = False, 0, 0, "end_turn"
    async for line in chat_stream:
        if not line.startswith("data: ") or line.strip() == "data: [DONE]":
            continue
        ch = json.loads(line[6:])
        if ch.get("usage"):
            n_out = ch["usage"].get("completion_tokens", n_out)
        for c in ch.get("choices", []):
            d = c.get("delta", {})
            if d.get("content"):
                if not block_open:
                    yield ev("content_block_start", {"type": "content_block_start", "index": idx, "content_block": {"type": "text", "text": ""}})
                    block_open = True
                n_out += 1
                yield ev("content_block_delta", {"type": "content_block_delta", "index": idx, "delta": {"type": "text_delta", "text": d["content"]}})
            for tc in d.get("tool_calls") or []:
                if block_open:
                    yield ev("content_block_stop", {"type": "content_block_stop", "index": idx})
                    idx, block_open = idx + 1, False
                fn = tc.get("function", {})
                yield ev("content_block_start", {"type": "content_block_start", "index": idx,
                                                 "content_block": {"type": "tool_use", "id": tc.get("id"), "name": fn.get("name"), "input": {}}})
                yield ev("content_block_delta", {"type": "content_block_delta", "index": idx,
                                                 "delta": {"type": "input_json_delta", "partial_json": fn.get("arguments") or ""}})
                yield ev("content_block_stop", {"type": "content_block_stop", "index": idx})
                idx += 1
            if c.get("finish_reason"):
                stop = _STOP_MAP.get(c["finish_reason"], "end_turn")
    if block_open:
        yield ev("content_block_stop", {"type": "content_block_stop", "index": idx})
    yield ev("message_delta", {"type": "message_delta", "delta": {"stop_reason": stop, "stop_sequence": None}, "usage": {"output_tokens": n_out}})
    yield ev("message_stop", {"type": "message_stop"})


# ---------------------------------------------------------------- Ollama
def _ollama_options(body: Dict[str, Any]) -> Dict[str, Any]:
    o = body.get("options") or {}
    kw: Dict[str, Any] = {}
    for src, dst in (("temperature", "temperature"), ("top_p", "top_p"), ("top_k", "top_k"), ("stop", "stop"), ("seed", "seed"),
                     ("repeat_penalty", "repetition_penalty"), ("presence_penalty", "presence_penalty"),
                     ("frequency_penalty", "frequency_penalty")):
        if o.get(src) is not None:
            kw[dst] = o[src]
    kw["max_tokens"] = o.get("num_predict") if (o.get("num_predict") or 0) > 0 else 512
    if body.get("format") == "json":
        kw["response_format"] = {"type": "json_object"}
    elif isinstance(body.get("format"), dict):
        kw["response_format"] = {"type": "json_schema", "json_schema": {"name": "out", "schema": body["format"]}}
    return kw


def parse_multipart(body: bytes, content_type: str) -> Dict[str, Any]:
    """minimal multipart/form-data reader (no python-multipart in the image): {field: str | bytes (file parts)}"""
    m = re.search(r'boundary="?([^";]+)"?', content_type or "")
    if not m:
        raise ValueError("multipart body without a boundary")
    out: Dict[str, Any] = {}
    for part in body.split(b"--" + m.group(1).encode()):
        part = part.strip(b"\r\n")
        if not part or part == b"--" or b"\r\n\r\n" not in part:
            continue
        head, data = part.split(b"\r\n\r\n", 1)
        hd = head.decode("utf-8", "replace")
        nm = re.search(r'name="([^"]*)"', hd)
        if not nm:
            continue
        out[nm.group(1)] = data if "filename=" in hd else data.decode("utf-8", "replace")
    return out


def _whisper_prompt_ids(tokenizer: Any, language: Optional[str], task: str = "transcribe") -> List[int]:
    toks = ["<|startoftranscript|>"] + ([f"<|{language}|>"] if language else []) + [f"<|{task}|>", "<|notimestamps|>"]
    ids = tokenizer.convert_tokens_to_ids(toks)
    if any(i is None for i in ids):
        raise ValueError("the tokenizer has no Whisper control tokens; pass prompt_ids")
    return [int(i) for i in ids]


def install_adapters(app: Any, engine: Any, serving: Any) -> None:
    model_name = engine.server_args.served_model_name

    # ----- speech to text (reference entrypoints/openai/serving_transcription.py)
    @app.post("/v1/audio/transcriptions")
    async def transcriptions(request: Request):
        ctype = request.headers.get("content-type", "")
        raw = await request.body()
        try:
            if ctype.startswith("multipart/"):
                form = parse_multipart(raw, ctype)
                audio = form.get("file")
            else:
                form = json.loads(raw or b"{}")
                audio = form.get("audio") or form.get("file")
            if audio is None:
                raise ValueError("no audio: send a multipart 'file' part or a JSON 'audio' (base64 WAV)")
            if engine.model_config.audio is None:
                raise ValueError(f"{model_name} is not a speech model")
            ids = form.get("prompt_ids")
            if isinstance(ids, str):
                ids = json.loads(ids)
            if ids is None:
                ids = _whisper_prompt_ids(engine.tokenizer, form.get("language") or "en")
            sp = {"max_new_tokens": int(form.get("max_tokens") or engine.model_config.max_position_embeddings - len(ids) - 1),
                  "temperature": float(form.get("temperature") or 0.0)}
            o = await engine.async_generate(input_ids=[int(i) for i in ids], sampling_params=sp, audio_data=audio)
        except (ValueError, KeyError, TypeError) as e:
            return JSONResponse({"error": {"message": str(e), "type": "invalid_request_error"}}, status_code=400)
        text = o.get("text")
        if text is None and engine.tokenizer is not None:
            text = engine.tokenizer.decode(o["output_ids"], skip_special_tokens=True)
        if (form.get("response_format") or "json") == "text":
            return PlainTextResponse(text or "")
        return {"text": text or "", "output_ids": o["output_ids"], "usage": {"type": "tokens", "input_tokens": o["meta_info"]["prompt_tokens"],
                                                                            "output_tokens": o["meta_info"]["completion_tokens"]}}

    # ----- Anthropic
    @app.post("/v1/messages")
    async def anthropic_messages(request: Request):
        body = await request.json()
        try:
            r = _anthropic_to_chat(body)
            out = await serving.chat_completions(r)
        except (ValueError, KeyError) as e:
            return JSONResponse({"type": "error", "error": {"type": "invalid_request_error", "message": str(e)}}, status_code=400)
        if r.stream:
            return StreamingResponse(_anthropic_stream(out, r.model), media_type="text/event-stream")
        return _chat_to_anthropic(out, r.model)

    @app.post("/v1/messages/count_tokens")
    async def anthropic_count(request: Request):
        body = await request.json()
        r = _anthropic_to_chat({**body, "max_tokens": 1})
        prompt, _, _ = serving._render_chat(r)
        return {"input_tokens": len(engine.tokenizer.encode(prompt))}

    # ----- Ollama
    @app.get("/api/tags")
    async def ollama_tags():
        return {"models": [{"name": model_name, "model": model_name, "modified_at": _now_iso(), "size": int(engine.runner.weight_bytes),
                            "digest": "", "details": {"family": engine.model_config.architecture, "format": "safetensors"}}]}

    @app.get("/api/version")
    async def ollama_version():
        from .. import __version__

        return {"version": __version__}

    @app.post("/api/show")
    async def ollama_show(request: Request):
        return {"modelfile": "", "parameters": "", "template": "", "details": {"family": engine.model_config.architecture},
                "model_info": {"context_length": engine.runner.max_context_len}}

    def _ollama_done(model: str, meta: Dict[str, Any], t0: float) -> Dict[str, Any]:
        fr = meta.get("finish_reason") or {}
        return {"model": model, "created_at": _now_iso(), "done": True, "done_reason": "length" if fr.get("type") == "length" else "stop",
                "total_duration": int((time.time() - t0) * 1e9), "prompt_eval_count": meta.get("prompt_tokens", 0),
                "eval_count": meta.get("completion_tokens", 0)}

    @app.post("/api/generate")
    async def ollama_generate(request: Request):
        body = await request.json()
        model = body.get("model", model_name)
        kw = _ollama_options(body)
        prompt = (body.get("system", "") + "\n" if body.get("system") else "") + body.get("prompt", "")
        r = P.CompletionRequest(model=model, prompt=prompt, stream=False, **kw)
        sp = r.base_sampling(r.max_tokens)
        t0 = time.time()
        if body.get("stream", True):
            async def gen():
                sent = 0
                g = await engine.async_generate(prompt, sampling_params=sp, stream=True)
                async for o in g:
                    delta, sent = o["text"][sent:], len(o["text"])
                    if o["meta_info"]["finish_reason"] is None:
                        yield json.dumps({"model": model, "created_at": _now_iso(), "response": delta, "done": False}) + "\n"
                    else:
                        if delta:
                            yield json.dumps({"model": model, "created_at": _now_iso(), "response": delta, "done": False}) + "\n"
                        yield json.dumps({**_ollama_done(model, o["meta_info"], t0), "response": ""}) + "\n"
            return StreamingResponse(gen(), media_type="application/x-ndjson")
        o = await engine.async_generate(prompt, sampling_params=sp)
        return {**_ollama_done(model, o["meta_info"], t0), "response": o["text"]}

    @app.post("/api/chat")
    async def ollama_chat(request: Request):
        body = await request.json()
        model = body.get("model", model_name)
        kw = _ollama_options(body)
        tools = body.get("tools")
        r = P.ChatCompletionRequest(model=model, messages=body.get("messages", []), stream=bool(body.get("stream", True)), tools=tools,
                                    tool_choice="auto" if tools else "none", **kw)
        t0 = time.time()
        out = await serving.chat_completions(r)
        if r.stream:
            async def gen():
                meta = {"prompt_tokens": 0, "completion_tokens": 0, "finish_reason": {"type": "stop"}}
                async for line in out:
                    if not line.startswith("data: ") or line.strip() == "data: [DONE]":
                        continue
                    for c in json.loads(line[6:]).get("choices", []):
                        d = c.get("delta", {})
                        if d.get("content"):
                            meta["completion_tokens"] += 1
                            yield json.dumps({"model": model, "created_at": _now_iso(), "message": {"role": "assistant", "content": d["content"]},
                                              "done": False}) + "\n"
                        if c.get("finish_reason"):
                            meta["finish_reason"] = {"type": c["finish_reason"]}
                yield json.dumps({**_ollama_done(model, meta, t0), "message": {"role": "assistant", "content": ""}}) + "\n"
            return StreamingResponse(gen(), media_type="application/x-ndjson")
        ch = out.choices[0]
        msg: Dict[str, Any] = {"role": "assistant", "content": ch.message.content or ""}
        if ch.message.tool_calls:
            msg["tool_calls"] = [{"function": {"name": t.function.name, "arguments": json.loads(t.function.arguments or "{}")}}
                                 for t in ch.message.tool_calls]
        meta = {"prompt_tokens": out.usage.prompt_tokens, "completion_tokens": out.usage.completion_tokens,
                "finish_reason": {"type": ch.finish_reason}}
        return {**_ollama_done(model, meta, t0), "message": msg}

    # ----- SageMaker: health on /ping, OpenAI chat body on /invocations
    @app.get("/ping")
    async def sagemaker_ping():
        return JSONResponse({"status": "ok"})

    @app.post("/invocations")
    async def sagemaker_invocations(request: Request):
        body = await request.json()
        if "messages" in body:
            r = P.ChatCompletionRequest(**body)
            out = await serving.chat_completions(r)
        elif "prompt" in body:
            r = P.CompletionRequest(**body)
            out = await serving.completions(r)
        else:      # native /generate body
            kw = {k: body[k] for k in ("sampling_params", "return_logprob", "logprob_start_len", "top_logprobs_num", "lora_path") if k in body}
            return await engine.async_generate(body.get("text"), input_ids=body.get("input_ids"), **kw)
        return StreamingResponse(out, media_type="text/event-stream") if getattr(r, "stream", False) else out

    # ----- Vertex AI prediction container contract: {"instances": [...], "parameters": {...}} -> {"predictions": [...]}
    @app.post("/vertex_generate")
    async def vertex_generate(request: Request):
        body = await request.json()
        inst = body.get("instances") or []
        params = body.get("parameters") or {}
        texts = [i.get("text") or i.get("prompt") for i in inst if isinstance(i, dict)] or [str(i) for i in inst]
        sp = params.get("sampling_params") or {k: v for k, v in params.items() if k in ("max_new_tokens", "temperature", "top_p", "top_k", "stop")}
        outs = await engine.async_generate(texts, sampling_params=sp)
        outs = outs if isinstance(outs, list) else [outs]
        return {"predictions": outs}

    # ----- OpenAI Responses API (non-streaming + streaming text), mapped onto chat completions
    @app.post("/v1/responses")
    async def v1_responses(request: Request):
        body = await request.json()
        inp = body.get("input")
        msgs: List[Dict[str, Any]] = []
        if body.get("instructions"):
            msgs.append({"role": "system", "content": body["instructions"]})
        if isinstance(inp, str):
            msgs.append({"role": "user", "content": inp})
        else:
            for it in inp or []:
                c = it.get("content")
                if isinstance(c, list):
                    c = "".join(p.get("text", "") for p in c if p.get("type") in ("input_text", "output_text", "text"))
                msgs.append({"role": it.get("role", "user"), "content": c})
        kw: Dict[str, Any] = {}
        for src, dst in (("temperature", "temperature"), ("top_p", "top_p"), ("max_output_tokens", "max_tokens")):
            if body.get(src) is not None:
                kw[dst] = body[src]
        tools = [{"type": "function", "function": {k: t[k] for k in ("name", "description", "parameters") if k in t}}
                 for t in body.get("tools") or [] if t.get("type") == "function"] or None
        r = P.ChatCompletionRequest(model=body.get("model", model_name), messages=msgs, stream=bool(body.get("stream")), tools=tools,
                                    tool_choice="auto" if tools else "none", **kw)
        rid = "resp_" + uuid.uuid4().hex[:24]
        out = await serving.chat_completions(r)
        if r.stream:
            async def gen():
                text = ""
                yield f"event: response.created\ndata: {json.dumps({'type': 'response.created', 'response': {'id': rid, 'status': 'in_progress'}})}\n\n"
                async for line in out:
                    if not line.startswith("data: ") or line.strip() == "data: [DONE]":
                        continue
                    for c in json.loads(line[6:]).get("choices", []):
                        d = c.get("delta", {}).get("content")
                        if d:
                            text += d
                            yield f"event: response.output_text.delta\ndata: {json.dumps({'type': 'response.output_text.delta', 'delta': d})}\n\n"
                done = {"type": "response.completed", "response": {"id": rid, "status": "completed", "output_text": text}}
                yield f"event: response.completed\ndata: {json.dumps(done)}\n\n"
            return StreamingResponse(gen(), media_type="text/event-stream")
        ch = out.choices[0]
        output: List[Dict[str, Any]] = []
        if ch.message.content:
            output.append({"type": "message", "id": "msg_" + uuid.uuid4().hex[:24], "role": "assistant", "status": "completed",
                           "content": [{"type": "output_text", "text": ch.message.content, "annotations": []}]})
        for t in ch.message.tool_calls or []:
            output.append({"type": "function_call", "id": t.id, "call_id": t.id, "name": t.function.name, "arguments": t.function.arguments})
        return {"id": rid, "object": "response", "created_at": int(time.time()), "status": "completed", "model": r.model, "output": output,
                "output_text": ch.message.content or "",
                "usage": {"input_tokens": out.usage.prompt_tokens, "output_tokens": out.usage.completion_tokens,
                          "total_tokens": out.usage.total_tokens}}

    # ----- classification (sequence-classification / reward heads): softmax over the pooled logits
    @app.post("/v1/classify")
    async def v1_classify(request: Request):
        import asyncio

        import torch

        body = await request.json()
        inp = body.get("input")
        texts = [inp] if isinstance(inp, str) else list(inp or [])
        outs = await asyncio.get_running_loop().run_in_executor(None, engine.encode, texts)
        data = []
        for i, o in enumerate(outs):
            logits = torch.tensor(o["embedding"], dtype=torch.float32)
            probs = torch.softmax(logits, -1).tolist()
            data.append({"index": i, "label": str(int(max(range(len(probs)), key=probs.__getitem__))), "probs": probs, "num_classes": len(probs)})
        return {"id": "classify-" + uuid.uuid4().hex[:16], "object": "list", "model": body.get("model", model_name), "data": data}
