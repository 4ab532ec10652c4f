"""OpenAI-compatible serving logic on top of ``Engine`` (completions, chat, embeddings, tokenize, score, rerank).

Parity: reference ``python/sglang/srt/entrypoints/openai/`` (``serving_base.py:26,73``, ``serving_chat.py:243``
``_convert_to_internal_request:965`` / ``_process_messages:1099`` / streaming ``:1514``, ``serving_completions.py:46``,
``serving_embedding.py``, tool-call + reasoning parsing in the streaming path).
"""
from __future__ import annotations

import json
from typing import Any, AsyncIterator, Dict, List, Optional, Tuple

from ...function_call.function_call_parser import FunctionCallParser
from ...parser.reasoning_parser import ReasoningParser
from . import protocol as P


def _finish(fr: Optional[Dict[str, Any]]) -> Tuple[Optional[str], Any]:
    if fr is None:
        return None, None
    t = fr.get("type")
    return ("stop" if t == "stop" else "length" if t == "length" else "abort"), fr.get("matched")


def _apply_response_format(sp: Dict[str, Any], rf: Optional[P.ResponseFormat]) -> None:
    if rf is None or rf.type == "text":
        return
    if rf.type == "json_object":
        sp["json_schema"] = '{"type": "object"}'
    elif rf.type == "json_schema" and rf.json_schema is not None:
        sp["json_schema"] = json.dumps(rf.json_schema.schema_ or {"type": "object"})
    elif rf.type == "regex":
        sp["regex"] = rf.pattern
    elif rf.type == "ebnf":
        sp["ebnf"] = rf.grammar


class OpenAIServing:
    def __init__(self, engine: Any):
        self.engine = engine
        self.args = engine.server_args
        self.tok = engine.tokenizer
        self.model_name = self.args.served_model_name

    # ------------------------------------------------------------------ models / tokenize
    def models(self) -> P.ModelList:
        return P.ModelList(data=[P.ModelCard(id=self.model_name, root=self.args.model_path,
                                             max_model_len=self.engine.scheduler.max_context_len)])

    def tokenize(self, r: P.TokenizeRequest) -> Dict[str, Any]:
        ps = r.prompt if isinstance(r.prompt, list) else [r.prompt]
        toks = [self.tok.encode(p) for p in ps]
        return {"tokens": toks if isinstance(r.prompt, list) else toks[0], "count": [len(t) for t in toks] if isinstance(r.prompt, list) else len(toks[0]),
                "max_model_len": self.engine.scheduler.max_context_len}

    def detokenize(self, r: P.DetokenizeRequest) -> Dict[str, Any]:
        batch = r.tokens and isinstance(r.tokens[0], list)
        ts = r.tokens if batch else [r.tokens]
        texts = [self.tok.decode(t, skip_special_tokens=r.skip_special_tokens) for t in ts]
        return {"text": texts if batch else texts[0]}

    # ------------------------------------------------------------------ completions
    def _completion_inputs(self, r: P.CompletionRequest):
        p = r.prompt
        if isinstance(p, str):
            return dict(prompt=p), 1
        if p and isinstance(p[0], int):
            return dict(input_ids=p), 1
        if p and isinstance(p[0], list):
            return dict(input_ids=p), len(p)
        return dict(prompt=p), len(p)

    def _logprobs_obj(self, meta: Dict[str, Any], k: int) -> P.LogProbs:
        lp = P.LogProbs()
        tops = meta.get("output_top_logprobs") or []
        for i, (v, tid, _) in enumerate(meta.get("output_token_logprobs") or []):
            lp.tokens.append(self.tok.decode([tid]))
            lp.token_logprobs.append(v)
            if k and i < len(tops) and tops[i]:
                lp.top_logprobs.append({self.tok.decode([t]): float(x) for x, t in tops[i][:k]})
        return lp

    async def completions(self, r: P.CompletionRequest):
        kw, n_prompts = self._completion_inputs(r)
        sp = r.base_sampling(r.max_tokens)
        _apply_response_format(sp, r.response_format)
        want_lp = r.logprobs is not None
        rid = r.rid or P.new_id("cmpl")
        gen_kw = dict(sampling_params=sp, return_logprob=want_lp, top_logprobs_num=r.logprobs or 0, lora_path=r.lora_path,
                      priority=r.priority)
        if r.stream:
            return self._stream_completion(r, kw, gen_kw, rid)
        outs = await self.engine.async_generate(**kw, **gen_kw)
        outs = outs if isinstance(outs, list) else [outs]
        choices, pt, ct = [], 0, 0
        seen_prompts = set()
        for i, o in enumerate(outs):
            m = o["meta_info"]
            fr, matched = _finish(m["finish_reason"])
            text = o["text"]
            if r.echo and isinstance(r.prompt, str):
                text = r.prompt + text
            choices.append(P.CompletionChoice(index=i, text=text, finish_reason=fr, matched_stop=matched,
                                              logprobs=self._logprobs_obj(m, r.logprobs) if want_lp else None))
            ct += m["completion_tokens"]
            key = i // max(1, r.n)
            if key not in seen_prompts:
                seen_prompts.add(key)
                pt += m["prompt_tokens"]
        return P.CompletionResponse(id=rid, model=r.model, choices=choices,
                                    usage=P.UsageInfo(prompt_tokens=pt, completion_tokens=ct, total_tokens=pt + ct))

    async def _stream_completion(self, r, kw, gen_kw, rid) -> AsyncIterator[str]:
        gen = await self.engine.async_generate(**kw, **gen_kw, stream=True)
        sent: Dict[str, int] = {}
        idx: Dict[str, int] = {}
        pt = ct = 0
        async for o in gen:
            m = o["meta_info"]
            i = idx.setdefault(m["id"], len(idx))
            delta = o["text"][sent.get(m["id"], 0):]
            sent[m["id"]] = len(o["text"])
            fr, matched = _finish(m["finish_reason"])
            if fr is not None:
                pt += m["prompt_tokens"]
                ct += m["completion_tokens"]
            ch = P.CompletionChoice(index=i, text=delta, finish_reason=fr, matched_stop=matched)
            yield "data: " + P.CompletionResponse(id=rid, model=r.model, choices=[ch], usage=P.UsageInfo()).model_dump_json() + "\n\n"
        if r.stream_options and r.stream_options.include_usage:
            u = P.UsageInfo(prompt_tokens=pt, completion_tokens=ct, total_tokens=pt + ct)
            yield "data: " + P.CompletionResponse(id=rid, model=r.model, choices=[], usage=u).model_dump_json() + "\n\n"
        yield "data: [DONE]\n\n"

    # ------------------------------------------------------------------ chat
    def _render_chat(self, r: P.ChatCompletionRequest) -> Tuple[str, Optional[List[Any]], Optional[List[Any]]]:
        msgs = [m.model_dump(exclude_none=True) for m in r.messages]
        images: List[Any] = []
        audios: List[Any] = []
        for m in msgs:
            c = m.get("content")
            if isinstance(c, list):
                for part in c:
                    if part.get("type") == "image_url":
                        images.append(part["image_url"]["url"] if isinstance(part["image_url"], dict) else part["image_url"])
                    elif part.get("type") == "audio_url":          # URL / path / data-URL of a WAV clip
                        audios.append(part["audio_url"]["url"] if isinstance(part["audio_url"], dict) else part["audio_url"])
                    elif part.get("type") == "input_audio":        # OpenAI form: {"data": <base64>, "format": "wav"}
                        audios.append(part["input_audio"]["data"] if isinstance(part["input_audio"], dict) else part["input_audio"])
        tools = [t.model_dump() for t in r.tools] if r.tools else None
        kw = dict(r.chat_template_kwargs or {})
        try:
            text = self.tok.apply_chat_template(msgs, tokenize=False, add_generation_prompt=not r.continue_final_message,
                                                tools=tools, **kw)
        except TypeError:
            text = self.tok.apply_chat_template(msgs, tokenize=False, add_generation_prompt=not r.continue_final_message)
        return text, (images or None), (audios or None)

    def _chat_sampling(self, r: P.ChatCompletionRequest) -> Dict[str, Any]:
        sp = r.base_sampling(r.max_completion_tokens or r.max_tokens)
        _apply_response_format(sp, r.response_format)
        if r.tools and r.tool_choice not in ("auto", "none") and not any(k in sp for k in ("regex", "ebnf", "json_schema")):
            parser = FunctionCallParser(r.tools, self.args.tool_call_parser or "json")
            st = parser.get_structural_constraint(r.tool_choice)
            if st is not None:
                sp["structural_tag"] = json.dumps(st)
        ck = r.chat_template_kwargs or {}
        if self.args.reasoning_parser and ck.get("enable_thinking", True) and ck.get("thinking", True) and \
                any(k in sp for k in ("regex", "ebnf", "json_schema", "structural_tag")):
            # the model will think first: the grammar must only bind the answer (reference serving_chat.py require_reasoning)
            sp["custom_params"] = dict(sp.get("custom_params") or {}, require_reasoning=True)
        return sp

    async def chat_completions(self, r: P.ChatCompletionRequest):
        prompt, images, audios = self._render_chat(r)
        sp = self._chat_sampling(r)
        rid = r.rid or P.new_id("chatcmpl")
        gen_kw = dict(sampling_params=sp, return_logprob=r.logprobs, top_logprobs_num=r.top_logprobs or 0,
                      lora_path=r.lora_path, priority=r.priority, image_data=images)
        if audios:
            gen_kw["audio_data"] = audios
        if r.stream:
            return self._stream_chat(r, prompt, gen_kw, rid)
        outs = await self.engine.async_generate(prompt, **gen_kw)
        outs = outs if isinstance(outs, list) else [outs]
        choices, ct = [], 0
        for i, o in enumerate(outs):
            m = o["meta_info"]
            fr, matched = _finish(m["finish_reason"])
            text, reasoning, calls = o["text"], None, None
            if self.args.reasoning_parser and r.separate_reasoning:
                reasoning, text = ReasoningParser(self.args.reasoning_parser).parse_non_stream(text)
            if r.tools and r.tool_choice != "none":
                parser = FunctionCallParser(r.tools, self.args.tool_call_parser or "json")
                if parser.has_tool_call(text):
                    text, items = parser.parse_non_stream(text)
                    if items:
                        calls = [P.ToolCall(id=P.new_id("call"), index=c.tool_index,
                                            function=P.FunctionCall(name=c.name, arguments=c.parameters)) for c in items]
                        fr = "tool_calls"
            msg = P.ChatMessage(role="assistant", content=text if (text or not calls) else None, tool_calls=calls,
                                reasoning_content=reasoning)
            lp = None
            if r.logprobs:
                lo = self._logprobs_obj(m, r.top_logprobs or 0)
                lp = {"content": [{"token": t, "logprob": v, "top_logprobs": [{"token": k, "logprob": x} for k, x in (tl or {}).items()]}
                                  for t, v, tl in zip(lo.tokens, lo.token_logprobs, lo.top_logprobs + [None] * len(lo.tokens))]}
            choices.append(P.ChatCompletionChoice(index=i, message=msg, finish_reason=fr, matched_stop=matched, logprobs=lp))
            ct += m["completion_tokens"]
        pt = outs[0]["meta_info"]["prompt_tokens"]
        cached = outs[0]["meta_info"].get("cached_tokens", 0)
        return P.ChatCompletionResponse(id=rid, model=r.model, choices=choices,
                                        usage=P.UsageInfo(prompt_tokens=pt, completion_tokens=ct, total_tokens=pt + ct,
                                                          prompt_tokens_details={"cached_tokens": cached}))

    async def _stream_chat(self, r, prompt, gen_kw, rid) -> AsyncIterator[str]:
        gen = await self.engine.async_generate(prompt, **gen_kw, stream=True)
        state: Dict[str, Dict[str, Any]] = {}
        pt = ct = 0

        def chunk(i, delta, fr=None, matched=None):
            c = P.ChatCompletionStreamChoice(index=i, delta=delta, finish_reason=fr, matched_stop=matched)
            return "data: " + P.ChatCompletionStreamResponse(id=rid, model=r.model, choices=[c]).model_dump_json() + "\n\n"

        async for o in gen:
            m = o["meta_info"]
            st = state.get(m["id"])
            if st is None:
                st = state[m["id"]] = {"i": len(state), "sent": 0, "calls": 0,
                                       "rp": ReasoningParser(self.args.reasoning_parser, r.stream_reasoning) if (self.args.reasoning_parser and r.separate_reasoning) else None,
                                       "fp": FunctionCallParser(r.tools, self.args.tool_call_parser or "json") if (r.tools and r.tool_choice != "none") else None}
                yield chunk(st["i"], P.DeltaMessage(role="assistant", content=""))
            delta = o["text"][st["sent"]:]
            st["sent"] = len(o["text"])
            reasoning_delta = ""
            if st["rp"] is not None and delta:
                reasoning_delta, delta = st["rp"].parse_stream_chunk(delta)
            if reasoning_delta:
                yield chunk(st["i"], P.DeltaMessage(reasoning_content=reasoning_delta))
            new_calls = []
            if st["fp"] is not None and delta:
                delta, new_calls = st["fp"].parse_stream_chunk(delta)
            for c in new_calls:
                st["calls"] += 1
                tc = P.ToolCall(id=P.new_id("call"), index=c.tool_index, function=P.FunctionCall(name=c.name, arguments=c.parameters))
                yield chunk(st["i"], P.DeltaMessage(tool_calls=[tc]))
            fr, matched = _finish(m["finish_reason"])
            if delta:
                yield chunk(st["i"], P.DeltaMessage(content=delta))
            if fr is not None:
                pt, ct = m["prompt_tokens"], ct + m["completion_tokens"]
                yield chunk(st["i"], P.DeltaMessage(), "tool_calls" if st["calls"] else fr, matched)
        if r.stream_options and r.stream_options.include_usage:
            u = P.UsageInfo(prompt_tokens=pt, completion_tokens=ct, total_tokens=pt + ct)
            yield "data: " + P.ChatCompletionStreamResponse(id=rid, model=r.model, choices=[], usage=u).model_dump_json() + "\n\n"
        yield "data: [DONE]\n\n"

    # ------------------------------------------------------------------ embeddings / score / rerank
    async def embeddings(self, r: P.EmbeddingRequest) -> P.EmbeddingResponse:
        inp = r.input
        if isinstance(inp, str):
            texts = [inp]
        elif inp and isinstance(inp[0], int):
            texts = [self.tok.decode(inp)]
        elif inp and isinstance(inp[0], list):
            texts = [self.tok.decode(x) for x in inp]
        else:
            texts = list(inp)
        import asyncio

        outs = await asyncio.get_running_loop().run_in_executor(None, self.engine.encode, texts)
        data, pt = [], 0
        for i, o in enumerate(outs):
            e = o["embedding"]
            if r.dimensions:
                e = e[: r.dimensions]
            data.append(P.EmbeddingObject(embedding=e, index=i))
            pt += o["meta_info"]["prompt_tokens"]
        return P.EmbeddingResponse(data=data, model=r.model, usage=P.UsageInfo(prompt_tokens=pt, total_tokens=pt))
