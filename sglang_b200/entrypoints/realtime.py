"""``/v1/realtime`` — the OpenAI Realtime protocol over a WebSocket.

Reference: ``python/sglang/srt/entrypoints/openai/realtime/{handler,session,protocol}.py`` (transcription sessions: PCM16 audio is
appended to an input buffer, a commit turns the buffer into a conversation item and its transcription streams back as deltas).
Speech models get exactly that; for text models the same socket also carries the text half of the protocol
(``conversation.item.create`` + ``response.create`` -> ``response.text.delta`` ... ``response.done``), so one endpoint serves both.

A session holds: its config (sample rate, language, prompt), the audio buffer, and the conversation items.  Work for a commit /
response runs as a task, so the socket keeps accepting appends (and a ``response.cancel`` aborts the running request).
"""
from __future__ import annotations

import asyncio
import base64
import json
import time
import uuid
from typing import Any, Dict, List, Optional

import numpy as np
from fastapi import WebSocket, WebSocketDisconnect

MAX_BUFFER_S = 600


def _id(prefix: str) -> str:
    return f"{prefix}_{uuid.uuid4().hex[:20]}"


class RealtimeSession:
    def __init__(self, engine: Any, serving: Any, send):
        self.engine, self.serving, self._send = engine, serving, send
        self.id = _id("sess")
        self.is_speech = engine.model_config.audio is not None
        self.cfg: Dict[str, Any] = {"rate": 16000 if self.is_speech else 24000, "language": None, "prompt": None, "prompt_ids": None, "temperature": 0.0,
                                    "instructions": None, "max_output_tokens": 256}
        self.pcm: List[np.ndarray] = []
        self.pcm_samples = 0
        self.items: List[Dict[str, Any]] = []
        self.task: Optional[asyncio.Task] = None
        self.rid: Optional[str] = None

    async def send(self, type_: str, **kw: Any) -> None:
        await self._send(json.dumps({"type": type_, "event_id": _id("event"), **kw}))

    def describe(self) -> Dict[str, Any]:
        kind = "realtime.transcription_session" if self.is_speech else "realtime.session"
        return {"id": self.id, "object": kind, "model": self.engine.server_args.served_model_name,
                "audio": {"input": {"format": {"type": "audio/pcm", "rate": self.cfg["rate"]},
                                    "transcription": {"language": self.cfg["language"], "prompt": self.cfg["prompt"]}, "turn_detection": None}},
                "instructions": self.cfg["instructions"], "max_output_tokens": self.cfg["max_output_tokens"], "temperature": self.cfg["temperature"]}

    async def error(self, message: str, code: str = "invalid_request_error") -> None:
        await self.send("error", error={"type": code, "message": message})

    # ---- events ---------------------------------------------------------------------------------------------------------------
    async def handle(self, ev: Dict[str, Any]) -> None:
        t = ev.get("type")
        if t == "session.update":
            s = ev.get("session") or {}
            a = ((s.get("audio") or {}).get("input") or {})
            fmt, tr = a.get("format") or {}, a.get("transcription") or s.get("input_audio_transcription") or {}
            if fmt.get("type") not in (None, "audio/pcm", "pcm16"):
                return await self.error(f"unsupported audio format {fmt.get('type')!r}: send PCM16")
            if a.get("turn_detection") or s.get("turn_detection"):
                return await self.error("server-side turn detection is not supported: commit the buffer explicitly")
            self.cfg["rate"] = int(fmt.get("rate") or self.cfg["rate"])
            for k in ("language", "prompt", "prompt_ids"):
                if tr.get(k) is not None:
                    self.cfg[k] = tr[k]
            for k in ("instructions", "temperature", "max_output_tokens"):
                if s.get(k) is not None:
                    self.cfg[k] = s[k]
            return await self.send("session.updated", session=self.describe())
        if t == "input_audio_buffer.append":
            try:
                chunk = np.frombuffer(base64.b64decode(ev["audio"]), dtype="<i2")
            except Exception:
                return await self.error("input_audio_buffer.append needs base64 PCM16 in 'audio'")
            if (self.pcm_samples + chunk.size) > MAX_BUFFER_S * self.cfg["rate"]:
                return await self.error(f"audio buffer is limited to {MAX_BUFFER_S} s")
            self.pcm.append(chunk)
            self.pcm_samples += int(chunk.size)
            return
        if t == "input_audio_buffer.clear":
            self.pcm, self.pcm_samples = [], 0
            return await self.send("input_audio_buffer.cleared")
        if t == "input_audio_buffer.commit":
            if not self.pcm_samples:
                return await self.error("input_audio_buffer.commit on an empty buffer")
            if not self.is_speech:
                return await self.error(f"{self.engine.server_args.served_model_name} is not a speech model")
            wave = np.concatenate(self.pcm).astype(np.float32) / 32768.0
            self.pcm, self.pcm_samples = [], 0
            item = {"id": _id("item"), "type": "message", "role": "user", "content": [{"type": "input_audio", "transcript": None}]}
            prev = self.items[-1]["id"] if self.items else None
            self.items.append(item)
            await self.send("input_audio_buffer.committed", item_id=item["id"], previous_item_id=prev)
            await self.send("conversation.item.created", item=item, previous_item_id=prev)
            return self._spawn(self._transcribe(item, wave))
        if t == "conversation.item.create":
            item = dict(ev.get("item") or {})
            if item.get("type", "message") != "message" or item.get("role") not in ("user", "assistant", "system"):
                return await self.error("conversation.item.create takes a message item with role user / assistant / system")
            item.setdefault("id", _id("item"))
            prev = self.items[-1]["id"] if self.items else None
            self.items.append(item)
            return await self.send("conversation.item.created", item=item, previous_item_id=prev)
        if t == "conversation.item.delete":
            self.items = [i for i in self.items if i["id"] != ev.get("item_id")]
            return await self.send("conversation.item.deleted", item_id=ev.get("item_id"))
        if t == "response.create":
            if self.is_speech:
                return await self.error("this is a transcription session: commit audio instead of response.create")
            if self.task is not None and not self.task.done():
                return await self.error("a response is already in progress", "conversation_already_has_active_response")
            return self._spawn(self._respond(ev.get("response") or {}))
        if t == "response.cancel":
            if self.task is not None and not self.task.done():
                if self.rid:
                    self.engine.abort_request(self.rid)
                self.task.cancel()
            return
        await self.error(f"unknown event type {t!r}")

    def _spawn(self, coro) -> None:
        self.task = asyncio.ensure_future(coro)

    def close(self) -> None:
        if self.task is not None and not self.task.done():
            if self.rid:
                self.engine.abort_request(self.rid)
            self.task.cancel()

    # ---- work -----------------------------------------------------------------------------------------------------------------
    async def _transcribe(self, item: Dict[str, Any], wave: np.ndarray) -> None:
        from .adapters import _whisper_prompt_ids

        try:
            ids = self.cfg["prompt_ids"] or _whisper_prompt_ids(self.engine.tokenizer, self.cfg["language"] or "en")
            self.rid = _id("rt")
            sp = {"max_new_tokens": min(int(self.cfg["max_output_tokens"]), self.engine.model_config.max_position_embeddings - len(ids) - 1),
                  "temperature": float(self.cfg["temperature"])}
            gen = await self.engine.async_generate(input_ids=ids, sampling_params=sp, audio_data=(wave, self.cfg["rate"]), stream=True, rid=self.rid)
            sent, last = 0, None
            async for o in gen:
                last = o
                text = o.get("text")
                if text is None:
                    text = " ".join(str(t) for t in o["output_ids"])            # token-id transcript when there is no tokenizer
                if len(text) > sent:
                    await self.send("conversation.item.input_audio_transcription.delta", item_id=item["id"], content_index=0, delta=text[sent:])
                    sent = len(text)
            final = (last.get("text") if last and last.get("text") is not None else " ".join(str(t) for t in (last or {}).get("output_ids", [])))
            item["content"][0]["transcript"] = final
            await self.send("conversation.item.input_audio_transcription.completed", item_id=item["id"], content_index=0, transcript=final,
                            usage={"type": "tokens", "input_tokens": last["meta_info"]["prompt_tokens"],
                                   "output_tokens": last["meta_info"]["completion_tokens"]} if last else None)
        except asyncio.CancelledError:
            raise
        except Exception as e:      # noqa: BLE001 - reported on the socket
            await self.send("conversation.item.input_audio_transcription.failed", item_id=item["id"], content_index=0,
                            error={"type": "transcription_error", "message": f"{type(e).__name__}: {e}"})

    async def _respond(self, opts: Dict[str, Any]) -> None:
        rid = _id("resp")
        out_item = {"id": _id("item"), "type": "message", "role": "assistant", "content": [{"type": "text", "text": ""}]}
        await self.send("response.created", response={"id": rid, "object": "realtime.response", "status": "in_progress", "output": []})
        status, usage = "completed", None
        try:
            msgs = ([{"role": "system", "content": opts.get("instructions") or self.cfg["instructions"]}]
                    if (opts.get("instructions") or self.cfg["instructions"]) else [])
            for it in self.items:
                text = "".join(c.get("text") or c.get("transcript") or "" for c in it.get("content", []))
                msgs.append({"role": it["role"], "content": text})
            try:
                prompt = self.engine.tokenizer.apply_chat_template(msgs, tokenize=False, add_generation_prompt=True)
            except Exception:
                prompt = "\n".join(f"{m['role']}: {m['content']}" for m in msgs) + "\nassistant:"
            sp = {"max_new_tokens": int(opts.get("max_output_tokens") or self.cfg["max_output_tokens"]),
                  "temperature": float(opts.get("temperature", self.cfg["temperature"]))}
            self.rid = _id("rt")
            gen = await self.engine.async_generate(prompt, sp, stream=True, rid=self.rid)
            await self.send("response.output_item.added", response_id=rid, output_index=0, item=out_item)
            sent, last = 0, None
            async for o in gen:
                last, text = o, o.get("text") or ""
                if len(text) > sent:
                    await self.send("response.text.delta", response_id=rid, item_id=out_item["id"], output_index=0, content_index=0, delta=text[sent:])
                    sent = len(text)
            out_item["content"][0]["text"] = (last or {}).get("text") or ""
            self.items.append(out_item)
            await self.send("response.text.done", response_id=rid, item_id=out_item["id"], output_index=0, content_index=0, text=out_item["content"][0]["text"])
            await self.send("response.output_item.done", response_id=rid, output_index=0, item=out_item)
            if last:
                usage = {"input_tokens": last["meta_info"]["prompt_tokens"], "output_tokens": last["meta_info"]["completion_tokens"],
                         "total_tokens": last["meta_info"]["prompt_tokens"] + last["meta_info"]["completion_tokens"]}
        except asyncio.CancelledError:
            status = "cancelled"
        except Exception as e:      # noqa: BLE001
            status = "failed"
            await self.error(f"{type(e).__name__}: {e}", "server_error")
        await self.send("response.done", response={"id": rid, "object": "realtime.response", "status": status, "output": [out_item], "usage": usage})


def install_realtime(app: Any, engine: Any, serving: Any) -> None:
    @app.websocket("/v1/realtime")
    async def realtime(ws: WebSocket):
        from ..utils import auth as _auth

        ok, _, _ = _auth.decide("GET", "/v1/realtime", ws.headers.get("authorization"), engine.server_args.api_key,
                                engine.server_args.admin_api_key)
        if not ok:
            await ws.close(code=4401)
            return
        await ws.accept()
        sess = RealtimeSession(engine, serving, ws.send_text)
        await sess.send("session.created", session=sess.describe())
        try:
            while True:
                raw = await ws.receive_text()
                try:
                    ev = json.loads(raw)
                    assert isinstance(ev, dict)
                except Exception:
                    await sess.error("events are JSON objects")
                    continue
                await sess.handle(ev)
        except WebSocketDisconnect:
            pass
        finally:
            sess.close()
