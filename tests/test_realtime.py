"""/v1/realtime WebSocket: transcription sessions on a speech model, text responses on a language model."""
import base64
import json

import numpy as np
from fastapi.testclient import TestClient

import sglang_b200 as sgl
from sglang_b200.entrypoints.http_server import create_app


def _recv_until(ws, type_):
    got = []
    while True:
        ev = json.loads(ws.receive_text())
        got.append(ev)
        if ev["type"] == type_:
            return got


def test_transcription_session_streams_deltas_and_matches_direct_call():
    eng = sgl.Engine(model_path="dummy:tiny-whisper", device="cpu", page_size=4, log_level="error")
    t = np.arange(16000 * 2) / 16000.0
    wave = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    pcm = (wave * 32767).astype("<i2")
    direct = eng.generate(input_ids=[3, 10, 11], audio_data=(pcm.astype(np.float32) / 32768.0, 16000), sampling_params={"max_new_tokens": 6, "temperature": 0})
    with TestClient(create_app(eng)) as c, c.websocket_connect("/v1/realtime") as ws:
        first = json.loads(ws.receive_text())
        assert first["type"] == "session.created" and first["session"]["object"] == "realtime.transcription_session"
        ws.send_text(json.dumps({"type": "session.update", "session": {"max_output_tokens": 6, "audio": {"input": {
            "format": {"type": "audio/pcm", "rate": 16000}, "transcription": {"language": "en", "prompt_ids": [3, 10, 11]}}}}}))
        assert json.loads(ws.receive_text())["type"] == "session.updated"
        ws.send_text(json.dumps({"type": "input_audio_buffer.commit"}))
        assert json.loads(ws.receive_text())["type"] == "error"                              # nothing buffered yet
        raw = pcm.tobytes()
        for i in range(0, len(raw), 16000):                                                  # audio arrives in chunks
            ws.send_text(json.dumps({"type": "input_audio_buffer.append", "audio": base64.b64encode(raw[i:i + 16000]).decode()}))
        ws.send_text(json.dumps({"type": "input_audio_buffer.commit"}))
        evs = _recv_until(ws, "conversation.item.input_audio_transcription.completed")
        types = [e["type"] for e in evs]
        assert types[:2] == ["input_audio_buffer.committed", "conversation.item.created"]
        deltas = [e["delta"] for e in evs if e["type"].endswith(".delta")]
        done = evs[-1]
        assert deltas and "".join(deltas) == done["transcript"] == direct["text"]
        assert done["usage"]["output_tokens"] == 6 and done["item_id"] == evs[0]["item_id"]
        ws.send_text(json.dumps({"type": "response.create"}))
        assert json.loads(ws.receive_text())["type"] == "error"                              # transcription sessions do not chat
        ws.send_text(json.dumps({"type": "input_audio_buffer.clear"}))
        assert json.loads(ws.receive_text())["type"] == "input_audio_buffer.cleared"
    eng.shutdown()


def test_text_session_multi_turn_response_stream():
    eng = sgl.Engine(model_path="dummy:tiny-llama", device="cpu", page_size=4, log_level="error")
    with TestClient(create_app(eng)) as c, c.websocket_connect("/v1/realtime") as ws:
        assert json.loads(ws.receive_text())["session"]["object"] == "realtime.session"
        ws.send_text(json.dumps({"type": "session.update", "session": {"instructions": "be brief", "max_output_tokens": 5}}))
        assert json.loads(ws.receive_text())["session"]["max_output_tokens"] == 5
        ws.send_text(json.dumps({"type": "conversation.item.create", "item": {"type": "message", "role": "user", "content": [{"type": "input_text", "text": "hi"}]}}))
        assert json.loads(ws.receive_text())["type"] == "conversation.item.created"
        ws.send_text(json.dumps({"type": "response.create"}))
        evs = _recv_until(ws, "response.done")
        types = [e["type"] for e in evs]
        assert types[0] == "response.created" and "response.output_item.added" in types and "response.text.done" in types
        text = "".join(e["delta"] for e in evs if e["type"] == "response.text.delta")
        done = evs[-1]["response"]
        assert done["status"] == "completed" and done["usage"]["output_tokens"] == 5 and done["output"][0]["content"][0]["text"] == text
        ws.send_text(json.dumps({"type": "response.create"}))                                 # second turn sees the assistant's reply in history
        evs2 = _recv_until(ws, "response.done")
        assert evs2[-1]["response"]["usage"]["input_tokens"] > done["usage"]["input_tokens"]
        ws.send_text("not json")
        assert json.loads(ws.receive_text())["type"] == "error"
        ws.send_text(json.dumps({"type": "bogus.event"}))
        assert "unknown event" in json.loads(ws.receive_text())["error"]["message"]
    eng.shutdown()
