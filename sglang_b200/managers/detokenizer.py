"""Incremental detokeniser (BatchTokenIDOutput -> BatchStrOutput).

Parity: reference ``python/sglang/srt/managers/detokenizer_manager.py:92`` (``handle_batch_token_id_out:431``,
incremental surrogate-safe decode ``:291``, stop-string trimming ``:177``).  Runs as a thread of the front-end
process (or in-line for the offline engine); the C++ HTTP front end can host the same logic.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List

from . import io_struct as ios


@dataclass
class DecodeStatus:
    ids: List[int] = field(default_factory=list)
    decoded_text: str = ""
    surr_offset: int = 0
    read_offset: int = 0
    sent_len: int = 0


class Detokenizer:
    def __init__(self, tokenizer: Any, max_states: int = 1 << 16):
        self.tokenizer = tokenizer
        self.states: Dict[str, DecodeStatus] = {}
        self.max_states = max_states

    def _incremental(self, st: DecodeStatus, skip_special: bool) -> str:
        """returns the full text decoded so far (stable prefix + provisional tail without broken UTF-8)"""
        tok = self.tokenizer
        surr = tok.decode(st.ids[st.surr_offset:st.read_offset], skip_special_tokens=skip_special)
        full = tok.decode(st.ids[st.surr_offset:], skip_special_tokens=skip_special)
        if len(full) > len(surr) and not full.endswith("�"):
            st.decoded_text += full[len(surr):]
            st.surr_offset = st.read_offset
            st.read_offset = len(st.ids)
            return st.decoded_text
        # tail may be an incomplete multi-byte sequence: show only the stable part
        return st.decoded_text

    def process(self, msg: ios.BatchTokenIDOutput) -> ios.BatchStrOutput:
        texts: List[str] = []
        for i, rid in enumerate(msg.rids):
            st = self.states.get(rid)
            if st is None:
                st = self.states[rid] = DecodeStatus()
            st.ids.extend(msg.output_ids[i])
            finished = msg.finished_reasons[i] is not None
            text = self._incremental(st, msg.skip_special_tokens[i])
            if finished:
                # flush whatever is left (possibly a broken tail) and trim a matched stop string
                tail = self.tokenizer.decode(st.ids[st.surr_offset:], skip_special_tokens=msg.skip_special_tokens[i])
                surr = self.tokenizer.decode(st.ids[st.surr_offset:st.read_offset], skip_special_tokens=msg.skip_special_tokens[i])
                text = st.decoded_text + tail[len(surr):]
                fr = msg.finished_reasons[i]
                matched = fr.get("matched") if isinstance(fr, dict) else None
                if isinstance(matched, str) and not msg.no_stop_trim[i]:
                    pos = text.find(matched)
                    if pos >= 0:
                        text = text[:pos]
            else:
                # never stream a partial stop string (reference detokenizer_manager.py:177)
                for s in msg.stop_strs[i] or []:
                    for k in range(min(len(s) - 1, len(text)), 0, -1):
                        if text.endswith(s[:k]):
                            text = text[: len(text) - k]
                            break
            delta = text[st.sent_len:] if len(text) >= st.sent_len else ""
            st.sent_len = max(st.sent_len, len(text))
            texts.append(delta)
            if finished:
                self.states.pop(rid, None)
        if len(self.states) > self.max_states:
            for k in list(self.states)[: len(self.states) - self.max_states]:
                self.states.pop(k, None)
        return ios.BatchStrOutput(
            rids=msg.rids, finished_reasons=msg.finished_reasons, output_strs=texts, output_ids=msg.output_ids,
            prompt_tokens=msg.prompt_tokens, completion_tokens=msg.completion_tokens, cached_tokens=msg.cached_tokens,
            output_logprobs=msg.output_logprobs, input_logprobs=msg.input_logprobs, top_logprobs=msg.top_logprobs,
            spec_verify_ct=msg.spec_verify_ct, spec_accepted_tokens=msg.spec_accepted_tokens, embeddings=msg.embeddings,
            hidden_states=msg.hidden_states, ttfts=msg.ttfts, queue_times=msg.queue_times)
