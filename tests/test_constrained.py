"""Constrained decoding: regex / JSON-schema grammars force valid output; compressed-FSM jump-forward fires and the
jumped text still satisfies the grammar (reference kits: regex_constrained_kit / json_constrained_kit)."""
import json
import re

import numpy as np
import pytest

import sglang_b200 as sgl
from sglang_b200.constrained.regex_fsm_backend import ByteDFA, schema_to_regex
from sglang_b200.utils.tokenizer import ByteTokenizer

SCHEMA = {"type": "object", "properties": {"name": {"type": "string", "maxLength": 6}, "age": {"type": "integer"},
                                            "ok": {"type": "boolean"}}, "required": ["name", "age", "ok"]}


def test_schema_regex_and_dfa():
    rx = schema_to_regex(SCHEMA, ws="")
    assert re.fullmatch(rx, '{"name":"bob","age":42,"ok":true}')
    assert not re.fullmatch(rx, '{"name":"bob","age":"x","ok":true}')
    dfa = ByteDFA(rx)
    run, _ = dfa.forced_run(dfa.initial)
    assert run == b'{"name":"'        # compressed FSM: the whole key prefix is forced


@pytest.mark.parametrize("backend", ["regex_fsm", "xgrammar"])
def test_regex_constrained_generation(backend):
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, grammar_backend=backend)
    rx = r"(yes|no), [0-9]{3}-[a-c]+!"
    outs = eng.generate(["q1", "another question"], {"max_new_tokens": 40, "temperature": 0.8, "regex": rx, "sampling_seed": 1})
    for o in outs:
        assert re.fullmatch(rx, o["text"]), o["text"]
        assert o["meta_info"]["finish_reason"]["type"] == "stop"
    eng.scheduler.check_invariants()


@pytest.mark.parametrize("backend,jump", [("regex_fsm", True), ("regex_fsm", False), ("xgrammar", True)])
def test_json_schema_with_jump_forward(backend, jump):
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, grammar_backend=backend, enable_jump_forward=jump)
    sp = {"max_new_tokens": 120, "temperature": 0.7, "json_schema": json.dumps(SCHEMA), "sampling_seed": 3}
    outs = eng.generate(["give me json", "and another"], sp)
    for o in outs:
        if o["meta_info"]["finish_reason"]["type"] == "stop":
            obj = json.loads(o["text"], strict=False)
            assert set(obj) == {"name", "age", "ok"} and isinstance(obj["age"], int)
        else:   # hit max_new_tokens inside an unbounded field: the prefix must still be a valid JSON prefix
            assert o["text"].lstrip().startswith("{")
    st = eng.scheduler.get_internal_state()
    if jump and backend == "regex_fsm":
        # jump-forward replaces decode steps by a re-prefill: far fewer forward passes than generated tokens
        toks = sum(o["meta_info"]["completion_tokens"] for o in outs)
        assert st["forward_ct"] < toks, (st["forward_ct"], toks)
    eng.scheduler.check_invariants()


def test_invalid_grammar_aborts_only_that_request():
    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, grammar_backend="regex_fsm")
    outs = eng.generate(["a", "b"], [{"max_new_tokens": 4, "regex": "(unclosed"}, {"max_new_tokens": 4, "temperature": 0}])
    assert outs[0]["meta_info"]["finish_reason"]["type"] == "abort"
    assert outs[1]["meta_info"]["finish_reason"]["type"] == "length"
