"""gRPC surface against an in-process server (reference: rust/sglang-grpc + proto/sglang/runtime/v1/*.proto)."""
import json

import pytest

import sglang_b200 as sgl


@pytest.fixture(scope="module")
def grpc_client():
    from sglang_b200.entrypoints.grpc_server import SglbGrpcClient, make_server

    eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, log_level="error")
    eng.start_background_loop()
    server, port = make_server(eng, "127.0.0.1", 0)
    server.start()
    c = SglbGrpcClient(f"127.0.0.1:{port}")
    yield c
    c.close()
    server.stop(0)
    eng.shutdown()


def test_grpc_native_rpcs(grpc_client):
    from sglang_b200.entrypoints.grpc_server import MSG

    c = grpc_client
    assert c.call("HealthCheck", MSG["Empty"]()).healthy
    assert json.loads(c.call("GetModelInfo", MSG["Empty"]()).json)["architecture"] == "LlamaForCausalLM"
    assert c.call("ListModels", MSG["Empty"]()).models == ["dummy:tiny-llama"]
    tk = c.call("Tokenize", MSG["TokenizeRequest"](text="hello"))
    assert tk.count == len(tk.tokens) > 0
    assert c.call("Detokenize", MSG["DetokenizeRequest"](tokens=list(tk.tokens))).text.endswith("hello")
    sp = MSG["SamplingParams"](max_new_tokens=5, temperature=0.0, ignore_eos=True)
    outs = list(c.call("TextGenerate", MSG["TextGenerateRequest"](text="hello", sampling_params=sp)))
    assert len(outs) == 1 and outs[0].finished and len(outs[0].output_ids) == 5 and outs[0].meta.completion_tokens == 5
    chunks = list(c.call("TextGenerate", MSG["TextGenerateRequest"](text="hello", sampling_params=sp, stream=True)))
    assert chunks[-1].finished and [t for ch in chunks for t in ch.output_ids] == list(outs[0].output_ids)
    assert "".join(ch.text for ch in chunks) == outs[0].text
    ids = list(c.call("Generate", MSG["GenerateRequest"](input_ids=list(tk.tokens), sampling_params=sp)))[0]
    ids_s = list(c.call("Generate", MSG["GenerateRequest"](input_ids=list(tk.tokens), sampling_params=sp, stream=True)))
    assert len(ids.output_ids) == 5 and [t for ch in ids_s for t in ch.output_ids] == list(ids.output_ids)
    import time
    for _ in range(50):       # the scheduler retires the last request on its next iteration
        ld = c.call("GetLoad", MSG["Empty"]())
        if ld.num_running_reqs == 0:
            break
        time.sleep(0.05)
    assert ld.num_running_reqs == 0
    assert c.call("FlushCache", MSG["Empty"]()).success
    assert c.call("Abort", MSG["AbortRequest"](abort_all=True)).success
    assert json.loads(c.call("GetServerInfo", MSG["Empty"]()).json)["internal_states"][0]["page_size"] == 4


def test_grpc_openai_passthrough_and_errors(grpc_client):
    import grpc

    from sglang_b200.entrypoints.grpc_server import MSG, emit_proto

    c = grpc_client
    body = {"model": "m", "messages": [{"role": "user", "content": "ping"}], "max_tokens": 4, "temperature": 0}
    out = list(c.call("ChatComplete", MSG["OpenAIRequest"](json=json.dumps(body))))
    assert out[-1].done and json.loads(out[0].json)["usage"]["completion_tokens"] == 4
    out = list(c.call("ChatComplete", MSG["OpenAIRequest"](json=json.dumps({**body, "stream": True}))))
    assert out[-1].done and json.loads(out[0].json)["choices"][0]["delta"]["role"] == "assistant"
    out = list(c.call("Complete", MSG["OpenAIRequest"](json=json.dumps({"model": "m", "prompt": "hi", "max_tokens": 3}))))
    assert json.loads(out[0].json)["usage"]["completion_tokens"] == 3
    with pytest.raises(grpc.RpcError) as ei:
        list(c.call("TextGenerate", MSG["TextGenerateRequest"](text="x", sampling_params=MSG["SamplingParams"](top_p=7.0))))
    assert ei.value.code() in (grpc.StatusCode.INVALID_ARGUMENT, grpc.StatusCode.INTERNAL)
    proto = emit_proto()
    assert "service SglbService" in proto and "rpc TextGenerate(TextGenerateRequest) returns (stream TextGenerateResponse);" in proto
    assert open("proto/sglb/runtime/v1/sglb.proto").read() == proto, "checked-in .proto is stale: re-run --emit-proto"
