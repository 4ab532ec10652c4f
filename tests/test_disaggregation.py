"""PD disaggregation on one box: prefill engine -> (tcp | fake) KV transfer -> decode engine == monolithic engine
(reference: test/registered/disaggregation/test_disaggregation_basic.py + pd_parity_kit)."""
import numpy as np
import pytest

import sglang_b200 as sgl
from sglang_b200.disaggregation.pd import install_disaggregation
from sglang_b200.environ import envs
from sglang_b200.managers import io_struct as ios


@pytest.mark.parametrize("backend", ["fake", "tcp"])
def test_pd_parity(backend):
    rng = np.random.default_rng(0)
    prompts = [rng.integers(3, 200, size=int(n)).tolist() for n in (9, 33, 70)]
    sp = {"max_new_tokens": 10, "temperature": 0, "ignore_eos": True}
    mono = sgl.Engine(model_path="dummy:tiny-llama", page_size=4)
    want = [o["output_ids"] for o in mono.generate(input_ids=prompts, sampling_params=sp)]

    dec = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, disaggregation_mode="decode", disaggregation_transfer_backend=backend)
    d = install_disaggregation(dec, listen_port=0 if backend == "tcp" else 8998)
    port = d.transport.port if backend == "tcp" else 8998
    pre = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, disaggregation_mode="prefill", disaggregation_transfer_backend=backend)
    p = install_disaggregation(pre)

    # the router sends the same request to both sides with a shared bootstrap room
    dec_states = []
    for i, ids in enumerate(prompts):
        g = ios.GenerateReqInput(input_ids=ids, sampling_params=dict(sp), rid=f"r{i}", bootstrap_host="127.0.0.1", bootstrap_port=port,
                                 bootstrap_room=100 + i)
        dec_states.append(dec._submit(g))
        pre._submit(ios.GenerateReqInput(input_ids=ids, sampling_params=dict(sp), rid=f"r{i}", bootstrap_host="127.0.0.1",
                                         bootstrap_port=port, bootstrap_room=100 + i))
    while pre.scheduler.has_work():
        pre.scheduler.step()
    assert p.sent == 3 and p.failed == 0
    import time

    t0 = time.time()
    while not all(s.finished for s in dec_states) and time.time() - t0 < 30:
        dec.scheduler.step()
    assert all(s.finished for s in dec_states)
    assert [s.output_ids for s in dec_states] == want
    assert d.received == 3
    assert dec.scheduler.num_prefill_tokens == 0        # the decode instance never ran a prefill
    dec.scheduler.check_invariants(); pre.scheduler.check_invariants()
    d.transport.close()


def test_pd_transfer_failure_is_reported():
    with envs.SGLB_TEST_DISAGG_FAILURE_PROB.override(1.0):
        pre = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, disaggregation_mode="prefill", disaggregation_transfer_backend="fake")
        p = install_disaggregation(pre)
        pre._submit(ios.GenerateReqInput(input_ids=[5, 6, 7, 8, 9], sampling_params={"max_new_tokens": 4}, rid="x",
                                         bootstrap_host="127.0.0.1", bootstrap_port=1, bootstrap_room=1))
        while pre.scheduler.has_work():
            pre.scheduler.step()
    assert p.failed == 1 and p.sent == 0
    pre.scheduler.check_invariants()


def test_tcp_transport_rejects_pickles_and_wrong_secret(monkeypatch):
    """the listener used to pickle.loads() the header of any inbound connection (RCE).  The wire format is now magic + JSON + HMAC:
    a pickle blob, a wrong secret and a non-allow-listed dtype are all dropped without touching the inbox."""
    import pickle
    import socket
    import struct
    import time

    import torch

    from sglang_b200.disaggregation.pd import KVPayload, TcpTransport

    monkeypatch.setenv("SGLB_DISAGG_SECRET", "s3cret")
    rx = TcpTransport(("127.0.0.1", 0))
    tx = TcpTransport()
    pay = KVPayload(room=3, rid="r", input_ids=[1, 2, 3], first_token=9, sampling_params={"max_new_tokens": 4},
                    kv=torch.arange(24, dtype=torch.bfloat16).view(1, 2, 1, 3, 2, 2), n_tokens=3)
    assert tx.send(("127.0.0.1", rx.port), pay)
    got = rx.poll()
    assert len(got) == 1 and got[0].rid == "r" and torch.equal(got[0].kv, pay.kv) and got[0].sampling_params == {"max_new_tokens": 4}

    class Boom:
        def __reduce__(self):
            return (exec, ("raise SystemExit('pickle executed')",))

    blob = pickle.dumps(Boom())
    with socket.create_connection(("127.0.0.1", rx.port), timeout=5) as c:       # the round-1 frame: u64 length + pickle
        c.sendall(struct.pack("<Q", len(blob)) + blob)
        c.settimeout(2)
        try:
            assert c.recv(2) != b"OK"
        except (socket.timeout, ConnectionError):
            pass
    assert not TcpTransport(secret="other").send(("127.0.0.1", rx.port), pay)     # signed with the wrong key: no OK comes back ...
    time.sleep(0.2)
    assert rx.poll() == []                                                          # ... and nothing reached the inbox
    rx.close()
