"""DP replicas on CPU: dispatch, identical results across replicas, replica death is detected (reference: DP controller tests)."""
import os
import signal
import time

import pytest

from sglang_b200.managers.data_parallel_controller import DataParallelController
from sglang_b200.server_args import ServerArgs


def test_dp_dispatch_and_failover():
    c = DataParallelController(ServerArgs(model_path="dummy:tiny-llama", page_size=4, dp_size=2, log_level="error"), policy="round_robin")
    try:
        sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
        a = c.generate("hello dp", sp, routed_dp_rank=0)
        b = c.generate("hello dp", sp, routed_dp_rank=1)
        assert a["output_ids"] == b["output_ids"]              # replicas hold identical weights
        outs = [c.generate(f"req {i}", sp) for i in range(4)]
        assert all(len(o["output_ids"]) == 6 for o in outs)
        infos = c.control("get_load")
        assert len(infos) == 2
        os.kill(c.procs[1].pid, signal.SIGKILL)
        t0 = time.time()
        while c.alive[1] and time.time() - t0 < 10:
            time.sleep(0.1)
        assert not c.alive[1]
        assert len(c.generate("after failover", sp)["output_ids"]) == 6   # traffic continues on the survivor
        # elastic: bring rank 1 back, add a third replica, drain and remove it again
        c.restart_replica(1)
        assert c.alive == [True, True] and c.generate("hello dp", sp, routed_dp_rank=1)["output_ids"] == a["output_ids"]
        r = c.add_replica()
        assert r == 2 and c.dp == 3 and c.generate("hello dp", sp, routed_dp_rank=2)["output_ids"] == a["output_ids"]
        c.remove_replica(2)
        assert not c.alive[2] and len(c.generate("after scale-down", sp)["output_ids"]) == 6
        # PD: both fleets derive the dp rank from the bootstrap room
        c.policy = "follow_bootstrap_room"
        with c._lock:
            picks = [c._pick(room) for room in (0, 1, 2, 3, 10, 11)]
        assert picks == [0, 1, 0, 1, 0, 1]
    finally:
        c.shutdown()
