"""Host-side cost of one scheduler iteration at large running batches, with the model replaced by a stub worker that returns instantly
(so everything measured is Python: batch formation, KV allocation, metadata packing excluded [it lives in the runner], result processing,
output streaming).  This is the number that bounds TP scaling: at TP=8 a bs=1024 decode step is ~3 ms on the device.
    python tools/host_step_bench.py [running=1024] [steps=200]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from sglang_b200.managers import io_struct as ios
from sglang_b200.managers.scheduler import Scheduler
from sglang_b200.model_executor.model_runner import BatchResult
from sglang_b200.server_args import ServerArgs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 200


class StubWorker:
    def get_memory_info(self):
        return {"num_pages": 40000, "max_context_len": 8192, "max_running": N}

    def forward_batch_generation(self, sb):
        r = BatchResult()
        B = len(sb.reqs)
        r.batch_size = B
        r.sample_rows = np.arange(B) if sb.sample_mask is None or sb.forward_mode.is_decode() else np.flatnonzero(sb.sample_mask)
        r.next_token_ids = np.full(B, -1, dtype=np.int64)
        r.next_token_ids[r.sample_rows] = 7
        return r

    def flush(self):
        pass

    def attach_kv_manager(self, kv):
        self.kv = kv

    def get_eos_token_ids(self):
        return [2]

    def __getattr__(self, name):          # anything optional the scheduler probes for
        raise AttributeError(name)

    def supports_overlap(self):
        return True


a = ServerArgs(model_path="dummy:tiny-llama", page_size=32, max_running_requests=N, disable_radix_cache=False, chunked_prefill_size=8192,
               max_prefill_tokens=16384, device="cpu")
outs = []
sch = Scheduler(a, StubWorker(), None, outs.append, 0)
rng = np.random.default_rng(0)
for i in range(N):
    ids = rng.integers(3, 30000, size=int(rng.integers(100, 400))).tolist()
    sch.handle_request(ios.TokenizedGenerateReqInput(rid=f"r{i}", input_text=None, input_ids=ids,
                                                     sampling_params={"max_new_tokens": 100000, "temperature": 0, "ignore_eos": True}))
while sch.waiting_queue or sch.chunked_req is not None or (sch.last_batch is not None and sch.last_batch.forward_mode.is_extend()):
    sch.step()
print("running:", len(sch.running_batch.reqs))
t0 = time.perf_counter()
for _ in range(STEPS):
    sch.step()
dt = (time.perf_counter() - t0) / STEPS
print(f"host ms per decode iteration at {len(sch.running_batch.reqs)} running requests: {dt * 1e3:.3f}")
if os.environ.get("PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50):
        sch.step()
    pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(18)
