import json, sys, faulthandler
import numpy as np
import sglang_b200 as sgl
rng = np.random.default_rng(3)
shared = rng.integers(5, 2000, size=200).tolist()
prompts = [shared + rng.integers(5, 2000, size=int(n)).tolist() for n in (3, 40, 300)] + [rng.integers(5, 2000, size=int(n)).tolist() for n in (5, 133, 700)]
sp = {"max_new_tokens": 48, "temperature": 0, "ignore_eos": True}
kw = dict(model_path="dummy:tiny-swa-gpu", page_size=32, chunked_prefill_size=256, log_level="error", max_running_requests=8, mem_fraction_static=0.2, cuda_graph_max_bs=8)
ref = sgl.Engine(**kw)
want = [o["output_ids"] for o in ref.generate(input_ids=prompts, sampling_params=sp)]
ref.shutdown()
print("ref done", flush=True)
eng = sgl.Engine(swa_full_tokens_ratio=0.5, **kw)
import threading, time, os
def mon():
    for i in range(8):
        time.sleep(5)
        s = eng.scheduler
        print("t", 5 * (i + 1), "fwd", s.forward_ct, "iters", s.step_calls, "retracted", s.num_retracted, "running", len(s.running_batch.reqs), "waiting",
              len(s.waiting_queue), "chunked", s.chunked_req is not None, "pending", len(s.pending), "swa free", s.swa_allocator.available_pages(), "/",
              s.swa_allocator.total_pages, "full free", s.allocator.available_pages(), "slots", s.kv.slots.available(), flush=True)
    faulthandler.dump_traceback(all_threads=True)
    sys.stderr.flush()
    os._exit(3)
threading.Thread(target=mon, daemon=True).start()
got = [o["output_ids"] for o in eng.generate(input_ids=prompts, sampling_params=sp)]
eng.scheduler.check_invariants()
print("pool done", sum(a == b for a, b in zip(got, want)), eng.scheduler.kv.swa_slid_pages, flush=True)
got2 = [o["output_ids"] for o in eng.generate(input_ids=prompts[:3], sampling_params=sp)]
print("second", sum(a == b for a, b in zip(got2, want[:3])), eng.scheduler.num_cached_tokens)
