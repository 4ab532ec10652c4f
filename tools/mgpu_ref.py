"""TP=1 reference for tools/mgpu_test.py: greedy generations + teacher-forced prompt logprobs (the latter is the numerics
oracle — with random-init weights the greedy argmax is near-degenerate, so token sequences may legitimately diverge)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import sglang_b200 as sgl
rng = np.random.default_rng(0)
prompts = [rng.integers(3, 2000, size=int(n)).tolist() for n in rng.integers(5, 300, size=12)]
eng = sgl.Engine(model_path="dummy:" + os.environ.get("MGPU_MODEL", "tiny-llama-gpu"), page_size=32, chunked_prefill_size=256, cuda_graph_max_bs=16, mem_fraction_static=0.2,
                 max_running_requests=16, log_level="warning")
outs = eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 24, "temperature": 0, "ignore_eos": True})
lp = eng.generate(input_ids=prompts, sampling_params={"max_new_tokens": 1, "temperature": 0, "ignore_eos": True}, return_logprob=True,
                  logprob_start_len=0)
logprobs = [[float(t[0]) for t in o["meta_info"]["input_token_logprobs"] if t[0] is not None] for o in lp]
json.dump({"output_ids": [o["output_ids"] for o in outs], "input_logprobs": logprobs}, open(sys.argv[1], "w"))
print("ref ok")
