"""TP=2 in background-loop (server) mode: rank 0 submits requests through its inbox, every rank runs the scheduler thread and
receives each iteration's batch over the shared-memory broadcast ring."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sglang_b200 as sgl
rank = int(os.environ["RANK"])
eng = sgl.Engine(model_path="dummy:tiny-llama", page_size=4, tp_size=int(os.environ["WORLD_SIZE"]), device="cpu", log_level="error")
eng.start_background_loop()
assert (eng._shm_bcast is not None) == (os.environ.get("SGLB_SHM_BROADCAST", "1") != "0")
if rank == 0:
    sp = {"max_new_tokens": 6, "temperature": 0, "ignore_eos": True}
    outs = [eng.generate(input_ids=p, sampling_params=sp) for p in ([5, 6, 7, 8, 9], [3, 4], list(range(20, 50)))]
    outs += eng.generate(input_ids=[[9, 8, 7], [1, 2, 3, 4]], sampling_params=sp)
    print("RESULT " + json.dumps([o["output_ids"] for o in outs]), flush=True)
    eng.shutdown()
    os._exit(0)
time.sleep(float(os.environ.get("PEER_LIFETIME", "20")))
os._exit(0)
