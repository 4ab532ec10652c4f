import pickle, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sglang_b200.parallel.shm_broadcast import ShmBroadcaster, ShmHandle
h = ShmHandle(*pickle.loads(bytes.fromhex(sys.argv[1])))
r = int(sys.argv[2])
b = ShmBroadcaster.attach(h, r)
tot = 0
while True:
    o = b.broadcast(timeout=30)
    if o is None:
        break
    tot += (sum(o["x"]) + len(o["blob"])) if isinstance(o, dict) else 0
print("SUM", tot, flush=True)
b.close()
