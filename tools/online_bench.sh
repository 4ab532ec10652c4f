#!/bin/bash
# Online serving latency: launch the HTTP server (Llama-3-8B, random weights) with the multi-process frontend, then bench_serving at low and
# medium request rates.  usage: tools/online_bench.sh [extra server flags...]   -> gpurun_out/online_*.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
PORT=31077
python -m sglang_b200.launch_server --model-path dummy:llama-3-8b --port $PORT --host 127.0.0.1 --log-level warning "$@" > gpurun_out/online_server.log 2>&1 &
SRV=$!
for i in $(seq 1 120); do
  if curl -s -o /dev/null "http://127.0.0.1:$PORT/health"; then break; fi
  if ! kill -0 $SRV 2>/dev/null; then echo "server died"; tail -20 gpurun_out/online_server.log; exit 1; fi
  sleep 2
done
echo "server up after ${i}x2 s"
for rate in 1 8; do
  n=$((rate * 40)); [ $n -gt 200 ] && n=200
  timeout 300 python -m sglang_b200.bench_serving --base-url http://127.0.0.1:$PORT --dataset-name random --random-input-len 1024 --random-output-len 128 \
    --random-range-ratio 0.5 --num-prompts $n --request-rate $rate --vocab-size 128000 --output-file gpurun_out/online_rate${rate}.json 2>&1 | tail -22
done
kill $SRV; sleep 3; kill -9 $SRV 2>/dev/null
tail -5 gpurun_out/online_server.log
