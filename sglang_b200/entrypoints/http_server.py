"""HTTP server: native routes + OpenAI-compatible routes over one ``Engine``.

Parity: reference ``python/sglang/srt/entrypoints/http_server.py`` (routes ``:654-2041``: /health, /health_generate,
/get_model_info, /get_server_info, /get_load, /flush_cache, /generate, /encode, /abort_request, profiling,
weight-update family, LoRA load/unload, sessions, pause/continue, /metrics, /v1/*; admin-key guard; warm-up request
``_execute_server_warmup:2186``; ``launch_server:2766``).

Process model: rank 0 hosts FastAPI + tokenizer/detokenizer + scheduler thread; ranks 1..N-1 (TP/PP peers) are spawned
as plain worker processes running the same scheduler loop and receive each iteration's requests by Gloo broadcast.
"""
import asyncio
import json
import logging
import multiprocessing as mp
import os
import threading
import socket
import sys
import time
from typing import Any, Optional

from fastapi import FastAPI, Request
from fastapi.responses import JSONResponse, Response, StreamingResponse

from ..managers import io_struct as ios
from ..server_args import ServerArgs, prepare_server_args
from .openai import protocol as P
from .openai.serving import OpenAIServing

logger = logging.getLogger(__name__)


def create_app(engine: Any):
    app = FastAPI(title="sglang_b200")
    serving = OpenAIServing(engine)
    args: ServerArgs = engine.server_args
    app.state.engine = engine

    def err(msg: str, code: int = 400) -> JSONResponse:
        return JSONResponse({"error": {"message": msg, "type": "BadRequestError", "code": code}}, status_code=code)

    @app.middleware("http")
    async def auth(request: Request, call_next):
        # deny by default (utils/auth.py): every route except /health*, /ping and /metrics needs a key once one is configured
        from ..utils import auth as _auth

        ok, code, msg = _auth.decide(request.method, request.url.path, request.headers.get("authorization"), args.api_key,
                                     args.admin_api_key)
        if not ok:
            return err(msg, code)
        return await call_next(request)

    # ------------------------------------------------------------------ health / info
    @app.get("/health")
    async def health():
        return Response(status_code=200)

    @app.get("/health_generate")
    async def health_generate():
        try:
            out = await asyncio.wait_for(engine.async_generate(input_ids=[1, 2, 3], sampling_params={"max_new_tokens": 1, "temperature": 0}), 30)
            return Response(status_code=200 if out["meta_info"]["completion_tokens"] >= 1 else 503)
        except Exception:
            return Response(status_code=503)

    @app.get("/get_model_info")
    async def model_info():
        return {"model_path": args.model_path, "tokenizer_path": args.tokenizer_path, "is_generation": not args.is_embedding,
                "architecture": engine.model_config.architecture, "weight_version": engine.weight_version}

    @app.get("/get_server_info")
    async def server_info():
        return await _ctl(engine.get_server_info)

    @app.get("/get_load")
    @app.get("/v1/loads")
    async def get_load():
        return await _ctl(engine.get_load)

    @app.get("/metrics")
    async def metrics():
        if engine.metrics is None:
            return err("metrics are disabled (--enable-metrics)", 404)
        st = dict(engine.scheduler.get_internal_state())
        fe = getattr(app.state, "native_frontend", None)
        if fe is not None:          # the C++ socket layer's own counters (connections, requests, probes it answered itself, bytes)
            st.update({"http_" + k: v for k, v in fe.stats().items()})
        return Response(engine.metrics.render(st), media_type="text/plain; version=0.0.4")

    async def _ctl(fn, *a):
        return await asyncio.get_running_loop().run_in_executor(None, fn, *a)

    # ------------------------------------------------------------------ native generate
    @app.post("/generate")
    async def generate(request: Request):
        body = await request.json()
        allowed = {f for f in ios.GenerateReqInput.__dataclass_fields__}
        kw = {k: v for k, v in body.items() if k in allowed}
        stream = bool(kw.pop("stream", False))
        prompt = kw.pop("text", None)
        try:
            if stream:
                gen = await engine.async_generate(prompt, stream=True, **kw)

                async def sse():
                    async for o in gen:
                        yield "data: " + json.dumps(o) + "\n\n"
                    yield "data: [DONE]\n\n"

                return StreamingResponse(sse(), media_type="text/event-stream")
            return await engine.async_generate(prompt, **kw)
        except ValueError as e:
            return err(str(e))

    @app.post("/encode")
    async def encode(request: Request):
        body = await request.json()
        return await _ctl(engine.encode, body.get("text"))

    @app.post("/abort_request")
    async def abort(request: Request):
        body = await request.json()
        await _ctl(engine.abort_request, body.get("rid", ""), bool(body.get("abort_all", False)))
        return {"success": True}

    @app.api_route("/flush_cache", methods=["GET", "POST"])
    async def flush():
        ok = await _ctl(engine.flush_cache)
        return Response("Cache flushed.\n" if ok else "Flush refused: requests in flight.\n", status_code=200 if ok else 400)

    # ------------------------------------------------------------------ admin: profiling / weights / memory / LoRA / sessions
    @app.api_route("/start_profile", methods=["GET", "POST"])
    async def start_profile(request: Request):
        body = await request.json() if request.method == "POST" and await request.body() else {}
        r = await _ctl(engine.start_profile, body.get("output_dir"), body.get("num_steps"), body.get("activities"))
        return {"success": r.success, "message": r.message}

    @app.api_route("/stop_profile", methods=["GET", "POST"])
    async def stop_profile():
        r = await _ctl(engine.stop_profile)
        return {"success": r.success, "message": r.message}

    @app.post("/update_weights_from_disk")
    async def update_disk(request: Request):
        b = await request.json()
        r = await _ctl(engine.update_weights_from_disk, b["model_path"], b.get("load_format"))
        return JSONResponse({"success": r.success, "message": r.message}, status_code=200 if r.success else 400)

    @app.post("/update_weights_from_tensor")
    async def update_tensor(request: Request):
        b = await request.json()
        r = await _ctl(engine.update_weights_from_tensor, [(k, v) for k, v in b["named_tensors"]], b.get("flush_cache", True))
        return JSONResponse({"success": r.success, "message": r.message}, status_code=200 if r.success else 400)

    @app.api_route("/get_weights_by_name", methods=["GET", "POST"])
    async def get_weights(request: Request):
        b = await request.json()
        return await _ctl(engine.get_weights_by_name, b["name"], b.get("truncate_size", 100))

    @app.api_route("/hicache/storage-backend", methods=["GET", "PUT", "DELETE"])
    async def hicache_storage_backend(request: Request):
        """inspect / attach / detach the L3 storage of the hierarchical cache at run time (reference http_server.py:1078-1141)"""
        from ..mem_cache.hicache_storage import create_storage

        tree = engine.scheduler.tree
        if not hasattr(tree, "storage"):
            return JSONResponse({"success": False, "message": "hierarchical cache is not enabled"}, status_code=400)
        if request.method == "GET":
            return {"attached": tree.storage is not None, "stats": tree.storage.stats() if tree.storage is not None else {}}
        if request.method == "DELETE":
            if tree.storage is not None:
                tree.storage.close()
            tree.storage = None
            return {"success": True, "message": "detached"}
        b = await request.json()
        try:
            new = create_storage(b["backend"], b.get("config") or {})
        except (ValueError, KeyError) as e:
            return JSONResponse({"success": False, "message": str(e)}, status_code=400)
        if tree.storage is not None:
            tree.storage.close()
        tree.storage = new
        return {"success": True, "message": f"attached {b['backend']}"}

    @app.post("/send_weights_to_remote_instance")
    async def send_weights_to_remote_instance(request: Request):
        b = await request.json()
        r = await _ctl(engine.send_weights_to_remote_instance, b["master_address"], b["master_port"], b.get("group_name", "remote_instance_load"),
                       b.get("backend", "nccl"))
        return JSONResponse({"success": r.success, "message": r.message}, status_code=200 if r.success else 400)

    @app.api_route("/weights_checker", methods=["GET", "POST"])
    async def weights_checker(request: Request):
        b = await request.json() if await request.body() else {}
        r = await _ctl(engine.weights_checker, b.get("action", "checksum"))
        return JSONResponse({"success": r.success, "message": r.message}, status_code=200 if r.success else 400)

    def _rpc_route(path: str, method: str):
        async def handler(request: Request):
            b = await request.json() if await request.body() else {}
            try:
                return {"success": True, "result": await _ctl(lambda: engine.collective_rpc(method, **b))}
            except Exception as e:      # noqa: BLE001 - surfaced to the operator
                return JSONResponse({"success": False, "message": f"{type(e).__name__}: {e}"}, status_code=400)

        app.add_api_route(path, handler, methods=["POST"])

    # expert placement: live EPLB re-plan (optionally with redundant slots), host backups, and rank-fault recovery (eplb/placement.py)
    _rpc_route("/eplb_rebalance", "eplb_rebalance")
    _rpc_route("/build_expert_backup", "build_expert_backup")
    _rpc_route("/elastic_ep_recover", "elastic_ep_recover")

    @app.api_route("/slow_down", methods=["GET", "POST"])
    async def slow_down(request: Request):
        b = await request.json() if await request.body() else {}
        r = await _ctl(engine.slow_down, b.get("forward_sleep_time"))
        return {"success": r.success, "message": r.message}

    @app.api_route("/set_trace_level", methods=["GET", "POST"])
    async def set_trace_level(request: Request):
        from ..observability.trace import get_tracer, init_tracing

        b = await request.json() if await request.body() else {}
        lv = int(b.get("level", request.query_params.get("level", 1)))
        (get_tracer() or init_tracing(lv)).set_level(lv)
        return {"success": True, "level": lv}

    @app.api_route("/release_memory_occupation", methods=["GET", "POST"])
    async def release(request: Request):
        b = await request.json() if await request.body() else {}
        r = await _ctl(engine.release_memory_occupation, b.get("tags"))
        return {"success": r.success, "message": r.message}

    @app.api_route("/resume_memory_occupation", methods=["GET", "POST"])
    async def resume(request: Request):
        b = await request.json() if await request.body() else {}
        r = await _ctl(engine.resume_memory_occupation, b.get("tags"))
        return {"success": r.success, "message": r.message}

    @app.post("/load_lora_adapter")
    async def load_lora(request: Request):
        b = await request.json()
        r = await _ctl(engine.load_lora_adapter, b["lora_name"], b["lora_path"])
        return JSONResponse({"success": r.success, "message": r.message}, status_code=200 if r.success else 400)

    @app.post("/unload_lora_adapter")
    async def unload_lora(request: Request):
        b = await request.json()
        r = await _ctl(engine.unload_lora_adapter, b["lora_name"])
        return JSONResponse({"success": r.success, "message": r.message}, status_code=200 if r.success else 400)

    @app.api_route("/open_session", methods=["GET", "POST"])
    async def open_session(request: Request):
        b = await request.json() if await request.body() else {}
        return await _ctl(engine.open_session, b.get("capacity_of_str_len", 0), b.get("session_id"))

    @app.api_route("/close_session", methods=["GET", "POST"])
    async def close_session(request: Request):
        b = await request.json()
        await _ctl(engine.close_session, b["session_id"])
        return Response(status_code=200)

    @app.post("/pause_generation")
    async def pause(request: Request):
        b = await request.json() if await request.body() else {}
        await _ctl(engine.pause_generation, b.get("mode", "abort"))
        return {"success": True}

    @app.post("/continue_generation")
    async def cont():
        await _ctl(engine.continue_generation)
        return {"success": True}

    @app.api_route("/configure_logging", methods=["GET", "POST"])
    async def configure_logging(request: Request):
        b = await request.json() if await request.body() else {}
        if "log_level" in b:
            logging.getLogger().setLevel(getattr(logging, str(b["log_level"]).upper(), logging.INFO))
        if "log_requests" in b:
            args.log_requests = bool(b["log_requests"])
        return Response(status_code=200)

    # ------------------------------------------------------------------ OpenAI-compatible
    @app.get("/v1/models")
    async def v1_models():
        return serving.models()

    @app.post("/v1/completions")
    async def v1_completions(r: P.CompletionRequest):
        try:
            out = await serving.completions(r)
        except ValueError as e:
            return err(str(e))
        return StreamingResponse(out, media_type="text/event-stream") if r.stream else out

    @app.post("/v1/chat/completions")
    async def v1_chat(r: P.ChatCompletionRequest):
        try:
            out = await serving.chat_completions(r)
        except ValueError as e:
            return err(str(e))
        return StreamingResponse(out, media_type="text/event-stream") if r.stream else out

    @app.post("/v1/embeddings")
    async def v1_embeddings(r: P.EmbeddingRequest):
        return await serving.embeddings(r)

    @app.post("/v1/tokenize")
    @app.post("/tokenize")
    async def v1_tokenize(r: P.TokenizeRequest):
        return serving.tokenize(r)

    @app.post("/v1/detokenize")
    @app.post("/detokenize")
    async def v1_detokenize(r: P.DetokenizeRequest):
        return serving.detokenize(r)

    @app.post("/v1/score")
    async def v1_score(r: P.ScoringRequest):
        s = await _ctl(engine.score, r.query, r.items, r.label_token_ids, r.apply_softmax)
        return {"scores": s, "model": r.model, "object": "scoring"}

    @app.post("/v1/rerank")
    async def v1_rerank(r: P.RerankRequest):
        res = await _ctl(engine.rerank, r.query, r.documents)
        return res[: r.top_n] if r.top_n else res

    @app.get("/kv_events")
    async def kv_events():
        return await _ctl(engine.take_kv_events)

    @app.post("/init_weights_update_group")
    async def init_weights_update_group(request: Request):
        b = await request.json()
        r = await _ctl(engine.init_weights_update_group, b["master_address"], int(b["master_port"]), int(b.get("rank_offset", 1)),
                       int(b["world_size"]), b.get("group_name", "weight_update_group"), b.get("backend", "nccl"))
        return {"success": r.success, "message": r.message}

    @app.post("/update_weights_from_distributed")
    async def update_weights_from_distributed(request: Request):
        b = await request.json()
        r = await _ctl(engine.update_weights_from_distributed, b["names"], b["dtypes"], b["shapes"], b.get("group_name", "weight_update_group"),
                       bool(b.get("flush_cache", True)))
        return {"success": r.success, "message": r.message}

    @app.post("/destroy_weights_update_group")
    async def destroy_weights_update_group(request: Request):
        b = await request.json()
        r = await _ctl(engine.destroy_weights_update_group, b.get("group_name", "weight_update_group"))
        return {"success": r.success, "message": r.message}

    from .adapters import install_adapters

    install_adapters(app, engine, serving)
    from .realtime import install_realtime

    install_realtime(app, engine, serving)
    return app


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _peer_main(args: ServerArgs, rank: int, world: int, master_port: int, master_addr: str = "127.0.0.1", local_rank: Optional[int] = None) -> None:
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank if local_rank is None else local_rank), WORLD_SIZE=str(world), MASTER_ADDR=master_addr,
                      MASTER_PORT=str(master_port))
    from ..utils.process_tree import die_with_parent

    die_with_parent()          # a killed rank-0 server must not leave GPU-holding peers behind
    from .engine import Engine

    eng = Engine(server_args=args)
    eng.start_background_loop()
    eng._bg_thread.join()


def spawn_peers(server_args: ServerArgs) -> list:
    """start this node's TP x PP ranks > 0 as worker processes and export RANK / MASTER_* for the calling (first local) rank"""
    world = server_args.tp_size * server_args.pp_size
    peers = []
    if world > 1 and "RANK" not in os.environ:
        # multi-node (reference --nnodes / --node-rank / --dist-init-addr): node i owns the consecutive global ranks
        # [i * world / nnodes, (i + 1) * world / nnodes); every node spawns its own ranks, rendezvous at dist_init_addr; node 0 serves HTTP
        nn, nr = max(1, server_args.nnodes), server_args.node_rank
        if world % nn:
            raise ValueError("tp_size * pp_size must be divisible by nnodes")
        per = world // nn
        if nn > 1:
            if not server_args.dist_init_addr:
                raise ValueError("--nnodes > 1 needs --dist-init-addr host:port")
            addr, port = server_args.dist_init_addr.rsplit(":", 1)
            port = int(port)
        else:
            addr, port = "127.0.0.1", server_args.nccl_port or _free_port()
        first = nr * per
        ctx = mp.get_context("spawn")
        for lr in range(1, per):
            p = ctx.Process(target=_peer_main, args=(server_args, first + lr, world, port, addr, lr), daemon=True)
            p.start()
            peers.append(p)
        os.environ.update(RANK=str(first), LOCAL_RANK="0", WORLD_SIZE=str(world), MASTER_ADDR=addr, MASTER_PORT=str(port))
    return peers


def launch_server(server_args: ServerArgs, ready_event: Optional[Any] = None) -> None:
    import uvicorn

    from .engine import Engine

    if "RANK" not in os.environ and not os.environ.get("SGLB_IN_SCHEDULER_PROCESS") and threading.current_thread() is threading.main_thread():
        from ..utils.process_tree import register_server

        register_server(server_args.port)          # `python -m sglang_b200.cli killall` stops registered servers by pid

    if int(getattr(server_args, "tokenizer_worker_num", 1) or 1) > 1 and "RANK" not in os.environ and not os.environ.get("SGLB_IN_SCHEDULER_PROCESS"):
        from ..managers.multiprocess import serve_multi_tokenizer

        serve_multi_tokenizer(server_args)       # supervisor: N HTTP + tokenizer workers | scheduler | detokenizer
        return
    if server_args.enable_multi_process_frontend and "RANK" not in os.environ and not os.environ.get("SGLB_IN_SCHEDULER_PROCESS"):
        # this process keeps HTTP + tokenizer + request states; the scheduler (and through it the TP x PP peers) and the detokenizer run in
        # their own processes (reference engine.py:905 _launch_subprocesses)
        from ..managers.multiprocess import FrontendEngine

        engine = FrontendEngine(server_args)
        _serve(engine, server_args, ready_event, [])
        return
    peers = spawn_peers(server_args)
    engine = Engine(server_args=server_args)
    engine.start_background_loop()
    if int(os.environ.get("RANK", 0)) != 0:
        engine._bg_thread.join()
        return
    _serve(engine, server_args, ready_event, peers)


def _serve(engine: Any, server_args: ServerArgs, ready_event: Optional[Any], peers: list) -> None:
    import uvicorn

    app = create_app(engine)
    engine.exit_on_watchdog = True          # a server whose scheduler stopped making progress dumps its requests and goes down

    @app.on_event("startup")
    async def _warm():
        try:   # one real request through the whole stack (reference _wait_and_warmup)
            await engine.async_generate(input_ids=[1, 2, 3, 4], sampling_params={"max_new_tokens": 2, "temperature": 0})
            logger.info("The server is fired up and ready to roll!")
        except Exception as e:
            logger.error("warm-up request failed: %s", e)
        if ready_event is not None:
            ready_event.set()

    from ..utils.process_tree import SubprocessWatchdog, graceful_sigterm, kill_process_tree

    graceful_sigterm()         # uvicorn re-raises the SIGTERM it captured once its loop stopped: make that run the clean-up below
    children = list(peers) + [p for p in (getattr(engine, "_p_sched", None), getattr(engine, "_p_detok", None)) if p is not None]
    dog = SubprocessWatchdog(children) if children else None
    if dog is not None:
        dog.start()
    try:
        # graceful drain: on SIGTERM uvicorn closes the listener and waits for the in-flight HTTP requests (= running generations) to complete
        if getattr(server_args, "http_frontend", "uvicorn") == "native":
            # C++ epoll front end: sockets, HTTP framing, SSE chunking and liveness probes never take the GIL (entrypoints/native_http.py)
            from .native_http import NativeAsgiServer

            srv = NativeAsgiServer(app, server_args.host, server_args.port, graceful_timeout=max(1.0, float(server_args.graceful_shutdown_timeout)),
                                   static=None if server_args.api_key else {"/health": (200, "application/json", b"{}")})
            app.state.native_frontend = srv.fe
            srv.run()
        else:
            uvicorn.run(app, host=server_args.host, port=server_args.port, log_level=server_args.log_level, timeout_keep_alive=5,
                        timeout_graceful_shutdown=max(1, int(server_args.graceful_shutdown_timeout)))
    finally:
        if dog is not None:
            dog.stop()
        try:
            engine.shutdown()
        finally:
            kill_process_tree(children, grace_s=3.0)


def main(argv=None) -> None:
    launch_server(prepare_server_args(sys.argv[1:] if argv is None else argv))


if __name__ == "__main__":
    main()
