"""Frontend backends: ``RuntimeEndpoint`` (HTTP to a running server), ``Runtime`` (spawns the server), and
``EngineBackend`` (in-process engine; no HTTP hop).

Parity: reference ``python/sglang/lang/backend/runtime_endpoint.py:26`` (``cache_prefix:80``, ``generate:159``,
``generate_stream``, ``select:248`` via prompt logprobs, ``concatenate_and_append:317``) and ``Runtime:356``.
"""
from __future__ import annotations

import atexit
import json
import os
import socket
import subprocess
import sys
import time
from typing import Any, Dict, Iterator, List, Optional, Tuple

from ..chat_template import get_chat_template, get_chat_template_by_model_path
from ..choices import ChoicesDecision, ChoicesSamplingMethod, token_length_normalized
from ..ir import SglSamplingParams


class BaseBackend:
    support_concate_and_append = False

    def __init__(self) -> None:
        self.chat_template = get_chat_template("default")

    def get_chat_template(self):
        return self.chat_template

    def cache_prefix(self, prefix_str: str) -> None:
        pass

    def end_program(self, s: Any) -> None:
        pass

    def commit_lazy_operations(self, s: Any) -> None:
        pass

    def flush_cache(self) -> bool:
        return True

    def get_server_info(self) -> Any:
        return {}

    # subclasses implement _generate_raw(payload) -> dict | list and _generate_stream(payload)
    def _payload(self, s: Any, sp: SglSamplingParams, **extra: Any) -> Dict[str, Any]:
        p: Dict[str, Any] = {"text": s.prompt, "sampling_params": {"skip_special_tokens": True, **sp.to_srt_kwargs()}}
        if s.images:
            p["image_data"] = s.images if len(s.images) > 1 else s.images[0]
        for k in ("return_logprob", "logprob_start_len", "top_logprobs_num"):
            v = getattr(sp, k)
            if v is not None:
                p[k] = v
        p.update(extra)
        return p

    def generate(self, s: Any, sampling_params: SglSamplingParams) -> Tuple[Any, Dict[str, Any]]:
        obj = self._generate_raw(self._payload(s, sampling_params))
        if isinstance(obj, list):
            return [o["text"] for o in obj], [o["meta_info"] for o in obj]
        return obj["text"], obj["meta_info"]

    def generate_stream(self, s: Any, sampling_params: SglSamplingParams) -> Iterator[Tuple[str, Dict[str, Any]]]:
        pos = 0
        for o in self._generate_stream(self._payload(s, sampling_params, stream=True)):
            chunk = o["text"][pos:]
            pos = len(o["text"])
            yield chunk, o["meta_info"]

    def select(self, s: Any, choices: List[str], temperature: float, choices_method: Optional[ChoicesSamplingMethod]) -> ChoicesDecision:
        """score each choice by the logprobs of its tokens given the current text (server-side prefix reuse makes the
        shared prompt free after the first choice)"""
        assert temperature <= 1e-5, "select only supports greedy decisions"
        method = choices_method or token_length_normalized
        # 1. cache the prefix and learn its token length
        base = self._generate_raw({"text": s.prompt, "sampling_params": {"max_new_tokens": 0, "temperature": 0}})
        prompt_len = base["meta_info"]["prompt_tokens"]
        # 2. prompt logprobs of every continuation
        objs = self._generate_raw({"text": [s.prompt + c for c in choices], "sampling_params": {"max_new_tokens": 0, "temperature": 0},
                                   "return_logprob": True, "logprob_start_len": max(prompt_len - 2, 0)})
        input_lp, norm = [], []
        for o in objs:
            lps = [t for t in o["meta_info"].get("input_token_logprobs", []) if t[0] is not None]
            n_choice = max(1, o["meta_info"]["prompt_tokens"] - prompt_len)
            lps = lps[-n_choice:]
            input_lp.append(lps)
            norm.append(sum(t[0] for t in lps) / max(1, len(lps)))
        uncond = None
        if method.requires_unconditional_logprobs:
            u = self._generate_raw({"text": list(choices), "sampling_params": {"max_new_tokens": 0, "temperature": 0},
                                    "return_logprob": True, "logprob_start_len": 0})
            uncond = [o["meta_info"].get("input_token_logprobs", []) for o in u]
        return method(choices=choices, normalized_prompt_logprobs=norm, input_token_logprobs=input_lp,
                      output_token_logprobs=[[] for _ in choices], unconditional_token_logprobs=uncond)


class RuntimeEndpoint(BaseBackend):
    def __init__(self, base_url: str, api_key: Optional[str] = None, verify: Optional[str] = None, chat_template_name: Optional[str] = None):
        super().__init__()
        import requests

        self.base_url = base_url.rstrip("/")
        self.s = requests.Session()
        if api_key:
            self.s.headers["Authorization"] = f"Bearer {api_key}"
        info = self.s.get(self.base_url + "/get_model_info", timeout=30).json()
        self.model_info = info
        self.chat_template = get_chat_template(chat_template_name) if chat_template_name else get_chat_template_by_model_path(info["model_path"])

    def get_model_name(self) -> str:
        return self.model_info["model_path"]

    def flush_cache(self) -> bool:
        return self.s.post(self.base_url + "/flush_cache", timeout=60).status_code == 200

    def get_server_info(self) -> Any:
        return self.s.get(self.base_url + "/get_server_info", timeout=30).json()

    def cache_prefix(self, prefix_str: str) -> None:
        self.s.post(self.base_url + "/generate", json={"text": prefix_str, "sampling_params": {"max_new_tokens": 0}}, timeout=600)

    def _generate_raw(self, payload: Dict[str, Any]) -> Any:
        r = self.s.post(self.base_url + "/generate", json=payload, timeout=3600)
        if r.status_code != 200:
            raise RuntimeError(f"/generate failed ({r.status_code}): {r.text[:500]}")
        return r.json()

    def _generate_stream(self, payload: Dict[str, Any]) -> Iterator[Dict[str, Any]]:
        with self.s.post(self.base_url + "/generate", json=payload, stream=True, timeout=3600) as r:
            for line in r.iter_lines(decode_unicode=True):
                if line and line.startswith("data:") and "[DONE]" not in line:
                    yield json.loads(line[5:])


class EngineBackend(BaseBackend):
    """the same frontend semantics against an in-process ``Engine`` (B200-first: no HTTP serialisation on the path)"""

    def __init__(self, engine: Any, chat_template_name: Optional[str] = None):
        super().__init__()
        self.engine = engine
        self.chat_template = get_chat_template(chat_template_name) if chat_template_name else get_chat_template_by_model_path(
            engine.server_args.model_path)
        import threading

        self._lock = threading.Lock()   # the synchronous engine drive is single-threaded

    def _split(self, payload: Dict[str, Any]) -> Dict[str, Any]:
        kw = dict(payload)
        kw["prompt"] = kw.pop("text")
        return kw

    def _generate_raw(self, payload: Dict[str, Any]) -> Any:
        with self._lock:
            return self.engine.generate(**self._split(payload))

    def _generate_stream(self, payload: Dict[str, Any]) -> Iterator[Dict[str, Any]]:
        kw = self._split(payload)
        with self._lock:
            yield from self.engine.generate(**kw)

    def cache_prefix(self, prefix_str: str) -> None:
        self._generate_raw({"text": prefix_str, "sampling_params": {"max_new_tokens": 0}})

    def flush_cache(self) -> bool:
        return self.engine.flush_cache()

    def get_server_info(self) -> Any:
        return self.engine.get_server_info()


def _wait_http(url: str, timeout: float, proc: Optional[subprocess.Popen] = None) -> None:
    import requests

    t0 = time.time()
    while time.time() - t0 < timeout:
        if proc is not None and proc.poll() is not None:
            raise RuntimeError(f"server process exited with code {proc.returncode}")
        try:
            if requests.get(url + "/health", timeout=2).status_code == 200:
                return
        except Exception:
            pass
        time.sleep(0.3)
    raise TimeoutError(f"server at {url} did not become healthy in {timeout}s")


class Runtime:
    """spawn ``sglang_b200.launch_server`` as a subprocess and talk to it through a RuntimeEndpoint
    (reference ``Runtime:356``)"""

    def __init__(self, log_level: str = "error", launch_timeout: float = 600.0, **server_kwargs: Any):
        from ...server_args import ServerArgs

        if "port" not in server_kwargs:
            s = socket.socket()
            s.bind(("127.0.0.1", 0))
            server_kwargs["port"] = s.getsockname()[1]
            s.close()
        self.server_args = ServerArgs(log_level=log_level, **server_kwargs)
        argv = [sys.executable, "-m", "sglang_b200.launch_server"]
        import dataclasses

        defaults = ServerArgs.__dataclass_fields__
        for k, v in server_kwargs.items():
            flag = "--" + k.replace("_", "-")
            if isinstance(v, bool):
                if v:
                    argv.append(flag)
            elif isinstance(v, (list, tuple)):
                argv += [flag] + [str(x) for x in v]
            elif v is not None:
                argv += [flag, str(v)]
        argv += ["--log-level", log_level]
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        self.proc = subprocess.Popen(argv, env=env)
        self.url = self.server_args.url()
        atexit.register(self.shutdown)
        try:
            _wait_http(self.url, launch_timeout, self.proc)
        except Exception:
            self.shutdown()
            raise
        self.endpoint = RuntimeEndpoint(self.url, api_key=self.server_args.api_key)

    def shutdown(self) -> None:
        p = getattr(self, "proc", None)
        if p is not None and p.poll() is None:
            p.terminate()
            try:
                p.wait(timeout=15)
            except subprocess.TimeoutExpired:
                p.kill()
        self.proc = None

    def generate(self, prompt: Any, sampling_params: Optional[Dict] = None, **kw: Any) -> Any:
        return self.endpoint._generate_raw({"text": prompt, "sampling_params": sampling_params or {}, **kw})

    def get_tokenizer(self) -> Any:
        from ...utils.tokenizer import get_tokenizer

        return get_tokenizer(self.server_args.tokenizer_path)

    def cache_prefix(self, prefix: str) -> None:
        self.endpoint.cache_prefix(prefix)

    def __del__(self) -> None:
        self.shutdown()
