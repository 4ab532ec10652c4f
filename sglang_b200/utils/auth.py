"""Request authorisation policy shared by the HTTP, WebSocket and gRPC front ends.

Deny by default: once ``--api-key`` (or ``--admin-api-key``) is configured, EVERY route needs a bearer token except liveness probes
(``/health*``, ``/ping``) and the Prometheus scrape (``/metrics``).  Control-plane routes are classed explicitly:

  ADMIN         needs the admin key when one is configured, otherwise the API key, otherwise open   (cache flush, profiling, sessions …)
  ADMIN_FORCE   needs the admin key; WITHOUT an admin key the route is refused (403) even for API-key holders
                (weight mutation / weight read-back / memory release / fault injection)

Behavioural parity: reference ``python/sglang/srt/utils/auth.py`` (``decide_request_auth``: NORMAL / ADMIN_OPTIONAL / ADMIN_FORCE) — the
classification here is by route table instead of per-endpoint decorators so that the gRPC method names map onto the same table.
Tokens are compared with ``secrets.compare_digest``.
"""
from __future__ import annotations

import secrets
from typing import Optional, Tuple

OPEN_PREFIXES: Tuple[str, ...] = ("/health", "/metrics", "/ping")

ADMIN_ROUTES = frozenset({
    "/flush_cache", "/start_profile", "/stop_profile", "/pause_generation", "/continue_generation", "/configure_logging",
    "/set_trace_level", "/abort_request", "/open_session", "/close_session", "/load_lora_adapter", "/unload_lora_adapter",
    "/hicache/storage-backend", "/eplb_rebalance", "/build_expert_backup", "/elastic_ep_recover", "/kv_events", "/get_server_info",
})
ADMIN_FORCE_ROUTES = frozenset({
    "/update_weights_from_disk", "/update_weights_from_tensor", "/update_weights_from_distributed", "/init_weights_update_group",
    "/destroy_weights_update_group", "/send_weights_to_remote_instance", "/get_weights_by_name", "/weights_checker",
    "/release_memory_occupation", "/resume_memory_occupation", "/slow_down",
})

NORMAL, ADMIN, ADMIN_FORCE = "normal", "admin", "admin_force"


def classify(path: str) -> str:
    p = path.rstrip("/") or "/"
    if p in ADMIN_FORCE_ROUTES:
        return ADMIN_FORCE
    if p in ADMIN_ROUTES:
        return ADMIN
    return NORMAL


def bearer(header: Optional[str]) -> Optional[str]:
    if not header:
        return None
    parts = header.split(" ", 1)
    if len(parts) != 2 or parts[0].lower() != "bearer":
        return None
    return parts[1].strip()


def _eq(token: Optional[str], expected: Optional[str]) -> bool:
    return bool(token) and bool(expected) and secrets.compare_digest(token.encode(), expected.encode())


def decide(method: str, path: str, authorization: Optional[str], api_key: Optional[str], admin_api_key: Optional[str]) -> Tuple[bool, int, str]:
    """-> (allowed, http status when refused, message)"""
    if method.upper() == "OPTIONS" or path.startswith(OPEN_PREFIXES):
        return True, 200, ""
    level = classify(path)
    tok = bearer(authorization)
    if level == ADMIN_FORCE:
        if not admin_api_key:
            if not api_key:
                return True, 200, ""          # no keys configured at all: a private deployment, nothing to compare against
            return False, 403, "this route needs --admin-api-key to be configured"
        return (True, 200, "") if _eq(tok, admin_api_key) else (False, 401, "admin API key required")
    if level == ADMIN:
        if admin_api_key:
            return (True, 200, "") if _eq(tok, admin_api_key) else (False, 401, "admin API key required")
        if api_key:
            return (True, 200, "") if _eq(tok, api_key) else (False, 401, "invalid API key")
        return True, 200, ""
    if api_key:
        ok = _eq(tok, api_key) or _eq(tok, admin_api_key)
        return (True, 200, "") if ok else (False, 401, "invalid API key")
    return True, 200, ""


# gRPC method name -> the HTTP route whose policy it shares (anything not listed is a NORMAL data-plane call)
GRPC_ADMIN_EQUIV = {
    "FlushCache": "/flush_cache", "Abort": "/abort_request", "PauseGeneration": "/pause_generation",
    "ContinueGeneration": "/continue_generation", "StartProfile": "/start_profile", "StopProfile": "/stop_profile",
    "UpdateWeightsFromDisk": "/update_weights_from_disk", "LoadLoRAAdapter": "/load_lora_adapter",
    "UnloadLoRAAdapter": "/unload_lora_adapter", "ReleaseMemoryOccupation": "/release_memory_occupation",
    "ResumeMemoryOccupation": "/resume_memory_occupation", "GetServerInfo": "/get_server_info", "OpenSession": "/open_session",
    "CloseSession": "/close_session",
}


def decide_grpc(method_name: str, authorization: Optional[str], api_key: Optional[str], admin_api_key: Optional[str]) -> Tuple[bool, str]:
    short = method_name.rsplit("/", 1)[-1]
    if short in ("HealthCheck", "Health", "Check"):
        return True, ""
    path = GRPC_ADMIN_EQUIV.get(short, "/v1/grpc")
    ok, _, msg = decide("POST", path, authorization, api_key, admin_api_key)
    return ok, msg
