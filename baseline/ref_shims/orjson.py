"""json-backed stand-in for the `orjson` wheel (not in this offline image).  Reference-arm only: lives on the sys.path that
bench.py --impl reference builds, never on the product's path."""
import dataclasses
import json as _json

OPT_APPEND_NEWLINE = 1 << 10
OPT_INDENT_2 = 1 << 0
OPT_NAIVE_UTC = 1 << 1
OPT_NON_STR_KEYS = 1 << 2
OPT_OMIT_MICROSECONDS = 1 << 3
OPT_PASSTHROUGH_DATACLASS = 1 << 11
OPT_PASSTHROUGH_DATETIME = 1 << 9
OPT_PASSTHROUGH_SUBCLASS = 1 << 8
OPT_SERIALIZE_DATACLASS = 0
OPT_SERIALIZE_NUMPY = 1 << 4
OPT_SERIALIZE_UUID = 0
OPT_SORT_KEYS = 1 << 5
OPT_STRICT_INTEGER = 1 << 6
OPT_UTC_Z = 1 << 7

JSONDecodeError = _json.JSONDecodeError


class JSONEncodeError(TypeError):
    pass


class Fragment:
    def __init__(self, contents):
        self.contents = contents


def _default_chain(user_default):
    def d(o):
        if dataclasses.is_dataclass(o) and not isinstance(o, type):
            return dataclasses.asdict(o)
        try:
            import numpy as np

            if isinstance(o, np.ndarray):
                return o.tolist()
            if isinstance(o, np.generic):
                return o.item()
        except ImportError:
            pass
        if hasattr(o, "isoformat"):
            return o.isoformat()
        if isinstance(o, (set, frozenset, tuple)):
            return list(o)
        if isinstance(o, bytes):
            return o.decode("utf-8", "replace")
        if isinstance(o, Fragment):
            c = o.contents
            return _json.loads(c if isinstance(c, str) else c.decode())
        if user_default is not None:
            return user_default(o)
        raise JSONEncodeError(f"Type is not JSON serializable: {type(o).__name__}")

    return d


def dumps(obj, default=None, option=None) -> bytes:
    option = option or 0
    kw = {"separators": (",", ":"), "ensure_ascii": False, "default": _default_chain(default)}
    if option & OPT_INDENT_2:
        kw["indent"] = 2
        kw["separators"] = (",", ": ")
    if option & OPT_SORT_KEYS:
        kw["sort_keys"] = True
    s = _json.dumps(obj, **kw)
    if option & OPT_APPEND_NEWLINE:
        s += "\n"
    return s.encode("utf-8")


def loads(data):
    if isinstance(data, (bytes, bytearray, memoryview)):
        data = bytes(data).decode("utf-8")
    return _json.loads(data)
