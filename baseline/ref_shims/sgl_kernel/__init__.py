"""Import-surface stand-in for the `sgl_kernel` wheel (no wheel in this offline image; its CMake build needs network FetchContent).

Reference-arm only (bench.py --impl reference puts baseline/ref_shims on sys.path; the product never imports this).

Every attribute / submodule resolves, so the reference's `from sgl_kernel import X` statements succeed.  The handful of ops the
bf16 Llama serving path can actually reach when the reference's own in-tree JIT kernels (sglang/kernels/jit, compiled on the box
through tvm-ffi) are unavailable are backed by LIBRARY kernels that ship in this image (vLLM's precompiled `_C` ops, FlashInfer);
everything else raises on call, loudly, so a silently wrong path cannot be timed.
"""
import importlib.abc
import importlib.machinery
import sys
import types

__version__ = "0.0.0+shim"


def _unavailable(name):
    def f(*a, **k):
        raise NotImplementedError(f"sgl_kernel.{name} is not available in this offline image (baseline/ref_shims stand-in)")

    f.__name__ = name.rsplit(".", 1)[-1]
    f.__sgl_shim_unavailable__ = True
    return f


# ------------------------------------------------------------------ library-backed ops (same math as sgl_kernel's own kernels)
def _vllm_ops():
    import vllm._custom_ops as ops  # noqa: F401  (loads vllm._C)
    import torch

    return torch.ops._C


def rmsnorm(input, weight, eps=1e-6, out=None, enable_pdl=None):
    import torch

    if out is None:
        out = torch.empty_like(input)
    _vllm_ops().rms_norm(out, input, weight, eps)
    return out


def fused_add_rmsnorm(input, residual, weight, eps=1e-6, enable_pdl=None):
    _vllm_ops().fused_add_rms_norm(input, residual, weight, eps)


def gemma_rmsnorm(input, weight, eps=1e-6, out=None, enable_pdl=None):
    import torch

    x = input.float()
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + weight.float())
    if out is None:
        return y.to(input.dtype)
    out.copy_(y.to(input.dtype))
    return out


def gemma_fused_add_rmsnorm(input, residual, weight, eps=1e-6, enable_pdl=None):
    residual.add_(input)
    input.copy_(gemma_rmsnorm(residual, weight, eps))


def _act_and_mul(opname):
    def f(input, out=None):
        import torch

        if out is None:
            out = torch.empty(*input.shape[:-1], input.shape[-1] // 2, dtype=input.dtype, device=input.device)
        getattr(_vllm_ops(), opname)(out, input)
        return out

    f.__name__ = opname
    return f


silu_and_mul = _act_and_mul("silu_and_mul")
gelu_and_mul = _act_and_mul("gelu_and_mul")
gelu_tanh_and_mul = _act_and_mul("gelu_tanh_and_mul")


def rotary_embedding(positions, query, key, head_size, cos_sin_cache, is_neox=True):
    _vllm_ops().rotary_embedding(positions, query, key, head_size, cos_sin_cache, is_neox)


def merge_state_v2(v_a, s_a, v_b, s_b, v_merged=None, s_merged=None):
    import torch

    m = torch.maximum(s_a, s_b)
    wa, wb = torch.exp(s_a - m), torch.exp(s_b - m)
    den = wa + wb
    v = (v_a.float() * (wa / den).unsqueeze(-1) + v_b.float() * (wb / den).unsqueeze(-1)).to(v_a.dtype)
    s = m + torch.log(den)
    if v_merged is not None:
        v_merged.copy_(v)
        s_merged.copy_(s)
        return v_merged, s_merged
    return v, s


def top_k_renorm_prob(probs, top_k):
    import flashinfer.sampling as fs

    return fs.top_k_renorm_probs(probs, top_k)


def top_p_renorm_prob(probs, top_p):
    import flashinfer.sampling as fs

    return fs.top_p_renorm_probs(probs, top_p)


def min_p_sampling_from_probs(probs, min_p, *a, **k):
    import flashinfer.sampling as fs

    return fs.min_p_sampling_from_probs(probs, min_p)


def top_k_top_p_sampling_from_probs(probs, top_k, top_p, filter_apply_order="top_k_first", *a, **k):
    import flashinfer.sampling as fs

    return fs.top_k_top_p_sampling_from_probs(probs, top_k, top_p, filter_apply_order=filter_apply_order)


def weak_ref_tensor(t):
    return t


class _StubModule(types.ModuleType):
    """sgl_kernel.<anything>: attributes resolve to the top-level implementation if there is one, else to a raising stub"""

    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        top = sys.modules["sgl_kernel"]
        if name in top.__dict__:
            return top.__dict__[name]
        if name and name[0].isupper():      # classes used in annotations / isinstance (ScalarType, FlashMLASchedMeta, ...)
            return type(name, (), {"__sgl_shim_unavailable__": True})
        return _unavailable(f"{self.__name__.split('.', 1)[-1]}.{name}")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.startswith("sgl_kernel."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        if module.__name__ == "sgl_kernel.utils":
            module.is_arch_support_pdl = lambda: True
            module._to_tensor_scalar_tuple = lambda x: (None, x) if not hasattr(x, "shape") else (x, 0)
        if module.__name__ == "sgl_kernel.version":
            module.__version__ = __version__
        if module.__name__ == "sgl_kernel.scalar_type":
            class _ST:
                def __getattr__(self, n):
                    return n
            module.scalar_types = _ST()


sys.meta_path.insert(0, _Finder())


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    if name in ("flash_ops", "flashmla_ops", "infllm_ops", "spatial", "metal"):
        return importlib.import_module(f"sgl_kernel.{name}")
    return _unavailable(name)
