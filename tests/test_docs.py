"""the generated reference pages (docs/*.md) must match the code: regenerate with ``python tools/gen_docs.py``"""
import os
import subprocess
import sys


def test_generated_docs_are_current():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_docs.py"), "--check"], capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stdout + r.stderr


def test_plugin_registration_before_builtins_keeps_builtins():
    code = ("from sglang_b200.models import registry as R\n"
            "class X: pass\n"
            "R.register_model('MyArch', X)\n"
            "assert R.get_model_class('LlamaForCausalLM').__name__ == 'LlamaForCausalLM' and R.get_model_class('MyArch') is X\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stderr
