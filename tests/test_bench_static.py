"""bench.py and __graft_entry__.py only run end to end on a GPU box: keep what can be checked here checked here.
Every name they load must be defined somewhere in the file (a missing import once survived until the GPU run)."""
import ast
import builtins
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def undefined_names(path):
    tree = ast.parse(open(path).read())
    defined = set(dir(builtins)) | {"__file__", "__name__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            defined.add(n.name)
        elif isinstance(n, ast.Import):
            defined.update((a.asname or a.name).split(".")[0] for a in n.names)
        elif isinstance(n, ast.ImportFrom):
            defined.update(a.asname or a.name for a in n.names)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
            defined.add(n.id)
        elif isinstance(n, ast.arg):
            defined.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            defined.add(n.name)
    return {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in defined}


@pytest.mark.parametrize("rel", ["bench.py", "__graft_entry__.py", "tools/profile_one.py", "tools/bench_config5.py",
                                 "tools/fuzz_parity.py", "tools/variant_bench.py", "tools/shape_bench.py"])
def test_no_undefined_names(rel):
    assert undefined_names(os.path.join(ROOT, rel)) == set()


def test_bench_refuses_to_run_without_a_gpu():
    """No CPU fallback: on a machine without CUDA the b200 arm must stop with a message, not produce a number."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this machine has a GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CUDA device" in r.stderr
