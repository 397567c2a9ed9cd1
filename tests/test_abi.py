"""The C-ABI library loads and exports every symbol include/lipread_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lipread_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from multimodal_lipread_b200 import _lib
    syms = _declared_symbols()
    assert "lr_logmel_fwd" in syms and "lr_version" in syms
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/lipread_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES and the header disagree"


def test_version_and_host_only_calls():
    from multimodal_lipread_b200 import _lib
    header = open(os.path.join(ROOT, "include", "lipread_b200.h")).read()
    assert _lib.lib.lr_version() == int(re.search(r"#define\s+LR_ABI_VERSION\s+(\d+)", header).group(1)) >= 2
    assert _lib.lib.lr_logmel_plan_bytes() % 16 == 0 and _lib.lib.lr_logmel_plan_bytes() > 9000
    assert _lib.launch_count() >= 0


def test_argument_errors_do_not_touch_the_gpu():
    from multimodal_lipread_b200 import _lib
    rc = _lib.lib.lr_logmel_fwd(None, None, None, 4, 117, 0, None)
    assert rc == -1 and b"null" in _lib.lib.lr_last_error()
    rc = _lib.lib.lr_logmel_fwd(16, 16, 16, 4, 200, 0, None)
    assert rc == -1 and b"n_out" in _lib.lib.lr_last_error()
    rc = _lib.lib.lr_logmel_fwd(16, 16, 24, 4, 117, 0, None)
    assert rc == -2
    assert _lib.lib.lr_logmel_fwd(None, None, None, 0, 117, 0, None) == 0     # empty batch is a no-op


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from multimodal_lipread_b200 import ops
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.logmel(torch.zeros(1, 20000), torch.zeros(16, dtype=torch.uint8), 117, 0)


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU arm may touch it.  The
    package and the B200 bench workloads must not import it (a product path through the oracle would void parity)."""
    import ast
    import glob
    files = glob.glob(os.path.join(ROOT, "multimodal_lipread_b200", "*.py")) + [os.path.join(ROOT, "bench_workloads.py")]
    assert len(files) > 10
    for path in files:
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f"{path} imports the oracle"
    # bench_checks.py (dp_parity / torch_gpu_baseline: checker legs, never timed as the product) is the one other user
    assert "oracle" in open(os.path.join(ROOT, "bench_checks.py")).read()
    # bench.py: the oracle appears only inside cpu_reference()
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle") for n in ast.walk(fn))
        assert uses == (fn.name == "cpu_reference") or not uses, fn.name
    src = open(os.path.join(ROOT, "multimodal_lipread_b200", "_lib.py")).read()
    assert "raise ImportError" in src                      # a missing library fails loudly; there is no fallback
