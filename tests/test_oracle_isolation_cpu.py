"""The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline leg may
import it.  The product package, the tools and bench.py's B200 arm must not."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_imports(path):
    """(lineno, enclosing function or None) of every `import oracle...` / `from oracle... import` in a file."""
    tree = ast.parse(open(path).read())
    hits = []

    def visit(node, fn):
        for child in ast.iter_child_nodes(node):
            name = child.name if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(child, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in child.names):
                hits.append((child.lineno, fn))
            if isinstance(child, ast.ImportFrom) and (child.module or "").split(".")[0] == "oracle":
                hits.append((child.lineno, fn))
            visit(child, name)

    visit(tree, None)
    return hits


def test_product_and_tools_do_not_import_the_oracle():
    bad = []
    for sub in ("multimodal_edema_prediction_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith(".py"):
                    p = os.path.join(dirpath, f)
                    bad += [(os.path.relpath(p, ROOT), h) for h in _oracle_imports(p)]
    assert not bad, bad


def test_bench_imports_the_oracle_only_in_the_reference_arm():
    hits = _oracle_imports(os.path.join(ROOT, "bench.py"))
    # the reference arm = run_reference and its two step builders (_oracle_port; _reference_in_place imports the reference's
    # own files through oracle/shims via sys.path, not the oracle module)
    assert hits and all(fn in ("run_reference", "_oracle_port") for _, fn in hits), hits
    hits = _oracle_imports(os.path.join(ROOT, "__graft_entry__.py"))
    assert all(fn == "smoke" for _, fn in hits), hits
