"""Static sweep of the Python sources: every global name a function reads is defined at module level, imported, or a builtin.  The GPU-only code paths
(NCCL stages of the h x w path, ctypes wrappers) are not executed by the CPU suite, so a stale name there would surface only on the GPU box -- as the
`L.launch_stream` NameError of round 2 did (makani_b200/distributed, fixed)."""
import builtins
import glob
import os
import symtable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _undefined(path):
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    out = []

    def walk(table):
        for child in table.get_children():
            for s in child.get_symbols():
                if s.is_global() and s.is_referenced() and not s.is_assigned():
                    n = s.get_name()
                    if n not in module_names and not hasattr(builtins, n) and n not in ("__file__", "__name__"):
                        out.append((child.get_name(), child.get_lineno(), n))
            walk(child)

    walk(top)
    return out


def test_no_undefined_global_names():
    files = glob.glob(os.path.join(ROOT, "makani_b200", "**", "*.py"), recursive=True) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += glob.glob(os.path.join(ROOT, "scripts", "*.py")) + glob.glob(os.path.join(ROOT, "oracle", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "*.py"))
    bad = {os.path.relpath(f, ROOT): u for f in files if (u := _undefined(f))}
    assert not bad, bad


def test_product_path_does_not_touch_the_oracle_or_torch_fft():
    """the CUDA product path must not route through the CPU oracle or through torch.fft / einsum"""
    import re

    for f in glob.glob(os.path.join(ROOT, "makani_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        code = "\n".join(line.split("#", 1)[0] for line in re.sub(r'"""(.|\n)*?"""', "", src).splitlines())
        assert not re.search(r"^\s*(from|import)\s+oracle\b", code, flags=re.M), f
        assert "torch.fft" not in code and "einsum(" not in code, f
