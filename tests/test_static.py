"""Static check of the whole tree for undefined names: the GPU-only Python branches never run in the CPU suite, so a typo
there would otherwise only surface on the GPU box."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _checker():
    spec = importlib.util.spec_from_file_location("undefined_names", os.path.join(ROOT, "tools", "undefined_names.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_no_undefined_names_in_the_tree(capsys):
    roots = [os.path.join(ROOT, p) for p in ("distributed-deep-learning-workshop_b200", "b200ddl", "examples", "benchmarks",
                                             "baseline", "tests", "tools", "bench.py", "__graft_entry__.py")]
    issues = _checker().main(roots)
    assert not issues, "\n".join(issues)


def test_the_checker_catches_planted_errors(tmp_path):
    (tmp_path / "planted.py").write_text(
        "import os\n\ndef f(a):\n    import json\n    if a:\n        return json.dumps(a)\n    return jsno.dumps(a)\n\n"
        "class K:\n    def m(self, x):\n        return yy + os.sep\n\n"
        "def ok(items, scale=2):\n    return [i * scale for i in items], (lambda z: z + scale)(1)\n")
    issues = _checker().main([str(tmp_path)])
    assert len(issues) == 2 and "jsno" in issues[0] and "yy" in issues[1]
