"""Tiny undefined-name finder (there is no pyflakes / ruff in this image): flags Name loads that no enclosing scope binds.

    python tools/undefined_names.py <files or directories ...>

Why it exists: the GPU-only Python branches (engine, CUDA paths of the distributed optimizer, bench.py) are never executed
by the CPU test suite, so a typo there would only show up on the GPU box.  `tests/test_static.py` runs it over the tree.
Deliberately simple: module / function / class / comprehension scopes, imports, globals; star-imports disable the check
for that file."""
import ast, builtins, sys, os

BUILTINS = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__path__", "__spec__", "__package__", "__builtins__", "__class__"}

class Scope:
    def __init__(self, node, parent, kind):
        self.node, self.parent, self.kind = node, parent, kind
        self.bound, self.globals_ = set(), set()

def bind_targets(t, scope):
    for n in ast.walk(t):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            scope.bound.add(n.id)

def collect(node, scope, scopes):
    for child in ast.iter_child_nodes(node):
        if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            if not isinstance(child, ast.Lambda):
                scope.bound.add(child.name)
                for d in child.decorator_list: collect_expr(d, scope, scopes)
            s = Scope(child, scope, "function"); scopes.append(s)
            a = child.args
            for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                s.bound.add(arg.arg)
            for d in a.defaults + [k for k in a.kw_defaults if k is not None]: collect_expr(d, scope, scopes)
            if isinstance(child, ast.Lambda):
                collect_expr(child.body, s, scopes)
            else:
                collect(ast.Module(body=child.body, type_ignores=[]), s, scopes)
        elif isinstance(child, ast.ClassDef):
            scope.bound.add(child.name)
            s = Scope(child, scope, "class"); scopes.append(s)
            for b in child.bases + child.decorator_list + [k.value for k in child.keywords]: collect_expr(b, scope, scopes)
            collect(ast.Module(body=child.body, type_ignores=[]), s, scopes)
        elif isinstance(child, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
            s = Scope(child, scope, "function"); scopes.append(s)
            for g in child.generators: bind_targets(g.target, s)
            collect(child, s, scopes)
        else:
            if isinstance(child, (ast.Import, ast.ImportFrom)):
                for al in child.names:
                    scope.bound.add((al.asname or al.name).split(".")[0])
            elif isinstance(child, ast.Global):
                scope.globals_.update(child.names)
            elif isinstance(child, ast.Nonlocal):
                scope.bound.update(child.names)
            elif isinstance(child, ast.ExceptHandler) and child.name:
                scope.bound.add(child.name)
            elif isinstance(child, ast.Name):
                if isinstance(child.ctx, (ast.Store, ast.Del)):
                    scope.bound.add(child.id)
                else:
                    scope.loads = getattr(scope, "loads", [])
                    scope.loads.append(child)
            elif isinstance(child, (ast.MatchAs, ast.MatchStar)) and getattr(child, "name", None):
                scope.bound.add(child.name)
            collect(child, scope, scopes)

def collect_expr(e, scope, scopes):
    holder = ast.Expr(value=e); collect(holder, scope, scopes)

def check(path):
    src = open(path).read()
    try: tree = ast.parse(src)
    except SyntaxError as ex: return [f"{path}: SYNTAX {ex}"]
    mod = Scope(tree, None, "module"); scopes = [mod]
    collect(tree, mod, scopes)
    star = any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree))
    out = []
    for s in scopes:
        for n in getattr(s, "loads", []):
            name = n.id; cur = s; found = False
            while cur is not None:
                if cur.kind != "class" or cur is s:
                    if name in cur.bound: found = True; break
                cur = cur.parent
            if not found and name not in BUILTINS and name not in mod.bound and not star:
                out.append(f"{path}:{n.lineno}: undefined name '{name}'")
    return out

def main(roots):
    files = []
    for r in roots:
        if os.path.isfile(r): files.append(r)
        else:
            for d, _, fs in os.walk(r):
                if "build" in d.split(os.sep) or "__pycache__" in d: continue
                files += [os.path.join(d, f) for f in fs if f.endswith(".py")]
    issues = []
    for f in sorted(files): issues += check(f)
    print(f"{len(files)} files checked, {len(issues)} findings")
    for i in issues[:60]: print(" ", i)
    return issues


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1:]) else 0)
