"""Poor man's pyflakes (none is installed here): report names that are loaded but bound nowhere in the enclosing function
scopes, the module or builtins.  Used to vet the GPU-only code paths of bench.py / simt_b200 / tests without a GPU.
    python scripts/undefined_names.py file.py ..."""
import ast
import builtins
import sys


class Scope:
    def __init__(self, parent=None):
        self.parent, self.names = parent, set()


def bound_names(node):
    """names bound directly in this scope's body (not in nested function / class scopes)"""
    out = set()

    def targets(t):
        for n in ast.walk(t):
            if isinstance(n, ast.Name):
                out.add(n.id)

    def visit(n, top=True):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
            if not top:
                return
        if isinstance(n, ast.Lambda) and not top:
            return
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split(".")[0])
        if isinstance(n, (ast.Assign,)):
            for t in n.targets:
                targets(t)
        if isinstance(n, (ast.AugAssign, ast.AnnAssign)):
            targets(n.target)
        if isinstance(n, (ast.For, ast.AsyncFor)):
            targets(n.target)
        if isinstance(n, (ast.With, ast.AsyncWith)):
            for it in n.items:
                if it.optional_vars is not None:
                    targets(it.optional_vars)
        if isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        if isinstance(n, ast.NamedExpr):
            targets(n.target)
        if isinstance(n, (ast.Global, ast.Nonlocal)):
            out.update(n.names)
        if isinstance(n, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
            for g in n.generators:
                targets(g.target)
        for c in ast.iter_child_nodes(n):
            if isinstance(c, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef, ast.Lambda)):
                if not isinstance(c, ast.Lambda):
                    out.add(c.name)
                continue
            visit(c, False)

    visit(node)
    return out


def check(path):
    tree = ast.parse(open(path).read(), path)
    problems = []

    def walk(node, scopes):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            s = set()
            a = node.args
            for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                s.add(arg.arg)
            if not isinstance(node, ast.Lambda):
                for d in node.decorator_list:
                    walk(d, scopes)
                for d in a.defaults + [k for k in a.kw_defaults if k is not None]:
                    walk(d, scopes)
                s |= bound_names(node)
                body = node.body
            else:
                body = [node.body]
            for b in body:
                walk(b, scopes + [s])
            return
        if isinstance(node, ast.ClassDef):
            s = bound_names(node)
            for b in node.bases + node.decorator_list:
                walk(b, scopes)
            for b in node.body:
                # class scope is visible to its own body statements but not to nested functions; approximate: visible
                walk(b, scopes + [s])
            return
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load):
            if not any(node.id in s for s in scopes) and not hasattr(builtins, node.id):
                problems.append((node.lineno, node.id))
        for c in ast.iter_child_nodes(node):
            walk(c, scopes)

    mod = bound_names(tree) | {"__file__", "__name__", "__doc__"}
    walk(tree, [mod])
    return problems


if __name__ == "__main__":
    bad = 0
    for p in sys.argv[1:]:
        for line, name in check(p):
            print(f"{p}:{line}: undefined name {name!r}")
            bad += 1
    sys.exit(1 if bad else 0)
