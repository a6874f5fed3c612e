"""tools/undefined_names.py -- poor man's pyflakes (the image has none): report names that are loaded in a function or
module but never bound anywhere in that module (imports, defs, assignments, args, comprehension targets, builtins).
Catches the NameErrors that only a GPU-side code path would otherwise reveal."""
import ast
import builtins
import sys


def bound_names(tree):
    names = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(node.name)
            if not isinstance(node, ast.ClassDef):
                for a in node.args.args + node.args.kwonlyargs + node.args.posonlyargs:
                    names.add(a.arg)
                for a in (node.args.vararg, node.args.kwarg):
                    if a:
                        names.add(a.arg)
        elif isinstance(node, ast.Lambda):
            for a in node.args.args + node.args.kwonlyargs:
                names.add(a.arg)
            for a in (node.args.vararg, node.args.kwarg):
                if a:
                    names.add(a.arg)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            names.add(node.id)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            names.add(node.name)
        elif isinstance(node, (ast.Global, ast.Nonlocal)):
            names.update(node.names)
    return names


bad = 0
for path in sys.argv[1:]:
    tree = ast.parse(open(path).read(), path)
    known = bound_names(tree)
    for node in ast.walk(tree):
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load) and node.id not in known:
            print(f"{path}:{node.lineno}: undefined name {node.id!r}")
            bad += 1
sys.exit(1 if bad else 0)
