#!/usr/bin/env python
"""Offline stand-in for pyflakes (no linter is installed in the image): reports names that are loaded somewhere in a module but
bound nowhere in it (module scope, any function / class / comprehension / except / with / import) and are not builtins, and
imports that are never used.  Coarse on purpose - scopes are merged per file - so it only finds the typo / missing-import class
of mistake.   usage: python tools/lint_names.py [paths...]"""
import ast
import builtins
import os
import sys


def check(path):
    src = open(path).read()
    tree = ast.parse(src, path)
    bound, loaded, imported = set(dir(builtins)) | {'__file__', '__name__', '__doc__', '__path__'}, {}, {}
    star = False
    for n in ast.walk(tree):
        if isinstance(n, ast.Name):
            if isinstance(n.ctx, ast.Load):
                loaded.setdefault(n.id, n.lineno)
            else:
                bound.add(n.id)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(n.name)
            if not isinstance(n, ast.ClassDef):
                a = n.args
                for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                    bound.add(arg.arg)
        elif isinstance(n, ast.Lambda):
            a = n.args
            for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                bound.add(arg.arg)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for al in n.names:
                if al.name == '*':
                    star = True
                    continue
                name = (al.asname or al.name).split('.')[0]
                bound.add(name)
                imported.setdefault(name, n.lineno)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            bound.update(n.names)
        elif isinstance(n, ast.MatchAs) and n.name:
            bound.add(n.name)
    out = []
    if not star:
        out += [(ln, f'undefined name {k!r}') for k, ln in loaded.items() if k not in bound]
    exported = set()
    for n in ast.walk(tree):                                     # names re-exported through __all__ count as used
        if isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == '__all__' for t in n.targets):
            exported |= {e.value for e in getattr(n.value, 'elts', []) if isinstance(e, ast.Constant)}
    attr_roots = {n.value.id for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name)}
    if os.path.basename(path) != '__init__.py':
        for k, ln in imported.items():
            if k not in loaded and k not in attr_roots and k not in exported and k != 'annotations':
                line = src.splitlines()[ln - 1]
                if 'noqa' not in line:
                    out.append((ln, f'unused import {k!r}'))
    return sorted(out)


def main(paths):
    files = []
    for p in paths:
        if os.path.isdir(p):
            for d, _, fs in os.walk(p):
                if any(s in d for s in ('.git', '__pycache__', 'gpurun_out', '_ref')):
                    continue
                files += [os.path.join(d, f) for f in fs if f.endswith('.py')]
        else:
            files.append(p)
    bad = 0
    for f in sorted(files):
        for ln, msg in check(f):
            print(f'{f}:{ln}: {msg}')
            bad += 1
    return bad


if __name__ == '__main__':
    sys.exit(1 if main(sys.argv[1:] or ['.']) else 0)
