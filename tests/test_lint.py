"""No linter ships in the image; tools/lint_names.py finds the missing-import / typo class of mistake (names loaded but bound
nowhere in the file) and dead imports.  Keeps the whole tree - package, oracle, tests, tools, bench - clean of both."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_undefined_names_or_unused_imports(capsys):
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    try:
        import lint_names
    finally:
        sys.path.pop(0)
    bad = lint_names.main([ROOT])
    assert bad == 0, capsys.readouterr().out
