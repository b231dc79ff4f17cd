"""Poor man's pyflakes (no linter in this image): names a function reads as globals that the module never binds.  python tools/undefined_names.py FILE..."""
import builtins
import symtable
import sys


def check(path):
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    bad = []

    def walk(tab):
        for s in tab.get_symbols():
            if tab.get_type() != "module" and s.is_referenced() and s.is_global() and not s.is_assigned():
                n = s.get_name()
                if n not in module_names and not hasattr(builtins, n):
                    bad.append((tab.get_name(), tab.get_lineno(), n))
        for c in tab.get_children():
            walk(c)

    walk(top)
    return bad


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        for scope, line, name in check(p):
            print(f"{p}:{line}: in {scope}: undefined name {name!r}")
            rc = 1
    sys.exit(rc)
