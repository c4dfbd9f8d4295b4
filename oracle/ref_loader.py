"""TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference extension built by oracle/build_ref.py
(oracle/_ref/MuchTree*.so) without any reference .py file in this repo:

  * `dendropy`            -> oracle/newick_ref.py (our stand-in; dendropy is
                             not installable here, see its header)
  * `SuchTree` package    -> an empty in-memory package module
  * `SuchTree.exceptions` -> an in-memory module exposing the four exception
                             classes of suchtree_b200.exceptions (same names
                             and constructor signatures as
                             SuchTree/exceptions.py:2-38)
  * `SuchTree.MuchTree`   -> the compiled reference

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may call this.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_cached = None


def _install_fakes():
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import newick_ref

    if "dendropy" not in sys.modules:
        dp = types.ModuleType("dendropy")
        dp.Tree = newick_ref.Tree
        dp.Node = newick_ref.Node
        dp.Taxon = newick_ref.Taxon
        dp.__standin__ = True
        sys.modules["dendropy"] = dp
    if "SuchTree" not in sys.modules:
        pkg = types.ModuleType("SuchTree")
        pkg.__path__ = []  # mark as package
        sys.modules["SuchTree"] = pkg
    if "SuchTree.exceptions" not in sys.modules:
        repo = os.path.dirname(HERE)
        if repo not in sys.path:
            sys.path.insert(0, repo)
        from suchtree_b200 import exceptions as ours

        ex = types.ModuleType("SuchTree.exceptions")
        for name in ("SuchTreeError", "NodeNotFoundError", "InvalidNodeError", "TreeStructureError"):
            setattr(ex, name, getattr(ours, name))
        sys.modules["SuchTree.exceptions"] = ex
        sys.modules["SuchTree"].exceptions = ex


def load_reference(build_if_missing=True):
    """Return the compiled reference module (SuchTree.MuchTree) or None."""
    global _cached
    if _cached is not None:
        return _cached
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import build_ref

    so = build_ref.ref_so_path()
    if not os.path.exists(so) and build_if_missing:
        so = build_ref.build(verbose=False)
    if not so or not os.path.exists(so):
        return None
    _install_fakes()
    name = "SuchTree.MuchTree"
    loader = importlib.machinery.ExtensionFileLoader(name, so)
    spec = importlib.util.spec_from_file_location(name, so, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    pkg = sys.modules["SuchTree"]
    pkg.MuchTree = mod
    pkg.SuchTree = mod.SuchTree
    pkg.SuchLinkedTrees = mod.SuchLinkedTrees
    for name in ("SuchTreeError", "NodeNotFoundError", "InvalidNodeError", "TreeStructureError"):
        setattr(pkg, name, getattr(pkg.exceptions, name))
    _cached = mod
    return mod
