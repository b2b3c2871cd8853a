"""Import shim: the package directory is named ``rust-local-rag_b200`` (a hyphen is not
importable), so this module turns itself into that package.  ``import rust_local_rag_b200``
and ``from rust_local_rag_b200 import engine`` both work with the repo root on sys.path."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "rust-local-rag_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
