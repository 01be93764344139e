"""Import shim: the package sources live in ``rag-cobweb_b200/`` (the directory name the
project layout prescribes, which is not a valid Python identifier).  Importing
``rag_cobweb_b200`` extends ``__path__`` to that directory and runs its ``_pkg_init``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rag-cobweb_b200")
__path__.append(_real)

from ._pkg_init import *  # noqa: F401,F403,E402
from ._pkg_init import __all__  # noqa: E402
