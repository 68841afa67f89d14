"""madrona-learn learner hot path, B200-native (sm_100a kernels behind the reference's API).

Public surface mirrors /root/reference/src/madrona_learn/__init__.py:1-45 for the hot path.
"""
from . import _lib  # noqa: F401
from . import kernels  # noqa: F401
