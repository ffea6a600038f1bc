from . import msg  # noqa: F401
