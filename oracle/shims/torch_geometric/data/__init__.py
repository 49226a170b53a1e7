from . import makedirs  # noqa: F401
