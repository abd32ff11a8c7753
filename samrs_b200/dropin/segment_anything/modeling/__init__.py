from .sam import Sam  # noqa: F401
