from .tucker import Tucker  # noqa: F401
