from .conv import MessagePassing  # noqa: F401
