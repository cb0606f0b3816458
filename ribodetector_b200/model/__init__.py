from .model import SeqModel  # noqa: F401
