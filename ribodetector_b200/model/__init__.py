from .model import SeqModel, ReadBatch  # noqa: F401
