from .fastx import get_seq_format, FastxReader, RecordChunk, open_for_write, partition_records  # noqa: F401
