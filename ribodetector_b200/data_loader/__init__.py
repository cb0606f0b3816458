from .fastx import get_seq_format, FastxReader, RecordChunk, open_for_write, partition_records, BgzfReader, GzStreamReader, open_text  # noqa: F401
from .fastx_parser import seq_parser  # noqa: F401
from . import seq_encoder  # noqa: F401
