"""Golden vectors for the record scanner / writer: runs the REFERENCE's own parser
(/root/reference/ribodetector/data_loader/fastx_parser.py:15-55, imported unmodified) and
get_seq_format (seq_encoder.py:21-39) on edge-case texts and stores the inputs with the records it
yields in tests/golden/fastx.json.  Run in the build container (the reference is not on the GPU
box):  python oracle/gen_golden_fastx.py"""
import io
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
bio, seqm = types.ModuleType("Bio"), types.ModuleType("Bio.Seq")
seqm.Seq = object
sys.modules["Bio"], sys.modules["Bio.Seq"] = bio, seqm
from ribodetector.data_loader.fastx_parser import seq_parser            # noqa: E402
from ribodetector.data_loader.seq_encoder import get_seq_format         # noqa: E402

CASES = {
    "fq_plain": ("fastq", "@r1 first read\nACGTNNacgtRYKM\n+\nIIIIIIIIIIIIII\n@r2\nGGGGCCCC\n+r2\n@@@@!!!!\n"),
    "fq_crlf_trailing_space": ("fastq", "@r1\r\nACGT  \r\n+\r\nIIII\r\n@r2 \nTT\t\n+ \nII\n"),
    "fq_truncated_final_record": ("fastq", "@r1\nACGT\n+\nIIII\n@r2\nACG\n+\n"),
    "fq_no_final_newline": ("fastq", "@r1\nACGT\n+\nIIII\n@r2\nAC\n+\n##"),
    "fq_quality_starts_with_at": ("fastq", "@r1\nACGT\n+\n@III\n@r2\nUUUU\n+\n@@@@\n"),
    "fq_empty": ("fastq", ""),
    "fa_multiline_lowercase": ("fasta", ">s1 desc\nacgtn\nACGU\n>s2\nGGGG\n"),
    "fa_blank_lines_and_spaces": ("fasta", "\n>s1\n  acgt  \n\nNN\n\n>s2\n tt\n"),
    "fa_empty_record_in_the_middle": ("fasta", ">s1\nAC\n>s2\n>s3\nGG\n"),
    "fa_trailing_header_dropped": ("fasta", ">s1\nAC\n>s2\n"),
    "fa_no_final_newline": ("fasta", ">s1\nAC\n>s2\nGT"),
    "fa_sequence_before_first_header": ("fasta", "ACGT\n>s1\nGG\n>s2\nTT\n"),
    "fa_crlf": ("fasta", ">s1\r\nacgt\r\nAC\r\n>s2\r\nGG\r\n"),
}
NAMES = ["a.fq", "a.fastq", "a.fa", "a.fasta", "a.fna", "a.fas", "a.fq.gz", "a.fasta.gz", "dir.x/b.fastq.gz",
         "a.txt", "a.fq.gzip", "a.fq.bz2", "a.fastq.xz", "a", "a.FQ"]


def main():
    out = {"cases": {}, "formats": {}}
    for name, (typ, text) in CASES.items():
        recs = [list(r) for r in seq_parser(io.StringIO(text, newline=None), typ)]
        out["cases"][name] = {"type": typ, "text": text, "records": recs}
    for n in NAMES:
        try:
            out["formats"][n] = get_seq_format(n)
        except ValueError as e:
            out["formats"][n] = "ValueError"
    path = os.path.join(ROOT, "tests", "golden", "fastx.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, {k: len(v["records"]) for k, v in out["cases"].items()}, out["formats"])


if __name__ == "__main__":
    main()
