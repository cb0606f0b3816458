"""``ribodetector_cpu`` — the reference's second command line (``ribodetector/detect_cpu.py:776-825``)
kept as an entry point with ITS numerical convention: reads zero-padded to ``-l`` and the output taken
at the last non-zero one-hot row (``model_cpu.py:29-37,57-62`` — "padded" semantics), so
results can be cross-checked against the published CPU tool.  Same flags as the reference (no ``-d``,
no ``-m``; ``-t`` default 20).  The arithmetic still runs on the B200 kernels — this package has no
CPU path — and, unlike the reference (worker completion order, ``detect_cpu.py:304-311``), output
order equals input order."""
import os

from .detect import Predictor as _Predictor, build_parser, cd
from .parse_config import ConfigParser


class Predictor(_Predictor):
    semantics = "padded"
    gpu_flags = False


def main(argv=None):
    args = build_parser(gpu=False).parse_args(argv)
    config_file = os.path.join(cd, 'config.json') if args.config is None else args.config
    config = ConfigParser.from_json(config_file)
    seq_pred = Predictor(config, args)
    seq_pred.load_model()
    seq_pred.detect()
    return seq_pred


if __name__ == '__main__':
    main()
