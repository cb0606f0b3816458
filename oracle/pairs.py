"""Label routing and pair combination, restated (test infrastructure; see oracle/__init__.py).

Reference: ``ribodetector/detect.py``
  * single end: ``torch.argmax(output, dim=1)``  :288  (ties → index 0)
  * ``separate_paired_reads``  :616-663
      rrna   : 1 iff both ends are 1, else 0             :620-630
      norrna : 0 iff both ends are 0, else 1             :631-641
      both   : concordant label, else -1 (unclassified)  :642-654
      none   : argmax(logits_r1 + logits_r2)             :655-661  (sum of LOGITS)
  * counts are (non-rRNA, rRNA, unclassified); for pairs they count pairs  :193-206
"""
import numpy as np

MODES = ("none", "rrna", "norrna", "both")


def argmax_labels(logits):
    return np.argmax(np.asarray(logits), axis=1).astype(np.int8)   # first max on ties


def pair_labels(logits1, logits2, mode):
    l1 = argmax_labels(logits1)
    l2 = argmax_labels(logits2)
    if mode == "rrna":
        return ((l1 == 1) & (l2 == 1)).astype(np.int8)
    if mode == "norrna":
        return (~((l1 == 0) & (l2 == 0))).astype(np.int8)
    if mode == "both":
        return np.where(l1 == l2, l1, -1).astype(np.int8)
    if mode == "none":
        s = np.asarray(logits1, dtype=np.float32) + np.asarray(logits2, dtype=np.float32)
        return argmax_labels(s)
    raise ValueError(mode)


def counts(labels):
    labels = np.asarray(labels)
    return np.array([(labels == 0).sum(), (labels == 1).sum(), (labels == -1).sum()],
                    dtype=np.int64)
