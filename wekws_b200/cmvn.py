"""Host-side CMVN constant preparation (mirror of the reference's wekws/utils/cmvn.py).

``load_cmvn`` follows wekws/utils/cmvn.py:23-45 (json stats -> [mean, 1/std]) and
``load_kaldi_cmvn`` follows :48-93 (Kaldi nnet text: AddShift / Rescale / Splice).
Both return a float64 numpy array of shape (2, dim) exactly like the reference, which
init_model then converts with ``torch.from_numpy(...).float()`` (kws_model.py:104-108).
"""
import json
import math
import re

import numpy as np


def load_cmvn(json_cmvn_file):
    with open(json_cmvn_file) as f:
        stats = json.load(f)
    count = stats["frame_num"]
    mean, istd = [], []
    for s, v in zip(stats["mean_stat"], stats["var_stat"]):
        mu = s / count
        var = max(v / count - mu * mu, 1.0e-20)
        mean.append(mu)
        istd.append(1.0 / math.sqrt(var))
    return np.array([mean, istd])


def _bracket_floats(line):
    inner = re.findall(r"[\[](.*?)[\]]", line)[0]
    return [float(x) for x in inner.strip().split(" ")]


def load_kaldi_cmvn(cmvn_file):
    with open(cmvn_file) as f:
        lines = f.readlines()
    means = scales = None
    copies = None
    for i, line in enumerate(lines):
        head = line.strip().split(" ")
        if "AddShift" in line:
            assert len(head) == 3
            means = [0.0 - v for v in _bracket_floats(lines[i + 1])]
            assert len(means) == int(head[1])
        elif "Rescale" in line:
            assert len(head) == 3
            scales = _bracket_floats(lines[i + 1])
            assert len(scales) == int(head[1])
        elif "Splice" in line:
            assert len(head) == 3
            splice = _bracket_floats(lines[i + 1])
            assert len(splice) * int(head[2]) == int(head[1])
            copies = len(splice)
    return np.tile(np.array([means, scales]), (1, copies))
