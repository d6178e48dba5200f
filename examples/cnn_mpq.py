#!/usr/bin/env python
"""Mixed-Precision Quantisation: tensors above MXNET_KVSTORE_SIZE_LOWER_BOUND travel fp32 + Bi-Sparse, the small ones dense float16."""
import os

from cnn_fp16 import run
from common import make_parser

if __name__ == "__main__":
    a = make_parser(extra=("bcr",)).parse_args()
    bound = int(float(os.getenv("MXNET_KVSTORE_SIZE_LOWER_BOUND", 2e5)))
    run(a, lambda arr: arr.size <= bound)
