"""Model families: the reference's demo CNN (+ its graph-captured HiPS training engine)."""
from .cnn import CNN_PARAM_SHAPES, HipsCNNTrainStep, build_cnn  # noqa: F401
