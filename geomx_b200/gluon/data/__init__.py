"""``gluon.data`` namespace."""
from . import sampler, vision  # noqa: F401
from .dataloader import DataLoader, default_batchify_fn  # noqa: F401
from .dataset import ArrayDataset, Dataset, RecordFileDataset, SimpleDataset  # noqa: F401
from .sampler import BatchSampler, RandomSampler, Sampler, SequentialSampler  # noqa: F401
