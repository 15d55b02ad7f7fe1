from .data_sampler import EnlargedSampler  # noqa: F401
from .synthetic import SyntheticPairedDataset, synth_pair  # noqa: F401
from .prefetch import CUDAPrefetcher  # noqa: F401
